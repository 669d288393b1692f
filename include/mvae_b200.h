/*
 * libmvae_b200.so -- C-ABI of the B200 (sm_100a) multiscale-VAE training-step kernels.
 *
 * The reference (NikolasMarkou/multiscale_variational_autoencoder) has NO native / FFI boundary: all of its
 * arithmetic is executed by TensorFlow ops reached through Keras layers.  Each entry point below therefore
 * cites the Keras call site(s) in the reference whose TensorFlow op(s) it replaces (file:line under
 * /root/reference).  The Python host (multiscale_variational_autoencoder_b200/*.py) is the only caller.
 *
 * Conventions
 *   - activations NHWC fp32, contiguous; weights in Keras layouts (Conv2D (kh,kw,Cin,Cout), Conv2DTranspose
 *     (kh,kw,Cout,Cin), depthwise (kh,kw,C,1), Dense (in,out)).
 *   - every buffer is owned by the caller; "zeroed" means the caller cleared it before the call (the host keeps
 *     all such accumulators in one arena cleared by a single mvae_memset_zero per step).
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation, no
 *     allocation: every call is CUDA-graph capturable.
 *   - return value 0 = ok, negative = error (text via mvae_last_error); nothing throws or exits.
 */
#ifndef MVAE_B200_H
#define MVAE_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* mvae_stream_t;

enum { MVAE_OK = 0, MVAE_ERR_ARG = -1, MVAE_ERR_CUDA = -2, MVAE_ERR_UNSUPPORTED = -3 };
enum { MVAE_ACT_NONE = 0, MVAE_ACT_RELU = 1, MVAE_ACT_ELU = 2 };
enum { MVAE_DIFF_NO_UPSAMPLE = 0, MVAE_DIFF_LAPLACIAN = 1 };
enum { MVAE_PREC_FP32 = 0, MVAE_PREC_TF32 = 1 };
enum { MVAE_REG_NONE = 0, MVAE_REG_L1 = 1, MVAE_REG_L2 = 2 };

int mvae_version(void);
int mvae_last_error(char* buf, size_t n);
/* returns 10*major+minor of the current device (100 on B200), negative on error */
int mvae_device_arch(void);
int mvae_memset_zero(void* ptr, size_t bytes, mvae_stream_t stream);
/* dst[i] += alpha * src[i], i < n: running sums of the step's loss scalars (the epoch means Keras `fit` reports,
 * multiscale_vae.py:550-557) without a host round trip per batch */
int mvae_accumulate(float* dst, const float* src, int n, float alpha, mvae_stream_t stream);
/* A CUDA stream of the current device that no other owner shares (cudaStreamNonBlocking; priority 0 = the device's lowest
 * (= default), 1 = middle of its range, 2 = highest).  The step forks into ~50 streams (levels, weight-gradient side lanes); streams
 * taken from a framework's fixed pool alias one another beyond its size, and two branches on one aliased stream are
 * serialised in the captured graph (measured: the same cfg2 step 1.33 ms on distinct streams, 1.46-1.53 ms on aliased ones). */
int mvae_stream_create(int priority, mvae_stream_t* stream);
int mvae_stream_destroy(mvae_stream_t stream);
/* number of kernel launches the library has made in this process (every launch goes through one helper).  Counted around
 * the capture of a step's CUDA graphs it is the number of kernel nodes one replay executes. */
long long mvae_kernel_launch_count(void);
/* number of tcgen05 (tensor-core) kernel launches made by this process: lets callers prove the TF32 path ran */
long long mvae_tc_launch_count(void);
/* Host-side launch policy: a single mvae_conv2d_wgrad launch takes `sms` SMs (0 = the default share of 32, chosen so that the
 * dgrad chain it runs beside keeps finding free SMs).  The engine raises it to the whole GPU for the weight gradients it
 * issues at the END of a level's backward pass, when nothing of that level is left to protect.  Returns the previous value. */
int mvae_set_wgrad_sm_share(int sms);
/* debug: device buffer of 1 + 3*1000 int64 (zeroed); CTA 0 of the TMA conv kernels records (event, tile, globaltimer ns)
 * triples (scripts/trace_conv.py); NULL switches it off */
int mvae_debug_trace(long long* buf);

/* ---------------------------------------------------------------------------------------------------------
 * Pyramid.  multiscale_vae.py:129-160 + 292-315 (normalize Lambda, gaussian_filter_block, MaxPool2D(1x1,s2),
 * Subtract) and layer_blocks.py:23-101 (diff_mode LAPLACIAN: Subtract of UpSampling2D(bilinear)).
 * bands[i] : (B, H>>i, W>>i, C), i < levels.  taps: kh*kw host floats (layer_blocks.py:980-1002), kh,kw odd <= 7.
 * workspace: mvae_pyramid_split_workspace_bytes().  H and W must be divisible by 2^(levels-1).
 * --------------------------------------------------------------------------------------------------------- */
size_t mvae_pyramid_split_workspace_bytes(int B, int H, int W, int C, int levels);
int mvae_pyramid_split(const float* x, float* const* bands, void* workspace, int B, int H, int W, int C,
                       int levels, float v0, float v1, const float* taps, int kh, int kw, int diff_mode,
                       mvae_stream_t stream);
/* gaussian_filter_block (layer_blocks.py:1008-1050): frozen depthwise conv, SAME zero pad, stride 1 */
int mvae_gaussian_filter(const float* x, float* y, int B, int H, int W, int C, const float* taps, int kh, int kw,
                         mvae_stream_t stream);

/* Training-time input corruption of the "multiscale" model, multiscale_vae.py:139-147 (GaussianNoise(stddev) in normalised
 * space, then SpatialDropout2D): out = denorm( keep[b,c]*keep_scale * (norm(x) + noise_std*noise) ), raw units, unclipped.
 * noise: N(0,1) like x (nullable); keep: (B,C) of 0/1 (nullable), keep_scale = 1/(1-rate). */
int mvae_input_corrupt(const float* x, const float* noise, const float* keep, float* out, int B, int HW, int C, float v0,
                       float v1, float noise_std, float keep_scale, mvae_stream_t stream);

/* Merge.  multiscale_vae.py:204-219 (UpSampling2D(2,'bilinear') + Add, coarse to fine).
 * r0 = merged, still in [-1,1] units (the denormalize Lambda, :221-222, is fused into the consumers below). */
size_t mvae_pyramid_merge_workspace_bytes(int B, int H, int W, int C, int levels);
int mvae_pyramid_merge_fwd(const float* const* ys, float* r0, void* workspace, int B, int H, int W, int C,
                           int levels, mvae_stream_t stream);
/* adjoint: dys[0] = dr0 (copy skipped when the pointers are equal), dys[i+1] = up2^T(dys[i]) */
int mvae_pyramid_merge_bwd(const float* dr0, float* const* dys, int B, int H, int W, int C, int levels,
                           mvae_stream_t stream);
/* denormalize Lambda, multiscale_vae.py:86-94 */
int mvae_denormalize_clip(const float* r0, float* out, long long n, float v0, float v1, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * ELBO.  compile(), multiscale_vae.py:453-495.
 * sums: (B, 1+2C) zeroed: [0] = sum|y-yh|, [1..C] = sum_hw(y-yh), [1+C..2C] = same over the centre crop.
 * out (nullable) receives clip(denorm(r0)).
 * --------------------------------------------------------------------------------------------------------- */
int mvae_recon_loss_fwd(const float* r0, const float* y, float* out, float* sums, int B, int H, int W, int C,
                        float v0, float v1, mvae_stream_t stream);
/* dr0 = d(loss)/d(r0) for loss = r_scale * sum_b L_b   (r_scale = r_loss_factor / B) */
int mvae_recon_loss_bwd(const float* r0, const float* y, const float* sums, float* dr0, int B, int H, int W,
                        int C, float v0, float v1, float r_scale, mvae_stream_t stream);
/* kl: (levels, B).  per_sample: (3,B) = r_loss, r_metric (vae_r_loss, :453-456), kl.  scalars: (4) =
 * mean_b(r*rf + kl*kf), mean r_loss, mean r_metric, mean kl. */
int mvae_loss_finalize(const float* sums, const float* kl, int levels, float* per_sample, float* scalars, int B,
                       int H, int W, int C, float r_factor, float kl_factor, mvae_stream_t stream);

/* sample Lambda, multiscale_vae.py:372-383 (logvar_scale 1.0) or multiscale_vae_.py:29-34 (0.5), fused with the
 * per-scale KL of :485-488.  mulv: (B, 2z) = [mu | log_var]; eps ~ N(0,1) supplied. */
int mvae_reparam_kl_fwd(const float* mulv, const float* eps, float* z, float* kl, int B, int zdim,
                        float logvar_scale, float sample_std, mvae_stream_t stream);
/* dmulv = [dz + kl_scale*mu | dz*s*exp(s*lv)*std*eps + kl_scale*0.5*(exp(lv)-1)],  kl_scale = kl_factor / B */
int mvae_reparam_kl_bwd(const float* mulv, const float* eps, const float* dz, float* dmulv, int B, int zdim,
                        float logvar_scale, float sample_std, float kl_scale, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM (Conv2D multiscale_vae.py:333-341, layer_blocks.py:594-602,625-634,946-949;
 * Conv2DTranspose layer_blocks.py:950-951 == dgrad; Dense multiscale_vae.py:359-370,402-406 == 1x1 on H=W=1).
 * The descriptor always describes the FORWARD convolution x(B,H,W,Cin) -> y(B,ceil(H/sh),ceil(W/sw),Cout) with
 * TensorFlow 'SAME' padding.  coord_mode 2/3 appends CoordConv channels xx,yy[,rr] (coord.py:88-133) to x on
 * the fly: w then has Cin+coord_mode input channels.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    int B, H, W, Cin;
    int kh, kw, sh, sw;
    int Cout;
    int coord_mode;
    int precision;
} mvae_conv_desc;

/* y = act( conv(x * gate[b,cin]) + bias ) + residual     (gate, bias, residual nullable) */
int mvae_conv2d_fwd(const mvae_conv_desc* d, const float* x, const float* w, const float* bias,
                    const float* gate, const float* residual, int act, float* y, mvae_stream_t stream);
/* dx = ( conv^T(dy) + bias + residual ) * act'(act_out)  (bias: Conv2DTranspose forward; act_out: the forward
 * OUTPUT of the activation whose input-gradient is wanted; all nullable) */
int mvae_conv2d_dgrad(const mvae_conv_desc* d, const float* dy, const float* w, const float* bias,
                      const float* residual, const float* act_out, int act, float* dx, mvae_stream_t stream);
/* dw += im2col(x * gate)^T dy ; dbias += sum dy   (accumulating: dw/dbias zeroed by the caller; dbias nullable) */
int mvae_conv2d_wgrad(const mvae_conv_desc* d, const float* x, const float* gate, const float* dy, float* dw,
                      float* dbias, mvae_stream_t stream);

/* Dense layers -- the fused mu||logvar head of every encoder level (multiscale_vae.py:359-370) and the Dense that opens
 * every decoder level (multiscale_vae.py:402-406).  w is the Keras (K, N) kernel.
 *   fwd  : y[M,N]  = act( x[M,K] w + bias )
 *   dgrad: dx[M,K] = ( dy[M,N] w^T ) * act'(act_out)        (act_out: forward OUTPUT of the activation that produced x)
 * With MVAE_PREC_TF32 and K, N multiples of 32 they run as tensor-core skinny GEMMs (column tiles or split-K; split-K
 * partial sums use `ws`, mvae_dense_workspace_bytes(M, K, N) bytes of caller scratch, added in a fixed order).  Any other
 * shape, fp32, or a missing / short workspace takes the mvae_conv2d_* path with H = W = 1.  The weight gradient is
 * mvae_conv2d_wgrad with the same H = W = 1 descriptor. */
size_t mvae_dense_workspace_bytes(int M, int K, int N);
int mvae_dense_fwd(int M, int K, int N, const float* x, const float* w, const float* bias, int act, float* y, float* ws,
                   size_t ws_bytes, int precision, mvae_stream_t stream);
int mvae_dense_dgrad(int M, int K, int N, const float* dy, const float* w, const float* act_out, int act, float* dx,
                     float* ws, size_t ws_bytes, int precision, mvae_stream_t stream);

/* Batched variants: n independent problems of the SAME layer (the reference builds one encoder/decoder per pyramid level
 * from one config, multiscale_vae.py:172-200, so layer k of every level has the same channels / kernel / stride and differs
 * only in image size and weights).  Semantics == n single calls in order; members that fit the TMA tensor-core kernels
 * share one launch.  Array arguments have n entries; nullable arrays may be NULL as a whole. */
int mvae_conv2d_fwd_batched(int n, const mvae_conv_desc* d, const float* const* x, const float* const* w,
                            const float* const* bias, const float* const* gate, const float* const* residual, int act,
                            float* const* y, mvae_stream_t stream);
int mvae_conv2d_dgrad_batched(int n, const mvae_conv_desc* d, const float* const* dy, const float* const* w,
                              const float* const* bias, const float* const* residual, const float* const* act_out, int act,
                              float* const* dx, mvae_stream_t stream);
int mvae_conv2d_wgrad_batched(int n, const mvae_conv_desc* d, const float* const* x, const float* const* gate,
                              const float* const* dy, float* const* dw, float* const* dbias, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * mobilenetV3 block internals (layer_blocks.py:604-623) and squeeze-excite (layer_blocks.py:418-462)
 * --------------------------------------------------------------------------------------------------------- */
/* u = relu(dw3x3(a) + bias); gap_sum[b,c] += sum_hw u   (gap_sum zeroed, nullable) */
int mvae_dwconv3x3_fwd(const float* a, const float* w, const float* bias, float* u, float* gap_sum, int B, int H,
                       int W, int C, mvae_stream_t stream);
/* with d_pre = (gate*dv + dgap) * (u>0):  da = (dw3x3^T(d_pre)) * (a>0);  dw += ..., dbias += ... */
int mvae_dwconv3x3_bwd(const float* a, const float* u, const float* dv, const float* gate, const float* dgap,
                       const float* w, float* da, float* dw, float* dbias, int B, int H, int W, int C,
                       mvae_stream_t stream);
/* gate = hard_sigmoid(BN(relu(gap W0 + b0)) W1 + b1), BN with batch statistics when training (moving stats
 * updated in place).  ws: mvae_se_gate_ws_floats(B, C) floats of scratch; fwd keeps gap, h1, hn, s, mean, rstd there for
 * the bwd.  Runs as one thread-block cluster of 8 CTAs that split the batch. */
long long mvae_se_gate_ws_floats(int B, int C);
int mvae_se_gate_fwd(const float* gap_sum, const float* w0, const float* b0, const float* gamma,
                     const float* beta, const float* w1, const float* b1, float* moving_mean, float* moving_var,
                     float* gate, float* ws, int B, int C, int HW, float eps, float momentum, int training,
                     mvae_stream_t stream);
/* dg[b,c] += sum_hw dv*u   (dg zeroed).  u == NULL: plain sum over hw (GlobalAveragePooling2D numerator) */
int mvae_se_dgate_reduce(const float* dv, const float* u, float* dg, int B, int HW, int C, mvae_stream_t stream);
/* dgap[b,c] = d(loss)/d(gap_sum)  (already divided by HW); parameter gradients accumulate */
int mvae_se_gate_bwd(const float* dg, const float* w0, const float* gamma, const float* beta, const float* w1, float* ws,
                     float* dgap, float* dw0, float* db0, float* dgamma, float* dbeta, float* dw1, float* db1,
                     int B, int C, int HW, mvae_stream_t stream);

/* batched variants of the four calls above (same B and C for every member; H, W / HW per member) */
int mvae_dwconv3x3_fwd_batched(int n, const float* const* a, const float* const* w, const float* const* bias,
                               float* const* u, float* const* gap_sum, int B, const int* H, const int* W, int C,
                               mvae_stream_t stream);
int mvae_dwconv3x3_bwd_batched(int n, const float* const* a, const float* const* u, const float* const* dv,
                               const float* const* gate, const float* const* dgap, const float* const* w, float* const* da,
                               float* const* dw, float* const* dbias, int B, const int* H, const int* W, int C,
                               mvae_stream_t stream);
int mvae_se_gate_fwd_batched(int n, const float* const* gap_sum, const float* const* w0, const float* const* b0,
                             const float* const* gamma, const float* const* beta, const float* const* w1,
                             const float* const* b1, float* const* moving_mean, float* const* moving_var,
                             float* const* gate, float* const* ws, int B, int C, const int* HW, float eps, float momentum,
                             int training, mvae_stream_t stream);
int mvae_se_gate_bwd_batched(int n, const float* const* dg, const float* const* w0, const float* const* gamma,
                             const float* const* beta, const float* const* w1, float* const* ws, float* const* dgap,
                             float* const* dw0, float* const* db0, float* const* dgamma, float* const* dbeta,
                             float* const* dw1, float* const* db1, int B, int C, const int* HW, mvae_stream_t stream);
int mvae_se_dgate_reduce_batched(int n, const float* const* dv, const float* const* u, float* const* dg, int B,
                                 const int* HW, int C, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused mobilenetV3 kernels (layer_blocks.py:556-648) for Cin == filters == 32, MVAE_PREC_TF32.  A block has one
 * batch-wide dependency per direction (the BatchNorm inside the squeeze-excite gate, layer_blocks.py:447-449); everything
 * between two such points runs in ONE launch on a 256-pixel tile held in shared / tensor memory:
 *   fwd:  [y = (u_prev * gate_prev) w2 + b2 + x_prev]  then  [a = relu(y w0 + b0); u = relu(dw3x3(a) + bd); gap_sum += sum_hw u]
 *   bwd:  [dv = dy w2^T; d_pre = (dv*gate + dgap)(u>0); da = dw3x3^T(d_pre)(a>0); dwd, dbd +=; dx = da w0^T + dy]
 *         then [dgate_prev += sum_hw (dx w2_prev^T) * u_prev]
 * Either half may be absent (its weight pointer NULL): the first launch of a chain has no first half (it reads `x` /
 * the landed `dy`), the last one has no second half.  The 1x1 weight gradients are mvae_conv2d_wgrad calls on the x, u
 * (with gate), da and dy tensors.  Shapes: H*W divides 256, or W in {8..128} dividing 256 with H a multiple of 256/W;
 * mvae_mbv3_fused_supported() tells.  All tensors NHWC fp32, 16-byte aligned.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    int B, H, W, C;
    const float* u_prev; const float* x_prev; const float* gate_prev; const float* w2; const float* b2; float* y;
    const float* x; const float* w0; const float* b0; const float* wd; const float* bd;
    float* a;               /* nullable: inference does not keep it */
    float* u; float* gap_sum;
    /* Squeeze-excite gate (layer_blocks.py:418-462) folded into the launches -- whole-image tiles (H*W <= 256) only, every
     * pointer NULL otherwise.  Second half: se_w0/se_b0 given -> it writes gap mean and relu(gap W0 + b0) of its images into
     * se_ws (the scratch of mvae_se_gate_fwd, same layout) instead of adding to gap_sum.  First half: se_w1_prev given -> every
     * CTA reduces that h over the batch for the BatchNorm statistics (moving statistics updated when training) and finishes
     * the gate of its images: gate_out_prev is WRITTEN (and gate_prev ignored); se_ws_prev receives s, mean, rstd. */
    const float* se_w0; const float* se_b0; float* se_ws;
    double* se_stat;               /* 64 doubles, zeroed: sum h, sum h^2 over the batch (second half adds, first half of
                                      the NEXT launch reads them through se_stat_prev) */
    const double* se_stat_prev;
    const float* se_gamma_prev; const float* se_beta_prev; const float* se_w1_prev; const float* se_b1_prev;
    float* se_mm_prev; float* se_mv_prev; float* se_ws_prev; float* gate_out_prev;
    float bn_eps, bn_momentum; int training;
} mvae_mbv3_fwd_args;
typedef struct {
    int B, H, W, C;
    const float* dy;
    const float* u; const float* a; const float* gate; const float* dgap; const float* w2; const float* wd; const float* w0;
    float* da; float* dx; float* dwd; float* dbd;
    const float* w2_prev; const float* u_prev; float* dgate_prev;
    /* folded gate, backward.  Second half: se_w1_prev given -> dgate_prev is plainly stored (not accumulated) and
     * ds = dgate * hard_sigmoid'(s), dhn = ds W1^T of the tile's images go into se_ws_prev.  First half: se_w0 given ->
     * every CTA reduces the two BatchNorm-backward sums over the batch and finishes dgap of its images (dgap ignored).
     * The squeeze-excite WEIGHT gradients remain mvae_se_gate_bwd(dgate_prev, ..., se_ws_prev) -- off the critical path. */
    const float* se_w1_prev; float* se_ws_prev;
    double* se_bstat_prev;         /* 64 doubles, zeroed: sum dhn * xh, sum dhn over the batch (second half adds) */
    const float* se_w0; const float* se_gamma; const float* se_ws;
    const double* se_bstat;        /* the same buffer of THIS block, read by the first half */
} mvae_mbv3_bwd_args;
int mvae_mbv3_fused_supported(int B, int H, int W, int Cin, int filters);
int mvae_mbv3_fused_fwd(const mvae_mbv3_fwd_args* a, mvae_stream_t stream);
int mvae_mbv3_fused_bwd(const mvae_mbv3_bwd_args* a, mvae_stream_t stream);
/* n <= 8 problems in ONE launch (the same step of the chains of several pyramid levels: shapes, weights and halves may
 * differ per member); semantics == n single calls.  Not thread-safe (one launch-parameter block per process). */
int mvae_mbv3_fused_fwd_batched(int n, const mvae_mbv3_fwd_args* a, mvae_stream_t stream);
int mvae_mbv3_fused_bwd_batched(int n, const mvae_mbv3_bwd_args* a, mvae_stream_t stream);

/* y[b,hw,c] = x[b,hw,c] * gate[b,c]: the Multiply of squeeze_excite_block (layer_blocks.py:458-460) when the block
 * is used on its own; inside the model the scale is fused into the next conv's operand load. */
int mvae_channel_scale(const float* x, const float* gate, float* y, int B, int HW, int C, mvae_stream_t stream);
/* out[c] += sum_p x[p,c]  (bias gradient of Conv2DTranspose, layer_blocks.py:950-951; out zeroed) */
int mvae_colsum(const float* x, float* out, long long M, int C, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Decoder tail: BatchNormalization(momentum .999, eps 1e-4) + Conv2D 1x1 -> C (multiscale_vae.py:420-431)
 * stat_sums: 2*Cf doubles zeroed; stats: (2,Cf) floats = mean, rstd (written by fwd, read by bwd)
 * --------------------------------------------------------------------------------------------------------- */
int mvae_bn_stats(const float* x, double* stat_sums, long long M, int Cf, mvae_stream_t stream);
int mvae_bn_convout_fwd(const float* x, const double* stat_sums, const float* gamma, const float* beta,
                        float* moving_mean, float* moving_var, const float* w, const float* bias, float* y,
                        float* stats, long long M, int Cf, int Co, float eps, float momentum, int training,
                        mvae_stream_t stream);
/* red: (Cf*Co + Co) floats zeroed.  dx written; dgamma,dbeta,dw,dbias accumulate. */
int mvae_bn_convout_bwd(const float* x, const float* dy, const float* stats, const float* gamma,
                        const float* beta, const float* w, float* red, float* dx, float* dgamma, float* dbeta,
                        float* dw, float* dbias, long long M, int Cf, int Co, mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimiser: Keras kernel_regularizer 'l1'/'l2' (factor 0.01) + Adagrad(clipnorm) (multiscale_vae.py:497-499).
 * segs: device (nseg,5) int64 = offset, count, width, ld, reg  (element j of a segment lives at
 *       offset + (j/width)*ld + j%width);  chunks: device (nchunk,2) int64 = seg, first element.
 * norms:  grads <- grads*grad_scale + reg'(params);  sumsq[seg] = |grads|^2;  reg_loss += reg(params).  Deterministic:
 *         every chunk writes one partial (partials: 2*nchunk floats of workspace) and a second launch sums each segment's
 *         partials in a fixed order, so replicas that hold the same gradient compute the same clip factors bit for bit.
 *         The chunk table must be sorted by segment.
 * adagrad: g = grads * clip/max(|g|,clip);  acc += g^2;  params -= lr * g / (sqrt(acc) + eps)
 * --------------------------------------------------------------------------------------------------------- */
int mvae_optim_norms(const float* params, float* grads, const long long* segs, const long long* chunks, int nseg,
                     int nchunk, int chunk_elems, float grad_scale, float* partials, float* sumsq, float* reg_loss,
                     mvae_stream_t stream);
int mvae_optim_adagrad(float* params, const float* grads, float* acc, const long long* segs,
                       const long long* chunks, int nchunk, int chunk_elems, const float* sumsq,
                       const float* lr, float clip_norm, float eps, mvae_stream_t stream);

/* CoordinateChannel2D (coord.py:88-133) as a standalone layer: y (B,H,W,C+2|3) */
int mvae_coord_channels(const float* x, float* y, int B, int H, int W, int C, int use_radius,
                        mvae_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink / NVSwitch peer memory (one process per GPU).  The reference trains on
 * one device (multiscale_vae.py:550-557 `fit`); the batch shards, the weights are replicated and the one exchange is a
 * sum all-reduce of the flat gradient buffer between the backward pass and the optimiser (:497-499).
 *   set-up (once): every rank allocates a signal block, exports it and its gradient buffer as CUDA IPC handles
 *   (mvae_comm_export: handle of the ALLOCATION that holds ptr + the offset of ptr inside it; handle buffer of
 *   mvae_comm_handle_bytes() bytes), exchanges the handles through any host channel (torch.distributed here) and maps
 *   the peers' (mvae_comm_open -> mapped base for mvae_comm_close, and the pointer at the offset).
 *   per exchange: mvae_comm_allreduce -- ONE kernel, capturable: cross-GPU barrier, then rank r sums slice r of all
 *   buffers in rank order (reading peers over NVLink) and stores the sum into every rank's buffer, barrier.  bufs / signals: host arrays
 *   of `world` device pointers as seen from THIS rank (own pointers at index `rank`), buffers 16-byte aligned.  The
 *   exchange covers `nranges` (1..8) element ranges [lo[k], lo[k] + n[k]) of the buffers (floats, multiples of 4; host
 *   arrays) in one launch.  Every rank must pass the same ranges, channel and ctas (0 = default 128, at most 256).  In
 *   place: on return every rank's buffer holds the elementwise sum over ranks, bit-identical on all ranks.  Exchanges on
 *   the SAME channel (0..15) must be ordered one after the other (stream order / dependencies) on every rank; different
 *   channels may overlap in time (e.g. the big Dense gradients of a level on its side stream while the backward pass
 *   continues, the rest at the end).  The caller keeps the buffers valid and does not write the ranges while the kernel runs.
 *   mvae_comm_status: *timed_out != 0 when a barrier gave up (MVAE_COMM_TIMEOUT_MS, default 20000) -- a peer died.
 * --------------------------------------------------------------------------------------------------------- */
size_t mvae_comm_handle_bytes(void);
int mvae_comm_alloc_signals(void** signals);
int mvae_comm_free_signals(void* signals);
int mvae_comm_export(const void* ptr, void* handle, unsigned long long* offset);
int mvae_comm_open(const void* handle, unsigned long long offset, void** mapped_base, void** ptr);
int mvae_comm_close(void* mapped_base);
int mvae_comm_allreduce(float* const* bufs, void* const* signals, int rank, int world, int nranges, const long long* lo,
                        const long long* n, int channel, int ctas, mvae_stream_t stream);
int mvae_comm_status(const void* signals, int* timed_out);

#ifdef __cplusplus
}
#endif
#endif
