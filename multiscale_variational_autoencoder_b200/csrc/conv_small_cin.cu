// Direct 3x3 stride-1 SAME convolution for the few-channel first layer (`conv_base`: C -> 32 with C = 1..4 image channels
// plus 0/2/3 generated CoordConv channels; multiscale_vae.py:333-341, coord.py:88-133) and its weight gradient.
// K = 9*CinT <= 72 is far too small for the tensor-core tile pipeline and the generic implicit-GEMM kernel spends its time
// on im2col index arithmetic (94 us forward / 125 us wgrad at 256 x 32x32x3); here a CTA stages a pixel tile (+halo, CoordConv
// channels generated while staging) and the weights in shared memory once and every thread does plain FMAs.
#include "common.cuh"

namespace mvae {

struct ConvGeom {
    int B, H, W, Cin;
    int Ho, Wo, Cout;
    int kh, kw, sh, sw, pt, pl;
    int coord;
    int CinT;
};

namespace sc {

constexpr int TH = 8, TW = 32;                 // output pixels per tile
constexpr int kThreads = 256;

__device__ __forceinline__ float coord_val(int which, int iy, int ix, int H, int W) {
    const float xx = (float)iy / (float)(H - 1) * 2.f - 1.f;
    const float yy = (float)ix / (float)(W - 1) * 2.f - 1.f;
    if (which == 0) return xx;
    if (which == 1) return yy;
    return sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f));
}

// staged tile: (TH+2) x (TW+2) pixels x CT channels, zero outside the image (SAME padding pads the CoordConv channels too)
__device__ __forceinline__ void stage_input(float* S, const float* __restrict__ x, int b, int y0, int x0, int H, int W, int C,
                                            int CT) {
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2) * CT; i += kThreads) {
        const int c = i % CT, p = i / CT;
        const int xl = p % (TW + 2), yl = p / (TW + 2);
        const int iy = y0 - 1 + yl, ix = x0 - 1 + xl;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
            v = c < C ? __ldg(x + (((long long)b * H + iy) * W + ix) * C + c) : coord_val(c - C, iy, ix, H, W);
        S[i] = v;
    }
}

// y = act(conv(x) + bias).  thread = (pixel, 4-channel group); CQ = Cout/4 groups.
template <int CT>
__global__ void __launch_bounds__(kThreads) fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                                                       int C, int Cout, int act, int tiles_x) {
    pdl_sync();
    extern __shared__ __align__(16) float sm[];
    float* SW = sm;                                    // 9*CT*Cout weights
    float* SX = sm + 9 * CT * Cout;                    // input tile
    const int b = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    for (int i = threadIdx.x; i < 9 * CT * Cout; i += kThreads) SW[i] = __ldg(w + i);
    stage_input(SX, x, b, y0, x0, H, W, C, CT);
    __syncthreads();
    // thread = (4 consecutive pixels of a row, 4-channel group): a weight float4 feeds 16 FMAs, an input value 12
    const int CQ = Cout >> 2;
    for (int o = threadIdx.x; o < TH * (TW / 4) * CQ; o += kThreads) {
        const int q = o % CQ, pg = o / CQ;
        const int xl = (pg % (TW / 4)) * 4, yl = pg / (TW / 4);
        const int oy = y0 + yl, ox = x0 + xl;
        if (oy >= H || ox >= W) continue;
        const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc[4] = {b4, b4, b4, b4};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float* xin = SX + ((yl + ky) * (TW + 2) + xl) * CT + c;
                float v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = xin[j * CT];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wv = *reinterpret_cast<const float4*>(SW + ((ky * 3 + kx) * CT + c) * Cout + 4 * q);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j].x = fmaf(v[j + kx], wv.x, acc[j].x); acc[j].y = fmaf(v[j + kx], wv.y, acc[j].y);
                        acc[j].z = fmaf(v[j + kx], wv.z, acc[j].z); acc[j].w = fmaf(v[j + kx], wv.w, acc[j].w);
                    }
                }
            }
        }
        float* yo = y + (((long long)b * H + oy) * W + ox) * Cout + 4 * q;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (ox + j < W) {
                float4 r = acc[j];
                r.x = act_apply(r.x, act); r.y = act_apply(r.y, act); r.z = act_apply(r.z, act); r.w = act_apply(r.w, act);
                *reinterpret_cast<float4*>(yo + (long long)j * Cout) = r;
            }
        }
    }
}

// dW[tap][c][co] += sum_p x(p + tap)[c] * dy[p][co], dbias[co] += sum_p dy[p][co].
// thread = (tap, 8-channel group of co, pixel group): CT x 8 accumulators in registers, so one pixel costs CT scalar + two
// 16-byte shared-memory loads for CT*8 FMAs (the previous (k', 4 co) mapping was bound by 2 loads per 4 FMAs); the
// kThreads / (9 * Cout/8) pixel groups are summed through shared memory, one global atomic per weight and CTA.
template <int CT>
__global__ void __launch_bounds__(kThreads) wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ dbias, int H, int W, int C,
                                                         int Cout, int tiles_x, int tiles_per_image, int tiles_total) {
    pdl_sync();
    extern __shared__ __align__(16) float sm[];
    float* SX = sm;                                    // (TH+2)*(TW+2)*CT
    float* SD = sm + (((TH + 2) * (TW + 2) * CT + 3) & ~3);   // TH*TW*Cout, 16-byte aligned
    const int CQ = Cout >> 2, C8 = Cout >> 3;
    const int per = 9 * C8, PG = kThreads / per;
    const int tid = threadIdx.x;
    const int grp = tid / per, r = tid - grp * per;
    const bool active = grp < PG;
    const int t = r / C8, o8 = r - t * C8;
    const int ky = t / 3, kx = t - ky * 3;
    float acc[CT][8];
    float bacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        bacc[j] = 0.f;
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[c][j] = 0.f;
    }
    const bool do_bias = dbias != nullptr && t == 0;
    // persistent CTAs over the tiles of ALL images: the accumulators live in registers across tiles and every CTA issues its
    // 9*CT*Cout + Cout global atomics ONCE.  With a CTA per one or two tiles (1024 CTAs at cfg2 level 0) each of the 896
    // addresses took 1024 serialised L2 atomics -- ~18 us of a 37 us launch whatever the image size.
    for (int gt = blockIdx.x; gt < tiles_total; gt += gridDim.x) {
        const int b = gt / tiles_per_image, tile = gt - b * tiles_per_image;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int y0 = ty * TH, x0 = tx * TW;
        __syncthreads();
        // the dy tile (32 KB at 32 output channels) goes global -> shared with cp.async: all of a thread's 16-byte copies are in
        // flight together and never pass through registers (a load -> store loop here paid one global latency per copy:
        // 58 us for the 33 MB of cfg2 level 0); pixels outside the image are zero-filled (src-size 0)
        for (int i = tid; i < TH * TW * CQ; i += kThreads) {
            const int q = i % CQ, p = i / CQ;
            const int xl = p % TW, yl = p / TW;
            const int oy = y0 + yl, ox = x0 + xl;
            const bool in = oy < H && ox < W;
            const float* src = in ? dy + (((long long)b * H + oy) * W + ox) * Cout + 4 * q : dy;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(SD + 4 * i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        stage_input(SX, x, b, y0, x0, H, W, C, CT);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (active) {
            for (int p = grp; p < TH * TW; p += PG) {
                const int xl = p % TW, yl = p / TW;
                const float4* d = reinterpret_cast<const float4*>(SD + p * Cout + o8 * 8);
                const float4 d0 = d[0], d1 = d[1];
                const float* xin = SX + ((yl + ky) * (TW + 2) + xl + kx) * CT;
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const float v = xin[c];
                    acc[c][0] = fmaf(v, d0.x, acc[c][0]); acc[c][1] = fmaf(v, d0.y, acc[c][1]);
                    acc[c][2] = fmaf(v, d0.z, acc[c][2]); acc[c][3] = fmaf(v, d0.w, acc[c][3]);
                    acc[c][4] = fmaf(v, d1.x, acc[c][4]); acc[c][5] = fmaf(v, d1.y, acc[c][5]);
                    acc[c][6] = fmaf(v, d1.z, acc[c][6]); acc[c][7] = fmaf(v, d1.w, acc[c][7]);
                }
                if (do_bias) {
                    bacc[0] += d0.x; bacc[1] += d0.y; bacc[2] += d0.z; bacc[3] += d0.w;
                    bacc[4] += d1.x; bacc[5] += d1.y; bacc[6] += d1.z; bacc[7] += d1.w;
                }
            }
        }
    }
    // sum the pixel groups in shared memory (group by group: no shared-memory atomics), then one global atomic per value
    float* RED = SD;                                   // 9*CT*Cout weights + Cout biases (<= TH*TW*Cout floats)
    float* REDB = RED + 9 * CT * Cout;
    for (int gsel = 0; gsel < PG; ++gsel) {
        __syncthreads();
        if (active && grp == gsel) {
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float* d = RED + (t * CT + c) * Cout + o8 * 8 + j;
                    *d = gsel == 0 ? acc[c][j] : *d + acc[c][j];
                }
            if (do_bias)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float* d = REDB + o8 * 8 + j;
                    *d = gsel == 0 ? bacc[j] : *d + bacc[j];
                }
        }
    }
    __syncthreads();
    for (int i = tid; i < 9 * CT * Cout; i += kThreads) atomicAdd(dw + i, RED[i]);
    if (dbias) for (int i = tid; i < Cout; i += kThreads) atomicAdd(dbias + i, REDB[i]);
}

static bool shape_ok(const ConvGeom& g, const float* gate) {
    return gate == nullptr && g.kh == 3 && g.kw == 3 && g.sh == 1 && g.sw == 1 && g.CinT >= 1 && g.CinT <= 8 &&
           (g.Cout % 8) == 0 && g.Cout <= 64 && g.B <= 65535 && 9 * g.CinT * (g.Cout / 4) <= 4 * kThreads;
}

template <int CT>
static int launch_fwd(const ConvGeom& g, const float* x, const float* w, const float* bias, int act, float* y, cudaStream_t s) {
    const int tiles_x = ceil_div(g.W, TW), tiles_y = ceil_div(g.H, TH);
    const size_t smem = ((size_t)9 * CT * g.Cout + (size_t)(TH + 2) * (TW + 2) * CT) * sizeof(float);
    MVAE_CUDA(launch_pdl(fwd_kernel<CT>, dim3(tiles_x * tiles_y, g.B), dim3(kThreads), smem, s, x, w, bias, y, g.H, g.W, g.Cin,
                         g.Cout, act, tiles_x));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

template <int CT>
static int launch_wgrad(const ConvGeom& g, const float* x, const float* dy, float* dw, float* dbias, cudaStream_t s) {
    const int tiles_x = ceil_div(g.W, TW), tiles_y = ceil_div(g.H, TH);
    const int tiles_per_image = tiles_x * tiles_y;
    const long long tiles_total = (long long)tiles_per_image * g.B;
    if (tiles_total > 0x7fffffffLL) return MVAE_ERR_UNSUPPORTED;
    // MVAE_SC_WGRAD_CTAS CTAs per SM (default 2) walk the tiles; few CTAs = few atomics per address, more = more tiles in flight
    long long gx = (long long)env_int("MVAE_SC_WGRAD_CTAS", 2) * kNumSMs;
    if (gx > tiles_total) gx = tiles_total;
    if (gx < 1) gx = 1;
    const size_t smem = ((size_t)(((TH + 2) * (TW + 2) * CT + 3) & ~3) + (size_t)TH * TW * g.Cout) * sizeof(float);
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(wgrad_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    if (smem > 96 * 1024) return MVAE_ERR_UNSUPPORTED;
    MVAE_CUDA(launch_pdl(wgrad_kernel<CT>, dim3((unsigned)gx), dim3(kThreads), smem, s, x, dy, dw, dbias, g.H, g.W, g.Cin, g.Cout,
                         tiles_x, tiles_per_image, (int)tiles_total));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

}  // namespace sc

#define MVAE_SC_DISPATCH(FN, ...)                         \
    switch (g.CinT) {                                     \
        case 1: return sc::FN<1>(__VA_ARGS__);            \
        case 2: return sc::FN<2>(__VA_ARGS__);            \
        case 3: return sc::FN<3>(__VA_ARGS__);            \
        case 4: return sc::FN<4>(__VA_ARGS__);            \
        case 5: return sc::FN<5>(__VA_ARGS__);            \
        case 6: return sc::FN<6>(__VA_ARGS__);            \
        case 7: return sc::FN<7>(__VA_ARGS__);            \
        default: return sc::FN<8>(__VA_ARGS__);           \
    }

int conv_fwd_small_cin(const ConvGeom& g, const float* x, const float* w, const float* bias, const float* gate,
                       const float* residual, int act, float* y, cudaStream_t s) {
    if (!sc::shape_ok(g, gate) || residual != nullptr) return MVAE_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15))) return MVAE_ERR_UNSUPPORTED;
    MVAE_SC_DISPATCH(launch_fwd, g, x, w, bias, act, y, s)
}

int conv_wgrad_small_cin(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                         cudaStream_t s) {
    if (!sc::shape_ok(g, gate) || (reinterpret_cast<uintptr_t>(dy) & 15)) return MVAE_ERR_UNSUPPORTED;
    MVAE_SC_DISPATCH(launch_wgrad, g, x, dy, dw, dbias, s)
}

}  // namespace mvae
