// tcgen05 / TMEM tensor-core implicit-GEMM convolution (MVAE_PREC_TF32): forward, dgrad (== Conv2DTranspose forward) and wgrad.
//
// One CTA computes 128 output pixels x N channels per tile (persistent loop over tiles).  The reduction runs over
// "chunks" of 32 channels of one filter tap: a chunk of the A operand is 128 rows x 128 bytes (fp32, read as TF32 by the
// tensor core) gathered pixel by pixel with TensorFlow-SAME zero fill -- the im2col matrix is never materialised.
//
//   warps 0-3  producers : global -> registers -> shared memory in the canonical UMMA SWIZZLE_128B layout
//                          (optionally x squeeze-excite gate), fence.proxy.async, mbarrier arrive
//   warps 4-7  epilogue  : tcgen05.ld accumulator (TMEM) -> bias / activation / residual / activation-gradient -> global
//   warp  8    MMA issue : one lane issues tcgen05.mma.kind::tf32 (M=128, N, K=8) x4 per chunk, tcgen05.commit frees the
//                          stage; two accumulator stages in TMEM overlap the epilogue of tile i with the MMAs of tile i+1
//
// Operand layouts (shared memory, 128-byte rows, 16-byte chunks XOR-swizzled with the row index modulo 8):
//   A (both modes)      K-major : [128 pixels][32 reduction channels]
//   B forward           MN-major: [N/32 groups][32 reduction channels (rows)][32 output channels]  == rows of Keras W[t][ci][:]
//                       (SWIZZLE_128B_BASE32B: 32-byte units XOR-swizzled with the row index modulo 4)
//   B dgrad             K-major : [N input channels (rows)][32 reduction channels]                 == rows of Keras W[t][ci][co..]
// so the Keras kernel layout (kh,kw,Cin,Cout) is consumed as stored by both passes, with no transposed copy.
#include <stdlib.h>
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mvae {

struct ConvGeom {
    int B, H, W, Cin;
    int Ho, Wo, Cout;
    int kh, kw, sh, sw, pt, pl;
    int coord;
    int CinT;
};

long long g_tc_launches = 0;
int g_wgrad_sms = 0;               // mvae_set_wgrad_sm_share: SM share of single weight-gradient launches (0 = default)
long long* g_trace = nullptr;      // debug timeline buffer (mvae_debug_trace)

namespace tc {

constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;          // 16 KB
constexpr int kProducerThreads = 128;
constexpr int kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;

struct Params {
    ConvGeom g;
    const float* src;        // fwd: x            dgrad: dy
    const float* wt;         // Keras (kh,kw,Cin,Cout)
    const float* bias;
    const float* gate;       // fwd only: (B, Cin)
    const float* residual;
    const float* act_out;
    float* out;
    int act, gact;
    int M, N;                // rows (pixels) and output channels of this GEMM
    int cgroups, nchunks;    // 32-channel groups per tap on the reduction side; taps * cgroups
    int tiles;
};

// MODE 0 forward, MODE 1 dgrad
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment of the operand tiles (SWIZZLE_128B atoms)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int N = p.N;
    const int bbytes = N * 128;
    const int stage_bytes = kABytes + bbytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStages + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // TMEM columns: two accumulator stages of N columns, power of two >= 32
    uint32_t ncols = 32;
    while (ncols < 2u * N) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), kProducerThreads); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpilogueThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();          // everything above (barriers, TMEM) overlapped the previous kernel's tail
    const ConvGeom& g = p.g;

    if (warp < 4) {
        // ================================================ producers ==============================================
        const int r = threadIdx.x;                                  // row of the tile owned by this thread
        const uint32_t row_off = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        int it = 0;                                                 // global chunk counter -> stage / phase
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int m = tile * kTileM + r;
            int b = -1, y0 = 0, x0 = 0;
            if (m < p.M) {
                if (MODE == 0) {
                    const int ox = m % g.Wo, t = m / g.Wo;
                    b = t / g.Ho; y0 = (t % g.Ho) * g.sh - g.pt; x0 = ox * g.sw - g.pl;
                } else {
                    const int ix = m % g.W, t = m / g.W;
                    b = t / g.H; y0 = (t % g.H) + g.pt; x0 = ix + g.pl;
                }
            }
            for (int c = 0; c < p.nchunks; ++c, ++it) {
                const int tap = c / p.cgroups, cg = c - tap * p.cgroups;
                const int ky = tap / g.kw, kx = tap - ky * g.kw;
                // ---- gather this thread's A row (128 bytes) ----
                const float4* srow = nullptr;
                if (b >= 0) {
                    if (MODE == 0) {
                        const int iy = y0 + ky, ix = x0 + kx;
                        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
                            srow = reinterpret_cast<const float4*>(p.src + (((long long)b * g.H + iy) * g.W + ix) * g.Cin + cg * 32);
                    } else {
                        const int ty = y0 - ky, tx = x0 - kx;
                        if (ty >= 0 && tx >= 0 && (ty % g.sh) == 0 && (tx % g.sw) == 0) {
                            const int oy = ty / g.sh, ox = tx / g.sw;
                            if (oy < g.Ho && ox < g.Wo)
                                srow = reinterpret_cast<const float4*>(p.src + (((long long)b * g.Ho + oy) * g.Wo + ox) * g.Cout + cg * 32);
                        }
                    }
                }
                float4 av[8];
                if (srow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) av[q] = __ldg(srow + q);
                    if (MODE == 0 && p.gate) {
                        const float4* gr = reinterpret_cast<const float4*>(p.gate + (long long)b * g.Cin + cg * 32);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 gt = __ldg(gr + q);
                            av[q].x *= gt.x; av[q].y *= gt.y; av[q].z *= gt.z; av[q].w *= gt.w;
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) av[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                // ---- this thread's share of the B chunk: N*8 16-byte pieces over 128 threads ----
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)((it / kStages) & 1);
                mbar_wait(empty_bar(s), ph ^ 1u);
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + kABytes;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(sa + row_off + (((uint32_t)q ^ sw) << 4)) = tf32_rn4(av[q]);
                const int npieces = N * 8;
                for (int idx = r; idx < npieces; idx += kProducerThreads) {
                    float4 v;
                    uint32_t off;
                    if (MODE == 0) {
                        // rows = reduction channel kr (32), 16-byte piece c16 along the N output channels
                        const int per_row = N >> 2;
                        const int kr = idx / per_row, c16 = idx - kr * per_row;
                        v = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + cg * 32 + kr) * N + c16 * 4));
                        // MN-major TF32: 32-byte units swizzled with (kr & 3)
                        off = (uint32_t)(c16 >> 3) * 4096u + (uint32_t)kr * 128u +
                              (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
                    } else {
                        // rows = output column n (forward input channel), 8 pieces of 4 reduction channels (forward Cout)
                        const int n = idx >> 3, cc = idx & 7;
                        v = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + n) * g.Cout + cg * 32 + cc * 4));
                        off = (uint32_t)n * 128u + ((((uint32_t)cc) ^ ((uint32_t)n & 7u)) << 4);
                    }
                    *reinterpret_cast<float4*>(sb + off) = tf32_rn4(v);
                }
                fence_proxy_async();
                mbar_arrive(full_bar(s));
            }
        }
    } else if (warp == 8) {
        // ================================================ MMA issue ==============================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 0 ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tl) {
                const int as = tl & 1;
                const uint32_t aph = (uint32_t)((tl >> 1) & 1);
                mbar_wait(tempty_bar(as), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * N);
                for (int c = 0; c < p.nchunks; ++c, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)((it / kStages) & 1);
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // A: K-major, 8-row groups 1024 B apart, K step of 8 tf32 = 32 B inside the 128-byte swizzle atom
                        const uint64_t da = make_desc(a_addr + 32u * k, 16u, 1024u);
                        // B fwd : MN-major (BASE32B), 8 reduction rows per MMA = two 512-byte atoms (SBO), 32-channel
                        //         groups 4096 B apart (LBO)
                        // B dgrad: K-major like A
                        const uint64_t db = (MODE == 0) ? make_desc(b_addr + 1024u * k, 4096u, 512u, 1u)
                                                        : make_desc(b_addr + 32u * k, 16u, 1024u);
                        umma_tf32(d_tmem, da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));            // stage reusable once these MMAs have read it
                }
                umma_commit(tfull_bar(as));               // accumulator complete
            }
        }
    } else {
        // ================================================ epilogue ===============================================
        const int q = warp - 4;                                     // TMEM lane quadrant of this warp
        int tl = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tl) {
            const int as = tl & 1;
            const uint32_t aph = (uint32_t)((tl >> 1) & 1);
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const int m = tile * kTileM + q * 32 + lane;
            const bool ok = m < p.M;
            for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N + n0), rr);
                if (ok) {
                    const long long o = (long long)m * N + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]),
                                               __uint_as_float(rr[j + 3]));
                        if (p.bias) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        }
                        if (p.act != MVAE_ACT_NONE) {
                            v.x = act_apply(v.x, p.act); v.y = act_apply(v.y, p.act);
                            v.z = act_apply(v.z, p.act); v.w = act_apply(v.w, p.act);
                        }
                        if (p.residual) {
                            const float4 rv = __ldg(reinterpret_cast<const float4*>(p.residual + o + j));
                            v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                        }
                        if (p.act_out) {
                            const float4 ov = __ldg(reinterpret_cast<const float4*>(p.act_out + o + j));
                            v.x *= act_grad_from_out(ov.x, p.gact); v.y *= act_grad_from_out(ov.y, p.gact);
                            v.z *= act_grad_from_out(ov.z, p.gact); v.w *= act_grad_from_out(ov.w, p.gact);
                        }
                        *reinterpret_cast<float4*>(p.out + o + j) = v;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

static inline bool al16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE>
static int launch(const Params& p, cudaStream_t s) {
    const size_t smem = (size_t)kStages * (kABytes + p.N * 128) + 256 + 1024;
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MVAE_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    const int per_sm = (smem <= 110 * 1024) ? 2 : 1;
    int grid = p.tiles < kNumSMs * per_sm ? p.tiles : kNumSMs * per_sm;
    MVAE_CUDA(launch_pdl(conv_tc_kernel<MODE>, dim3(grid), dim3(kThreads), smem, s, p));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------------------
// TMA-fed variant of the same implicit GEMM (forward, and dgrad of stride-1 convolutions):
//   warp  9    TMA producer : one lane issues cp.async.bulk.tensor (4-D map over (C, W, H, B), SWIZZLE_128B, box = 32 channels x
//                             the tile's tw x th x tb pixels, traversal stride = conv stride) per (tap, 32-channel group); out-of-
//                             image pixels arrive as zeros == TensorFlow SAME padding.  The im2col matrix is never materialised and
//                             no thread computes a gather address.
//   warps 0-3  transform    : the tensor core TRUNCATES fp32 operands to TF32; these warps round the landed A tile to nearest
//                             in place (and apply the squeeze-excite gate), stage the weight chunk, fence.proxy.async, arrive
//   warp  8    MMA issue, warps 4-7 epilogue: as above
// Deep ring (5-6 stages of 16 KB + weights) with all loads asynchronous: ~100 KB in flight per SM at 2 CTAs/SM.
// ---------------------------------------------------------------------------------------------------------------------
namespace tc2 {
using namespace tc;

constexpr int kThreads2 = 320;

struct Params2 {
    ConvGeom g;
    const float* wt;
    const float* bias;
    const float* gate;
    const float* residual;
    const float* act_out;
    float* out;
    int act, gact;
    int M, N;
    int cgroups, nchunks, tiles;
    int stages;
    // tile geometry: 128 consecutive output pixels = tb images x th rows x tw pixels
    int tw, th, tb;
    int OW, OH;            // spatial dims of the GEMM's row space (forward: Ho, Wo; dgrad: H, W)
    int sw, sh;            // traversal strides of the A box (forward: conv strides; dgrad: 1)
    int pix_per_img;       // OW * OH
    long long* trace;      // debug: block 0 records (event, tile, globaltimer) triples here when non-null (mvae_debug_trace)
    int gate_ppi;          // true pixels per image (the gate is per image even when the rows are flattened)
    int flat;              // 1x1 stride-1: the rows are one long line of M pixels
    // reduction taps of this problem: weight tap index and source offset (in A pixels) of each.  Forward / stride-1 dgrad
    // list all kh*kw taps; a parity class of a strided dgrad lists the taps that hit it.
    int ntaps;
    signed char tap_id[28], tap_dy[28], tap_dx[28];
    // pixel index (in the residual / act_out / out tensors) of GEMM row (b, y, x): b*o_b + y*o_y + x*o_x + o_off
    int o_b, o_y, o_x, o_off;
    // output tensor map: dims (N, OW, OH, images), byte strides of dims 1..3, base pointer
    unsigned long long om_stride[3];
    float* om_base;
    int om_images;
};

// One launch can serve several independent problems of the same GEMM shape (the pyramid levels: same layer, different image
// size and weights): CTAs [cta_begin[l], cta_begin[l+1]) work on problem l.  The coarse levels then cost a few extra tiles
// of an existing launch instead of a launch (+ pipeline fill + drain) each.
constexpr int kMaxBatch = 6;          // kernel parameters are limited to 4 KB
struct Batch2 {
    CUtensorMap map[kMaxBatch];
    CUtensorMap omap[kMaxBatch];       // output as a 2-D (N, M) tensor, box 32 x 128, SWIZZLE_128B: the epilogue's TMA store
    Params2 p[kMaxBatch];
    int cta_begin[kMaxBatch + 1];
    int n;
};

// debug timeline: stamps go to a small shared-memory log (cheap: no global round trip on the critical path) that CTA 0 of
// problem 0 copies out at the end
constexpr int kTraceMax = 96;
struct TraceLog { unsigned int n; long long ev[kTraceMax][3]; };
__device__ __forceinline__ void trace_ev(TraceLog* tl, int ev, int tile) {
    if (tl) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned int i = atomicAdd(&tl->n, 1u);
        if (i < (unsigned)kTraceMax) { tl->ev[i][0] = ev; tl->ev[i][1] = tile; tl->ev[i][2] = (long long)t; }
    }
}

// bias + activation of one accumulator chunk (32 columns of this thread's row), in place
template <int ACT>
__device__ __forceinline__ void bias_act(uint32_t (&rr)[32], const float* bias_s) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias_s) bv = *reinterpret_cast<const float4*>(bias_s + j);
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = __uint_as_float(rr[j + e]) + b[e];
            if (ACT == MVAE_ACT_RELU) v = fmaxf(v, 0.f);
            else if (ACT == MVAE_ACT_ELU) v = v > 0.f ? v : expm1f(v);
            rr[j + e] = __float_as_uint(v);
        }
    }
}

// WRES: all weight chunks stay resident in shared memory (rounded once per CTA) and the ring holds activations only
template <int MODE, bool WRES>
__global__ void __launch_bounds__(kThreads2, WRES ? 2 : 1) conv_tma_kernel(const __grid_constant__ Batch2 bt) {
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.cta_begin[lvl + 1]) ++lvl;
    const Params2& p = bt.p[lvl];
    const CUtensorMap& mapA = bt.map[lvl];
    const CUtensorMap& mapO = bt.omap[lvl];
    const int lbid = blockIdx.x - bt.cta_begin[lvl], lgrid = bt.cta_begin[lvl + 1] - bt.cta_begin[lvl];
    __shared__ TraceLog trace_log;
    TraceLog* tlog = (p.trace && blockIdx.x == 0) ? &trace_log : nullptr;
    if (tlog && threadIdx.x == 0) trace_log.n = 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int N = p.N;
    const int bbytes = N * 128;
    const int stage_bytes = WRES ? kABytes : kABytes + bbytes;      // multiples of 1024
    const int stages = p.stages;
    uint8_t* wres = smem;                                           // WRES: nchunks * bbytes of weights in front of the ring
    if (WRES) smem += p.nchunks * bbytes;
    uint8_t* obuf = smem + stages * stage_bytes;                    // 16 KB: one 128 x 32 output chunk, SWIZZLE_128B rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(obuf + kABytes);
    // bars: raw_full[stages], tf_full[stages], empty[stages], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * stages + 4);
    float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~(uintptr_t)15);   // N floats
    const uint32_t bar0 = smem_u32(bars);
    auto raw_bar = [&](int s) { return bar0 + 8u * s; };
    auto full_bar = [&](int s) { return bar0 + 8u * (stages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * stages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (3 * stages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (3 * stages + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = 32;
    while (ncols < 2u * N) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(full_bar(s), kProducerThreads); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpilogueThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();          // everything above (barriers, TMEM) overlapped the previous kernel's tail
    const ConvGeom& g = p.g;
    if (WRES) {
        // resident weights: ALL threads stage them (rounded to TF32, UMMA layout), eight independent 16-byte loads in flight
        // per thread -- a serial load->store loop here cost 4 us (1x1) to 20 us (3x3) of every CTA's life
        const int per_chunk = N * 8, total = p.nchunks * per_chunk;
        for (int base = threadIdx.x; base < total; base += kThreads2 * 8) {
            float4 v[8];
            uint32_t off[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + j * kThreads2;
                if (i < total) {
                    const int c = i / per_chunk, idx = i - c * per_chunk;
                    const int ti = c / p.cgroups, cg = c - ti * p.cgroups;
                    const int tap = p.tap_id[ti];
                    if (MODE == 0) {
                        const int per_row = N >> 2;
                        const int kr = idx / per_row, c16 = idx - kr * per_row;
                        v[j] = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + cg * 32 + kr) * N + c16 * 4));
                        off[j] = (uint32_t)(c * bbytes) + (uint32_t)(c16 >> 3) * 4096u + (uint32_t)kr * 128u +
                                 (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
                    } else {
                        const int n = idx >> 3, cc = idx & 7;
                        v[j] = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + n) * g.Cout + cg * 32 + cc * 4));
                        off[j] = (uint32_t)(c * bbytes) + (uint32_t)n * 128u + ((((uint32_t)cc) ^ ((uint32_t)n & 7u)) << 4);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (base + j * kThreads2 < total) *reinterpret_cast<float4*>(wres + off[j]) = tf32_rn4(v[j]);
        }
        fence_proxy_async();
        __syncthreads();
    }
    if (threadIdx.x == 0) trace_ev(tlog, 0, -1);          // setup done

    if (warp == 9) {
        // ================================================ TMA producer ===========================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
            int it = 0;
            for (int tile = lbid; tile < p.tiles; tile += lgrid) {
                const int m0 = tile * kTileM;
                const int b0 = m0 / p.pix_per_img, rem = m0 - b0 * p.pix_per_img;
                const int oy0 = rem / p.OW, ox0 = rem - oy0 * p.OW;
                for (int c = 0; c < p.nchunks; ++c, ++it) {
                    const int ti = c / p.cgroups, cg = c - ti * p.cgroups;
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)((it / stages) & 1);
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    trace_ev(tlog, 1, tile * 100 + c);            // TMA of chunk c issued
                    const int cx = ox0 * p.sw + p.tap_dx[ti], cy = oy0 * p.sh + p.tap_dy[ti];
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(raw_bar(s)), "r"((uint32_t)kABytes) : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                        ::"r"(smem_u32(smem + s * stage_bytes)), "l"(reinterpret_cast<uint64_t>(&mapA)), "r"(raw_bar(s)),
                          "r"(cg * 32), "r"(cx), "r"(cy), "r"(b0) : "memory");
                }
            }
        }
    } else if (warp < 4) {
        // ================================================ transform ==============================================
        const int r = threadIdx.x;
        const uint32_t row_off = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        const int npieces = N * 8;                       // 16-byte pieces of one weight chunk
        // one weight chunk (tap, 32 reduction channels) -> shared memory in the UMMA layout, rounded to TF32
        auto stage_weights = [&](int c, uint8_t* sb) {
            const int ti = c / p.cgroups, cg = c - ti * p.cgroups;
            const int tap = p.tap_id[ti];
            // four independent 16-byte loads in flight per thread (a load -> store loop exposes one L2 latency per piece)
            for (int base = r; base < npieces; base += kProducerThreads * 4) {
                float4 v[4];
                uint32_t off[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = base + j * kProducerThreads;
                    if (idx < npieces) {
                        if (MODE == 0) {
                            const int per_row = N >> 2;
                            const int kr = idx / per_row, c16 = idx - kr * per_row;
                            v[j] = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + cg * 32 + kr) * N + c16 * 4));
                            off[j] = (uint32_t)(c16 >> 3) * 4096u + (uint32_t)kr * 128u +
                                     (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
                        } else {
                            const int n = idx >> 3, cc = idx & 7;
                            v[j] = __ldg(reinterpret_cast<const float4*>(p.wt + ((long long)tap * g.CinT + n) * g.Cout + cg * 32 + cc * 4));
                            off[j] = (uint32_t)n * 128u + ((((uint32_t)cc) ^ ((uint32_t)n & 7u)) << 4);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (base + j * kProducerThreads < npieces) *reinterpret_cast<float4*>(sb + off[j]) = tf32_rn4(v[j]);
            }
        };
        int it = 0;
        for (int tile = lbid; tile < p.tiles; tile += lgrid) {
            const int m = tile * kTileM + r;
            const float* grow = nullptr;
            if (MODE == 0 && p.gate && m < p.M) grow = p.gate + (long long)(m / p.gate_ppi) * g.Cin;
            for (int c = 0; c < p.nchunks; ++c, ++it) {
                const int cg = c % p.cgroups;
                const int s = it % stages;
                const uint32_t ph = (uint32_t)((it / stages) & 1);
                // the gate row does not depend on the tile: fetch it while the tensor copy is still in flight
                float4 gt[8];
                if (MODE == 0 && grow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) gt[q] = __ldg(reinterpret_cast<const float4*>(grow + cg * 32) + q);
                }
                mbar_wait(raw_bar(s), ph);
                if (r == 0) trace_ev(tlog, 2, tile * 100 + c);      // chunk landed (seen by transform thread 0)
                uint8_t* sa = smem + s * stage_bytes;
                // in-place round-to-nearest TF32 of this thread's row (physical 16-byte chunk q holds logical chunk q ^ sw)
                // step q touches PHYSICAL chunk q ^ (row & 7) == logical chunk q: the eight rows of a quarter-warp hit eight
                // different 16-byte bank groups (walking the physical chunks in order is an 8-way bank conflict)
                // all eight loads first, then the stores: one shared-memory latency per row instead of eight (the compiler
                // cannot move a load of chunk q+1 above the store of chunk q)
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(sa + row_off + (((uint32_t)q ^ sw) << 4));
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (MODE == 0 && grow) { v[q].x *= gt[q].x; v[q].y *= gt[q].y; v[q].z *= gt[q].z; v[q].w *= gt[q].w; }
                    *reinterpret_cast<float4*>(sa + row_off + (((uint32_t)q ^ sw) << 4)) = tf32_rn4(v[q]);
                }
                if (!WRES) stage_weights(c, sa + kABytes);
                fence_proxy_async();
                mbar_arrive(full_bar(s));
                if (r == 0) trace_ev(tlog, 7, tile * 100 + c);      // chunk transformed
            }
        }
    } else if (warp == 8) {
        // ================================================ MMA issue ==============================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 0 ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            int it = 0, tl = 0;
            for (int tile = lbid; tile < p.tiles; tile += lgrid, ++tl) {
                const int as = tl & 1;
                const uint32_t aph = (uint32_t)((tl >> 1) & 1);
                mbar_wait(tempty_bar(as), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * N);
                for (int c = 0; c < p.nchunks; ++c, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (uint32_t)((it / stages) & 1);
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    trace_ev(tlog, 8, tile * 100 + c);             // MMA warp saw the chunk
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint32_t b_addr = WRES ? smem_u32(wres + c * bbytes) : a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_desc(a_addr + 32u * k, 16u, 1024u);
                        const uint64_t db = (MODE == 0) ? make_desc(b_addr + 1024u * k, 4096u, 512u, 1u)
                                                        : make_desc(b_addr + 32u * k, 16u, 1024u);
                        umma_tf32(d_tmem, da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(as));
                trace_ev(tlog, 3, tile);                          // MMAs of the tile issued
            }
        }
    } else {
        // ================================================ epilogue ===============================================
        const int q = warp - 4;
        int tl = 0;
        if (p.bias)
            for (int i = (int)threadIdx.x - 128; i < N; i += 128) sbias[i] = __ldg(p.bias + i);      // ordered by bar.sync 2 below
        for (int tile = lbid; tile < p.tiles; tile += lgrid, ++tl) {
            const int as = tl & 1;
            const uint32_t aph = (uint32_t)((tl >> 1) & 1);
            const int row = q * 32 + lane;
            const int m = tile * kTileM + row;
            const bool ok = m < p.M;
            const uint32_t rsw = (uint32_t)(row & 7);
            long long opix = m;                                    // pixel of this row in the residual / act_out tensors
            if (ok && (p.residual || p.act_out) && !p.flat) {
                const int bb = m / p.pix_per_img, rem = m - bb * p.pix_per_img;
                const int yy = rem / p.OW, xx = rem - yy * p.OW;
                opix = (long long)bb * p.o_b + (long long)yy * p.o_y + (long long)xx * p.o_x + p.o_off;
            }
            // tile origin for the 4-D output tensor map
            const int m0 = tile * kTileM;
            const int tb0 = m0 / p.pix_per_img, trem = m0 - tb0 * p.pix_per_img;
            const int ty0 = trem / p.OW, tx0 = trem - ty0 * p.OW;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            if (threadIdx.x == 128) trace_ev(tlog, 4, tile);      // accumulator ready
            for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N + n0), rr);
                // the previous TMA store must have finished READING the staging buffer before it is rewritten
                if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const long long o = opix * N + n0;
                // bias + activation in place, the activation chosen ONCE per chunk: with the switch inside the element
                // loop the compiler if-converted the ELU branch and every element paid for an expm1f (1.5 us per tile)
                if (p.act == MVAE_ACT_RELU) bias_act<MVAE_ACT_RELU>(rr, p.bias ? sbias + n0 : nullptr);
                else if (p.act == MVAE_ACT_ELU) bias_act<MVAE_ACT_ELU>(rr, p.bias ? sbias + n0 : nullptr);
                else if (p.bias) bias_act<MVAE_ACT_NONE>(rr, sbias + n0);
                // per-row global operands, applied in place 16 columns at a time with all loads of a half in flight (as part of
                // the store loop they were eight dependent loads: the compiler cannot move a load across the staging stores)
                if (ok && p.residual) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 rv[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) rv[jj] = __ldg(reinterpret_cast<const float4*>(p.residual + o + h * 16) + jj);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = h * 16 + jj * 4;
                            rr[j] = __float_as_uint(__uint_as_float(rr[j]) + rv[jj].x);
                            rr[j + 1] = __float_as_uint(__uint_as_float(rr[j + 1]) + rv[jj].y);
                            rr[j + 2] = __float_as_uint(__uint_as_float(rr[j + 2]) + rv[jj].z);
                            rr[j + 3] = __float_as_uint(__uint_as_float(rr[j + 3]) + rv[jj].w);
                        }
                    }
                }
                if (ok && p.act_out) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 ov[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) ov[jj] = __ldg(reinterpret_cast<const float4*>(p.act_out + o + h * 16) + jj);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = h * 16 + jj * 4;
                            rr[j] = __float_as_uint(__uint_as_float(rr[j]) * act_grad_from_out(ov[jj].x, p.gact));
                            rr[j + 1] = __float_as_uint(__uint_as_float(rr[j + 1]) * act_grad_from_out(ov[jj].y, p.gact));
                            rr[j + 2] = __float_as_uint(__uint_as_float(rr[j + 2]) * act_grad_from_out(ov[jj].z, p.gact));
                            rr[j + 3] = __float_as_uint(__uint_as_float(rr[j + 3]) * act_grad_from_out(ov[jj].w, p.gact));
                        }
                    }
                }
                // row-per-thread into the swizzled staging tile (conflict-free), then ONE TMA store of full 128-byte rows (rows
                // past M are clipped): per-thread 16-byte global stores made 8x the L2 write requests
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(obuf + (uint32_t)row * 128u + ((((uint32_t)j >> 2) ^ rsw) << 4)) =
                        make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]),
                                    __uint_as_float(rr[j + 3]));
                fence_proxy_async();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (threadIdx.x == 128) {
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&mapO)), "r"(smem_u32(obuf)), "r"(n0), "r"(tx0), "r"(ty0), "r"(tb0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
            if (threadIdx.x == 128) trace_ev(tlog, 5, tile);      // tile stored
        }
    }

    if (threadIdx.x == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && tlog) {
        trace_ev(tlog, 6, -1);                                    // CTA done
        const unsigned int n = min(trace_log.n, (unsigned)kTraceMax);
        p.trace[0] = n;
        for (unsigned int i = 0; i < n; ++i) { p.trace[1 + 3 * i] = trace_log.ev[i][0]; p.trace[2 + 3 * i] = trace_log.ev[i][1]; p.trace[3 + 3 * i] = trace_log.ev[i][2]; }
    }
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// rows of the GEMM = pixels of a (Bn, OH, OW) image stack; a tile is 128 consecutive rows.  Returns false when 128 rows
// do not form a tb x th x tw box (the caller then uses the register-gather kernel above).
static bool tile_geometry(int OH, int OW, bool flat, int& tw, int& th, int& tb) {
    if (OW >= kTileM) {
        if ((OW % kTileM) && !flat) return false;       // a flat line may end in a partial tile (TMA zero-fills, the store is masked)
        tw = kTileM; th = 1; tb = 1;
        return true;
    }
    if (kTileM % OW) return false;
    tw = OW;
    const int rows = kTileM / OW;
    if (OH >= rows) { if (OH % rows) return false; th = rows; tb = 1; return true; }
    if (rows % OH) return false;
    th = OH; tb = rows / OH;
    return true;
}

// fills p.tw/th/tb, p.stages and the tensor map of ONE problem; `wres`/`smem` describe the kernel variant it needs
template <int MODE>
static int plan2(Params2& p, CUtensorMap& map, CUtensorMap& omap, const float* src, int SC, int SW, int SH, int SB, bool& wres,
                 size_t& smem, bool allow_wres = true) {
    // src: (SB, SH, SW, SC) NHWC tensor the A operand is gathered from
    if (!tile_geometry(p.OH, p.OW, p.flat != 0, p.tw, p.th, p.tb)) return MVAE_ERR_UNSUPPORTED;
    if (p.tw * p.sw > 256 || p.th * p.sh > 256 || p.tb > 256) return MVAE_ERR_UNSUPPORTED;
    {
        // output rows as a (N, OW, OH, images) tensor: the same tb x th x tw box as the A tile, SWIZZLE_128B staging
        const unsigned long long odims[4] = {(unsigned long long)p.N, (unsigned long long)p.OW, (unsigned long long)p.OH,
                                             (unsigned long long)p.om_images};
        const unsigned int obox[4] = {32u, (unsigned)p.tw, (unsigned)p.th, (unsigned)p.tb};
        if (!tma::encode_f32_strided(&omap, p.om_base, 4, odims, p.om_stride, obox, CU_TENSOR_MAP_SWIZZLE_128B))
            return MVAE_ERR_UNSUPPORTED;
    }
    const unsigned long long dims[4] = {(unsigned long long)SC, (unsigned long long)SW, (unsigned long long)SH, (unsigned long long)SB};
    const unsigned int box[4] = {32u, (unsigned)(p.tw * p.sw), (unsigned)(p.th * p.sh), (unsigned)p.tb};
    const unsigned int es[4] = {1u, (unsigned)p.sw, (unsigned)p.sh, 1u};
    if (!tma::encode_f32(&map, src, 4, dims, box, es, CU_TENSOR_MAP_SWIZZLE_128B)) return MVAE_ERR_UNSUPPORTED;
    const int bbytes = p.N * 128;
    const int wres_bytes = p.nchunks * bbytes;
    // resident weights: two CTAs per SM when weights + ring + staging fit ~108 KB, else one CTA per SM up to ~150 KB of
    // weights; beyond that the weight chunk of every stage is streamed with the activations
    wres = allow_wres && wres_bytes <= 150 * 1024;
    if (wres) {
        const int budget = wres_bytes <= 40 * 1024 ? 108 * 1024 : 216 * 1024;
        int stages = (budget - wres_bytes - kABytes) / kABytes;
        if (stages > env_int("MVAE_CONV_STAGES", 6)) stages = env_int("MVAE_CONV_STAGES", 6);
        p.stages = stages;
        smem = (size_t)wres_bytes + (size_t)(stages + 1) * kABytes + (3 * stages + 4) * 8 + 64 + 528 + 1024;
        if (stages < 3) wres = false;
    }
    if (!wres) {
        const int stage_bytes = kABytes + bbytes;
        int stages = (200 * 1024 - kABytes) / stage_bytes;
        if (stages > 6) stages = 6;
        if (stages < 2) return MVAE_ERR_UNSUPPORTED;
        p.stages = stages;
        smem = (size_t)stages * stage_bytes + kABytes + (3 * stages + 4) * 8 + 64 + 528 + 1024;
    }
    return MVAE_OK;
}

// all problems of a batch must need the same kernel variant, shared-memory size and pipeline depth
template <int MODE>
static int launch_batch(Batch2& bt, bool wres, size_t smem, cudaStream_t s) {
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(conv_tma_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        MVAE_CUDA(cudaFuncSetAttribute(conv_tma_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    }
    // CTAs in proportion to the tile counts, at least one per problem, never more than a problem has tiles
    int budget = kNumSMs * (smem <= 110 * 1024 ? 2 : 1);
    if (env_int("MVAE_CONV_BUDGET", 0) > 0) budget = env_int("MVAE_CONV_BUDGET", 0);
    long long total_tiles = 0;
    for (int l = 0; l < bt.n; ++l) total_tiles += bt.p[l].tiles;
    bt.cta_begin[0] = 0;
    for (int l = 0; l < bt.n; ++l) {
        long long q = (long long)budget * bt.p[l].tiles / (total_tiles > 0 ? total_tiles : 1);
        if (q < 1) q = 1;
        if (q > bt.p[l].tiles) q = bt.p[l].tiles;
        bt.cta_begin[l + 1] = bt.cta_begin[l] + (int)q;
    }
    const int grid = bt.cta_begin[bt.n];
    if (wres) MVAE_CUDA(launch_pdl(conv_tma_kernel<MODE, true>, dim3(grid), dim3(kThreads2), smem, s, bt));
    else      MVAE_CUDA(launch_pdl(conv_tma_kernel<MODE, false>, dim3(grid), dim3(kThreads2), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

template <int MODE>
static int launch2(Params2& p, const float* src, int SC, int SW, int SH, int SB, cudaStream_t s) {
    Batch2 bt;
    bool wres;
    size_t smem;
    if (int e = plan2<MODE>(p, bt.map[0], bt.omap[0], src, SC, SW, SH, SB, wres, smem)) return e;
    bt.p[0] = p;
    bt.n = 1;
    return launch_batch<MODE>(bt, wres, smem, s);
}

struct SrcDims { int C, W, H, B; };

// contiguous output: GEMM row m is pixel m
static void dense_output(Params2& q, float* out) {
    q.o_b = q.pix_per_img; q.o_y = q.OW; q.o_x = 1; q.o_off = 0;
    q.om_base = out;
    q.om_images = q.flat ? 1 : ceil_div(q.M, q.pix_per_img);
    q.om_stride[0] = (unsigned long long)q.N * 4;
    q.om_stride[1] = (unsigned long long)q.OW * q.N * 4;
    q.om_stride[2] = (unsigned long long)q.pix_per_img * q.N * 4;
}

static void fwd_params(const ConvGeom& g, const float* w, const float* bias, const float* gate, const float* residual, int act,
                       float* y, Params2& q, SrcDims& sd) {
    const int M = g.B * g.Ho * g.Wo;
    q.g = g; q.wt = w; q.bias = bias; q.gate = gate; q.residual = residual; q.act_out = nullptr; q.out = y;
    q.act = act; q.gact = 0; q.M = M; q.N = g.Cout; q.cgroups = g.Cin / 32; q.nchunks = g.kh * g.kw * q.cgroups;
    q.trace = g_trace;
    q.tiles = ceil_div(M, kTileM);
    q.sw = g.sw; q.sh = g.sh; q.gate_ppi = g.Ho * g.Wo;
    q.ntaps = g.kh * g.kw;
    for (int t = 0; t < q.ntaps; ++t) { q.tap_id[t] = (signed char)t; q.tap_dy[t] = (signed char)(t / g.kw - g.pt); q.tap_dx[t] = (signed char)(t % g.kw - g.pl); }
    if (g.kh == 1 && g.kw == 1 && g.sh == 1 && g.sw == 1) {
        q.flat = 1; q.OW = M; q.OH = 1; q.pix_per_img = M;                        // plain GEMM rows
        sd = SrcDims{g.Cin, M, 1, 1};
    } else {
        q.flat = 0; q.OW = g.Wo; q.OH = g.Ho; q.pix_per_img = g.Wo * g.Ho;
        sd = SrcDims{g.Cin, g.W, g.H, g.B};
    }
    dense_output(q, y);
}

// stride-1 convolutions only
static void dgrad_params(const ConvGeom& g, const float* w, const float* bias, const float* residual, const float* act_out,
                         int act, float* dx, Params2& q, SrcDims& sd) {
    const int M = g.B * g.H * g.W;
    q.g = g; q.wt = w; q.bias = bias; q.gate = nullptr; q.residual = residual; q.act_out = act_out; q.out = dx;
    q.act = 0; q.gact = act; q.M = M; q.N = g.Cin; q.cgroups = g.Cout / 32; q.nchunks = g.kh * g.kw * q.cgroups;
    q.trace = g_trace;
    q.tiles = ceil_div(M, kTileM);
    q.sw = 1; q.sh = 1; q.gate_ppi = g.H * g.W;
    q.ntaps = g.kh * g.kw;
    for (int t = 0; t < q.ntaps; ++t) { q.tap_id[t] = (signed char)t; q.tap_dy[t] = (signed char)(g.pt - t / g.kw); q.tap_dx[t] = (signed char)(g.pl - t % g.kw); }
    if (g.kh == 1 && g.kw == 1) { q.flat = 1; q.OW = M; q.OH = 1; q.pix_per_img = M; sd = SrcDims{g.Cout, M, 1, 1}; }
    else { q.flat = 0; q.OW = g.W; q.OH = g.H; q.pix_per_img = g.W * g.H; sd = SrcDims{g.Cout, g.Wo, g.Ho, g.B}; }
    dense_output(q, dx);
}

// One parity class (py, px) of the dgrad of a strided convolution: the input pixels (sh*Y + py, sw*X + px) receive exactly
// the taps with ky = py + pt (mod sh), kx = px + pl (mod sw), read dy at (Y + (py + pt - ky)/sh, X + (px + pl - kx)/sw):
// a stride-1 convolution over dy with that tap subset whose output lands on every sh-th / sw-th pixel of dx.  No MMA is
// spent on the structural zeros of the transposed convolution.  Returns false when the class has no tap.
static bool dgrad_class_params(const ConvGeom& g, int py, int px, const float* w, const float* bias, const float* residual,
                               const float* act_out, int act, float* dx, Params2& q, SrcDims& sd) {
    const int Hc = g.H / g.sh, Wc = g.W / g.sw;
    const int M = g.B * Hc * Wc;
    q.g = g; q.wt = w; q.bias = bias; q.gate = nullptr; q.residual = residual; q.act_out = act_out; q.out = dx;
    q.act = 0; q.gact = act; q.M = M; q.N = g.Cin; q.cgroups = g.Cout / 32;
    q.trace = g_trace;
    q.ntaps = 0;
    for (int ky = 0; ky < g.kh; ++ky) {
        if ((py + g.pt - ky) % g.sh) continue;
        for (int kx = 0; kx < g.kw; ++kx) {
            if ((px + g.pl - kx) % g.sw) continue;
            q.tap_id[q.ntaps] = (signed char)(ky * g.kw + kx);
            q.tap_dy[q.ntaps] = (signed char)((py + g.pt - ky) / g.sh);
            q.tap_dx[q.ntaps] = (signed char)((px + g.pl - kx) / g.sw);
            ++q.ntaps;
        }
    }
    if (q.ntaps == 0) return false;
    q.nchunks = q.ntaps * q.cgroups;
    q.tiles = ceil_div(M, kTileM);
    q.sw = 1; q.sh = 1; q.gate_ppi = Hc * Wc;
    q.flat = 0; q.OW = Wc; q.OH = Hc; q.pix_per_img = Wc * Hc;
    sd = SrcDims{g.Cout, g.Wo, g.Ho, g.B};
    q.o_b = g.H * g.W; q.o_y = g.sh * g.W; q.o_x = g.sw; q.o_off = py * g.W + px;
    q.om_base = dx + (long long)q.o_off * q.N;
    q.om_images = g.B;
    q.om_stride[0] = (unsigned long long)g.sw * q.N * 4;
    q.om_stride[1] = (unsigned long long)g.sh * g.W * q.N * 4;
    q.om_stride[2] = (unsigned long long)g.H * g.W * q.N * 4;
    return true;
}

}  // namespace tc2

// A shape is taken by the tensor-core path when both channel counts are multiples of 32 (128-byte rows), N <= 128,
// there are no CoordConv channels, and there are enough rows to fill at least a few tiles.
static bool tc_shape_ok(const ConvGeom& g, int M, int N, int red_channels) {
    return g.coord == 0 && (red_channels % 32) == 0 && (N % 32) == 0 && N <= 128 && M >= 512;
}

int conv_fwd_tc(const ConvGeom& g, const float* x, const float* w, const float* bias, const float* gate,
                const float* residual, int act, float* y, cudaStream_t s) {
    const int M = g.B * g.Ho * g.Wo, N = g.Cout;
    if (!tc_shape_ok(g, M, N, g.Cin)) return MVAE_ERR_UNSUPPORTED;
    if (!(tc::al16(x) && tc::al16(w) && tc::al16(bias) && tc::al16(gate) && tc::al16(residual) && tc::al16(y)))
        return MVAE_ERR_UNSUPPORTED;
    tc::Params p;
    p.g = g; p.src = x; p.wt = w; p.bias = bias; p.gate = gate; p.residual = residual; p.act_out = nullptr; p.out = y;
    p.act = act; p.gact = 0; p.M = M; p.N = N; p.cgroups = g.Cin / 32; p.nchunks = g.kh * g.kw * p.cgroups;
    p.tiles = ceil_div(M, tc::kTileM);
    {
        tc2::Params2 q;
        tc2::SrcDims sd;
        tc2::fwd_params(g, w, bias, gate, residual, act, y, q, sd);
        const int r = tc2::launch2<0>(q, x, sd.C, sd.W, sd.H, sd.B, s);
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    return tc::launch<0>(p, s);
}

int conv_dgrad_tc(const ConvGeom& g, const float* dy, const float* w, const float* bias, const float* residual,
                  const float* act_out, int act, float* dx, cudaStream_t s) {
    const int M = g.B * g.H * g.W, N = g.Cin;
    if (!tc_shape_ok(g, M, N, g.Cout)) return MVAE_ERR_UNSUPPORTED;
    if (!(tc::al16(dy) && tc::al16(w) && tc::al16(bias) && tc::al16(residual) && tc::al16(act_out) && tc::al16(dx)))
        return MVAE_ERR_UNSUPPORTED;
    tc::Params p;
    p.g = g; p.src = dy; p.wt = w; p.bias = bias; p.gate = nullptr; p.residual = residual; p.act_out = act_out; p.out = dx;
    p.act = 0; p.gact = act; p.M = M; p.N = N; p.cgroups = g.Cout / 32; p.nchunks = g.kh * g.kw * p.cgroups;
    p.tiles = ceil_div(M, tc::kTileM);
    if (g.sh == 1 && g.sw == 1) {
        tc2::Params2 q;
        tc2::SrcDims sd;
        tc2::dgrad_params(g, w, bias, residual, act_out, act, dx, q, sd);
        const int r = tc2::launch2<1>(q, dy, sd.C, sd.W, sd.H, sd.B, s);
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    if ((g.sh > 1 || g.sw > 1) && g.sh * g.sw <= tc2::kMaxBatch && (g.H % g.sh) == 0 && (g.W % g.sw) == 0 &&
        g.Ho == g.H / g.sh && g.Wo == g.W / g.sw) {
        // strided: one sub-convolution per output parity class, all classes in one launch
        tc2::Batch2 bt;
        bool ok = true, wres0 = true;
        size_t smem0 = 0;
        int n = 0;
        for (int pass = 0; pass < 2; ++pass) {            // pass 1 repeats the planning with streamed weights for every class
            ok = true; n = 0; smem0 = 0;
            bool all_wres = true;
            for (int py = 0; py < g.sh && ok; ++py)
                for (int px = 0; px < g.sw && ok; ++px) {
                    tc2::SrcDims sd;
                    ok = tc2::dgrad_class_params(g, py, px, w, bias, residual, act_out, act, dx, bt.p[n], sd);
                    bool wres = false; size_t smem = 0;
                    if (ok) ok = tc2::plan2<1>(bt.p[n], bt.map[n], bt.omap[n], dy, sd.C, sd.W, sd.H, sd.B, wres, smem, pass == 0) == MVAE_OK;
                    if (ok) { all_wres = all_wres && wres; if (smem > smem0) smem0 = smem; ++n; }
                }
            wres0 = pass == 0;
            if (!ok || all_wres || pass == 1) break;
        }
        if (ok) {
            // classes have different tap counts: every member keeps its own ring depth, the launch takes the largest footprint
            bt.n = n;
            return tc2::launch_batch<1>(bt, wres0, smem0, s);
        }
    }
    return tc::launch<1>(p, s);
}


// ---- several problems of one layer shape (the pyramid levels) in one launch; MVAE_ERR_UNSUPPORTED -> caller loops ----
static bool tc_member_ok(const ConvGeom& g, int N, int red_channels) {
    return g.coord == 0 && (red_channels % 32) == 0 && (N % 32) == 0 && N <= 128;
}

int conv_fwd_tc_batched(int n, const ConvGeom* g, const float* const* x, const float* const* w, const float* const* bias,
                        const float* const* gate, const float* const* residual, int act, float* const* y, cudaStream_t s) {
    if (n < 1 || n > tc2::kMaxBatch) return MVAE_ERR_UNSUPPORTED;
    tc2::Batch2 bt;
    bool wres0 = false;
    size_t smem0 = 0;
    for (int l = 0; l < n; ++l) {
        if (!tc_member_ok(g[l], g[l].Cout, g[l].Cin)) return MVAE_ERR_UNSUPPORTED;
        const float* bi = bias ? bias[l] : nullptr; const float* ga = gate ? gate[l] : nullptr;
        const float* re = residual ? residual[l] : nullptr;
        if (!(tc::al16(x[l]) && tc::al16(w[l]) && tc::al16(bi) && tc::al16(ga) && tc::al16(re) && tc::al16(y[l]))) return MVAE_ERR_UNSUPPORTED;
        tc2::SrcDims sd;
        tc2::fwd_params(g[l], w[l], bi, ga, re, act, y[l], bt.p[l], sd);
        bool wres; size_t smem;
        if (int e = tc2::plan2<0>(bt.p[l], bt.map[l], bt.omap[l], x[l], sd.C, sd.W, sd.H, sd.B, wres, smem)) return e;
        if (l == 0) { wres0 = wres; smem0 = smem; }
        else if (wres != wres0 || smem != smem0 || bt.p[l].N != bt.p[0].N || bt.p[l].nchunks != bt.p[0].nchunks ||
                 bt.p[l].stages != bt.p[0].stages)
            return MVAE_ERR_UNSUPPORTED;
    }
    bt.n = n;
    return tc2::launch_batch<0>(bt, wres0, smem0, s);
}

int conv_dgrad_tc_batched(int n, const ConvGeom* g, const float* const* dy, const float* const* w, const float* const* bias,
                          const float* const* residual, const float* const* act_out, int act, float* const* dx,
                          cudaStream_t s) {
    if (n < 1 || n > tc2::kMaxBatch) return MVAE_ERR_UNSUPPORTED;
    tc2::Batch2 bt;
    bool wres0 = false;
    size_t smem0 = 0;
    for (int l = 0; l < n; ++l) {
        if (!tc_member_ok(g[l], g[l].Cin, g[l].Cout) || g[l].sh != 1 || g[l].sw != 1) return MVAE_ERR_UNSUPPORTED;
        const float* bi = bias ? bias[l] : nullptr; const float* re = residual ? residual[l] : nullptr;
        const float* ao = act_out ? act_out[l] : nullptr;
        if (!(tc::al16(dy[l]) && tc::al16(w[l]) && tc::al16(bi) && tc::al16(re) && tc::al16(ao) && tc::al16(dx[l]))) return MVAE_ERR_UNSUPPORTED;
        tc2::SrcDims sd;
        tc2::dgrad_params(g[l], w[l], bi, re, ao, act, dx[l], bt.p[l], sd);
        bool wres; size_t smem;
        if (int e = tc2::plan2<1>(bt.p[l], bt.map[l], bt.omap[l], dy[l], sd.C, sd.W, sd.H, sd.B, wres, smem)) return e;
        if (l == 0) { wres0 = wres; smem0 = smem; }
        else if (wres != wres0 || smem != smem0 || bt.p[l].N != bt.p[0].N || bt.p[l].nchunks != bt.p[0].nchunks ||
                 bt.p[l].stages != bt.p[0].stages)
            return MVAE_ERR_UNSUPPORTED;
    }
    bt.n = n;
    return tc2::launch_batch<1>(bt, wres0, smem0, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad on tcgen05:  dW[k' = (tap, ci)][co] += sum_p X_tap[p][ci] * dY[p][co]      (reduction over pixels p)
//
// Both operands are MN-major TF32 (SWIZZLE_128B_BASE32B): a "slab" is [32 pixels (reduction rows)][32 channels = 128 B],
// exactly a run of NHWC pixel rows, so x and dy are consumed as stored.  A CTA owns `ngroups` slabs of A (each slab = one
// (tap, 32-channel group) = 32 rows of dW; 4 slabs form one M = 128 MMA tile), all N/32 slabs of B, and a range of
// pixels; it accumulates in TMEM over its pixel range and adds the result to dW with red.global.add.f32.
// grid = (pixel splits, M splits).
// ---------------------------------------------------------------------------------------------------------------------
namespace tcw {
using namespace tc;

constexpr int kPix = 32;              // pixels (reduction rows) per stage
constexpr int kSlab = kPix * 128;     // 4 KB

struct Params {
    ConvGeom g;
    const float* x;
    const float* gate;
    const float* dy;
    float* dw;
    float* dbias;
    int P, N;                 // pixels of the forward output, Cout
    int cgroups, groups;      // Cin/32, taps*cgroups
    int groups_per_cta;       // <= 16
    int a_slabs;              // groups_per_cta rounded up to a multiple of 4 (an M = 128 MMA always reads 4 slabs)
    int pix_per_cta;          // multiple of kPix
    int stages;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const ConvGeom& g = p.g;
    const int N = p.N, nb = N >> 5;
    const int g0 = blockIdx.y * p.groups_per_cta;
    const int ng = min(p.groups_per_cta, p.groups - g0);
    const int mtiles = (ng + 3) >> 2;
    const int stage_bytes = (p.a_slabs + nb) * kSlab;
    const int stages = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    float* bias_red = reinterpret_cast<float*>(bars + 2 * stages + 2);        // N floats
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bias_red + N);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (stages + s); };
    const uint32_t tfull_bar = bar0 + 8u * (2 * stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(mtiles * N)) ncols <<= 1;

    const int pbeg = blockIdx.x * p.pix_per_cta;
    const int pend = min(p.P, pbeg + p.pix_per_cta);
    const int nchunks = (pend - pbeg + kPix - 1) / kPix;       // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), kProducerThreads); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < N; i += kThreads) bias_red[i] = 0.f;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();          // everything above (barriers, TMEM) overlapped the previous kernel's tail

    if (warp < 4) {
        // ---------------- producers: thread = (pixel px of the chunk, 32-byte unit of the 128-byte row) ----------------
        const int px = threadIdx.x >> 2, unit = threadIdx.x & 3;
        const uint32_t dst_row = (uint32_t)px * 128u + ((uint32_t)(unit ^ (px & 3)) << 5);
        const bool do_bias = p.dbias != nullptr && blockIdx.y == 0;
        float bsum[8][8];                                     // up to N = 256: nb <= 8
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[i][j] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int pp = pbeg + c * kPix + px;
            const bool pv = pp < pend;
            int b = 0, y0 = 0, x0 = 0;
            if (pv) {
                const int ox = pp % g.Wo, t = pp / g.Wo;
                b = t / g.Ho; y0 = (t % g.Ho) * g.sh - g.pt; x0 = ox * g.sw - g.pl;
            }
            const int s = c % stages;
            const uint32_t ph = (uint32_t)((c / stages) & 1);
            mbar_wait(empty_bar(s), ph ^ 1u);
            uint8_t* sa = smem + s * stage_bytes;
            uint8_t* sb = sa + p.a_slabs * kSlab;
            // A slabs
            for (int gi = 0; gi < ng; ++gi) {
                const int grp = g0 + gi;
                const int tap = grp / p.cgroups, cg = grp - tap * p.cgroups;
                const int ky = tap / g.kw, kx = tap - ky * g.kw;
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (pv) {
                    const int iy = y0 + ky, ix = x0 + kx;
                    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                        const float4* src = reinterpret_cast<const float4*>(
                            p.x + (((long long)b * g.H + iy) * g.W + ix) * g.Cin + cg * 32 + unit * 8);
                        v0 = __ldg(src); v1 = __ldg(src + 1);
                        if (p.gate) {
                            const float4* gr = reinterpret_cast<const float4*>(p.gate + (long long)b * g.Cin + cg * 32 + unit * 8);
                            const float4 g0v = __ldg(gr), g1v = __ldg(gr + 1);
                            v0.x *= g0v.x; v0.y *= g0v.y; v0.z *= g0v.z; v0.w *= g0v.w;
                            v1.x *= g1v.x; v1.y *= g1v.y; v1.z *= g1v.z; v1.w *= g1v.w;
                        }
                    }
                }
                float4* dst = reinterpret_cast<float4*>(sa + gi * kSlab + dst_row);
                dst[0] = tf32_rn4(v0); dst[1] = tf32_rn4(v1);
            }
            // B slabs (dy)
#pragma unroll
            for (int bi = 0; bi < 8; ++bi) {
                if (bi < nb) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (pv) {
                        const float4* src = reinterpret_cast<const float4*>(p.dy + (long long)pp * N + bi * 32 + unit * 8);
                        v0 = __ldg(src); v1 = __ldg(src + 1);
                    }
                    if (do_bias) {
                        bsum[bi][0] += v0.x; bsum[bi][1] += v0.y; bsum[bi][2] += v0.z; bsum[bi][3] += v0.w;
                        bsum[bi][4] += v1.x; bsum[bi][5] += v1.y; bsum[bi][6] += v1.z; bsum[bi][7] += v1.w;
                    }
                    float4* dst = reinterpret_cast<float4*>(sb + bi * kSlab + dst_row);
                    dst[0] = tf32_rn4(v0); dst[1] = tf32_rn4(v1);
                }
            }
            fence_proxy_async();
            mbar_arrive(full_bar(s));
        }
        if (do_bias) {
#pragma unroll
            for (int bi = 0; bi < 8; ++bi)
                if (bi < nb)
#pragma unroll
                    for (int j = 0; j < 8; ++j) atomicAdd(bias_red + bi * 32 + unit * 8 + j, bsum[bi][j]);
            asm volatile("bar.sync 1, 128;" ::: "memory");               // producers only
            for (int i = threadIdx.x; i < N; i += kProducerThreads) atomicAdd(p.dbias + i, bias_red[i]);
        }
    } else if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % stages;
                const uint32_t ph = (uint32_t)((c / stages) & 1);
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                const uint32_t b_addr = a_addr + p.a_slabs * kSlab;
                for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
                    for (int k = 0; k < kPix / 8; ++k) {
                        // 8 reduction rows (pixels) per MMA = two 4-row atoms (SBO 512 B); 32-channel groups one slab apart
                        const uint64_t da = make_desc(a_addr + mt * 4 * kSlab + 1024u * k, kSlab, 512u, 1u);
                        const uint64_t db = make_desc(b_addr + 1024u * k, kSlab, 512u, 1u);
                        umma_tf32(tmem_base + (uint32_t)(mt * N), da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(tfull_bar);
        }
    } else {
        // ---------------- epilogue: TMEM -> red.global.add ----------------
        const int q = warp - 4;
        mbar_wait(tfull_bar, 0u);
        tc_fence_after();
        for (int mt = 0; mt < mtiles; ++mt) {
            const int gi = mt * 4 + q;                       // slab of this warp's 32 lanes
            const bool ok = gi < ng;
            const long long row = (long long)(g0 + gi) * 32 + lane;        // row of dW: (tap*Cin + ci)
            for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N + n0), rr);
                if (ok) {
                    float* dst = p.dw + row * N + n0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(rr[j]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace tcw

// ---------------------------------------------------------------------------------------------------------------------
// TMA-fed wgrad.  Same GEMM as above (both operands MN-major TF32 slabs of [PIX pixels][32 channels]) but
//   warp 9     : one lane issues the TMA loads of a stage: `ng` slabs of x (4-D conv-geometry boxes, one per (tap, 32-channel
//                group), zero fill == SAME padding) + N/32 slabs of dy; SWIZZLE_NONE, so a slab lands as linear 128-byte rows
//   warps 0-3  : round to nearest TF32 (+ gate, + bias column sums) IN PLACE and permute each row's four 32-byte units into
//                the SWIZZLE_128B_BASE32B order the tensor core wants (unit ^= pixel & 3; the four lanes that own a row
//                read, __syncwarp, write)
//   warp 8 MMA, warps 4-7 epilogue (TMEM -> red.global.add) as above.
// PIX = 128 pixels per stage for few slabs (1x1 convs: 32 KB stages, 2 CTAs/SM), 32 for many (3x3).
// An M = 128 MMA reads 4 slabs; slabs that do not exist alias whatever follows in shared memory (their rows of D are never
// read back), so nothing is staged for them.
// ---------------------------------------------------------------------------------------------------------------------
namespace tcw2 {
using namespace tc;

struct Params {
    ConvGeom g;
    const float* gate;
    float* dw;
    float* dbias;
    int P, N;
    int cgroups, groups, groups_per_cta;
    int pix_per_cta;          // multiple of PIX
    int stages;
    int tw, th, tb;           // a stage's PIX pixels = tb images x th rows x tw pixels of the forward OUTPUT
    int flat;                 // 1x1 stride 1: pixels are one flat line
    long long* trace;         // debug timeline (mvae_debug_trace)
};

constexpr int kMaxBatch = 8;
struct BatchW {
    CUtensorMap mx[kMaxBatch], mdy[kMaxBatch];
    Params p[kMaxBatch];
    int cta_begin[kMaxBatch + 1];      // along grid.x (pixel splits); grid.y = slab-group splits, the same for every problem
    int n;
};

// NBMAX: upper bound of N/32 (sizes the per-thread bias accumulators; 1 keeps the kernel at two CTAs per SM)
template <int PIX, int NBMAX>
__global__ void __launch_bounds__(tc2::kThreads2, NBMAX == 1 ? 2 : 1) wgrad_tma_kernel(const __grid_constant__ BatchW bt) {
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.cta_begin[lvl + 1]) ++lvl;
    const Params& p = bt.p[lvl];
    const CUtensorMap& mapX = bt.mx[lvl];
    const CUtensorMap& mapDY = bt.mdy[lvl];
    const int lbid = blockIdx.x - bt.cta_begin[lvl];
    __shared__ tc2::TraceLog trace_log;
    tc2::TraceLog* tlog = (p.trace && blockIdx.x == 0 && blockIdx.y == 0) ? &trace_log : nullptr;
    if (tlog && threadIdx.x == 0) trace_log.n = 0;
    constexpr int kSlabB = PIX * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const ConvGeom& g = p.g;
    const int N = p.N, nb = N >> 5;
    const int g0 = blockIdx.y * p.groups_per_cta;
    const int ng = min(p.groups_per_cta, p.groups - g0);
    const int mtiles = (ng + 3) >> 2;
    const int stage_bytes = (ng + nb) * kSlabB;
    const int stages = p.stages;
    // [stages x stage][slack for the aliased reads of missing slabs][barriers][bias_red][tmem slot]
    // a single slab (1x1 convs) is read four times instead (LBO = 0: rows 32..127 of D repeat rows 0..31), no slack needed
    const int slack = (p.groups_per_cta == 1) ? 0 : 3 * kSlabB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes + slack);
    float* bias_red = reinterpret_cast<float*>(bars + 3 * stages + 2);        // [4 transform warps][N]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bias_red + 4 * N);
    const uint32_t bar0 = smem_u32(bars);
    auto raw_bar = [&](int s) { return bar0 + 8u * s; };
    auto full_bar = [&](int s) { return bar0 + 8u * (stages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * stages + s); };
    const uint32_t tfull_bar = bar0 + 8u * (3 * stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(mtiles * N)) ncols <<= 1;

    const int pbeg = lbid * p.pix_per_cta;
    const int pend = min(p.P, pbeg + p.pix_per_cta);
    const int nchunks = (pend - pbeg + PIX - 1) / PIX;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(full_bar(s), kProducerThreads); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();
    if (threadIdx.x == 0) tc2::trace_ev(tlog, 0, -1);

    if (warp == 9) {
        // ---------------- TMA producer ----------------
        // one lane per slab: a single thread issuing all (up to 24) tensor copies of a stage, with the tap / coordinate
        // arithmetic in front of each, took 1.6 us per stage -- longer than the copies' latency -- and serialised the ring
        {
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapX)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapDY)) : "memory");
            }
            const int ppi = p.flat ? p.P : g.Ho * g.Wo;
            const int OW = p.flat ? p.P : g.Wo;
            int cx = 0, dx = 0, dy_ = 0;
            if (lane < ng) {
                const int grp = g0 + lane;
                const int tap = grp / p.cgroups, cg = grp - tap * p.cgroups;
                const int ky = tap / g.kw, kx = tap - ky * g.kw;
                cx = cg * 32; dx = kx - g.pl; dy_ = ky - g.pt;
            }
            for (int c = 0; c < nchunks; ++c) {
                const int p0 = pbeg + c * PIX;
                const int b0 = p0 / ppi, rem = p0 - b0 * ppi;
                const int oy0 = rem / OW, ox0 = rem - oy0 * OW;
                const int s = c % stages;
                const uint32_t ph = (uint32_t)((c / stages) & 1);
                if (lane == 0) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    tc2::trace_ev(tlog, 1, c);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(raw_bar(s)), "r"((uint32_t)stage_bytes) : "memory");
                }
                __syncwarp();
                const uint32_t sa = smem_u32(smem + s * stage_bytes) + (uint32_t)lane * kSlabB;
                if (lane < ng) {
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                        ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&mapX)), "r"(raw_bar(s)),
                          "r"(cx), "r"(ox0 * g.sw + dx), "r"(oy0 * g.sh + dy_), "r"(b0) : "memory");
                } else if (lane < ng + nb) {
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                        ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&mapDY)), "r"(raw_bar(s)),
                          "r"((lane - ng) * 32), "r"(p0) : "memory");
                }
            }
            tc2::trace_ev(tlog, 11, 0);
        }
    } else if (warp < 4) {
        // ---------------- transform: thread = (pixel row, 32-byte unit), PIX/32 rows per slab ----------------
        const int unit = threadIdx.x & 3, r0 = threadIdx.x >> 2;
        const bool do_bias = p.dbias != nullptr && blockIdx.y == 0;
        const int gate_ppi = g.Ho * g.Wo;
        float bsum[NBMAX][8];
#pragma unroll
        for (int i = 0; i < NBMAX; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[i][j] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % stages;
            const uint32_t ph = (uint32_t)((c / stages) & 1);
            mbar_wait(raw_bar(s), ph);
            if (threadIdx.x == 0) tc2::trace_ev(tlog, 2, c);
            uint8_t* st = smem + s * stage_bytes;
            const int p0 = pbeg + c * PIX;
            if (NBMAX == 1 && PIX == 128 && ng == 1) {
                // 1x1 convolution with 32 output channels (every mobilenetV3 block): one x slab + one dy slab, four pixel rows
                // per thread -- all sixteen 16-byte loads in flight, ONE __syncwarp, then the permuted stores
                float4 v[8][2];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int px = r0 + 32 * (u >> 1);
                    const float4* q = reinterpret_cast<const float4*>(st + (u & 1) * kSlabB + (uint32_t)px * 128u + ((uint32_t)unit << 5));
                    v[u][0] = q[0]; v[u][1] = q[1];
                }
                if (p.gate) {
#pragma unroll
                    for (int u = 0; u < 8; u += 2) {
                        const int pp = p0 + r0 + 32 * (u >> 1);
                        if (pp < p.P) {
                            const float* grow = p.gate + (long long)(pp / gate_ppi) * g.Cin + unit * 8 + (g0 % p.cgroups) * 32;
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(grow));
                            const float4 a1 = __ldg(reinterpret_cast<const float4*>(grow) + 1);
                            v[u][0].x *= a0.x; v[u][0].y *= a0.y; v[u][0].z *= a0.z; v[u][0].w *= a0.w;
                            v[u][1].x *= a1.x; v[u][1].y *= a1.y; v[u][1].z *= a1.z; v[u][1].w *= a1.w;
                        }
                    }
                }
                if (do_bias) {
#pragma unroll
                    for (int u = 1; u < 8; u += 2) {
                        bsum[0][0] += v[u][0].x; bsum[0][1] += v[u][0].y; bsum[0][2] += v[u][0].z; bsum[0][3] += v[u][0].w;
                        bsum[0][4] += v[u][1].x; bsum[0][5] += v[u][1].y; bsum[0][6] += v[u][1].z; bsum[0][7] += v[u][1].w;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int px = r0 + 32 * (u >> 1);
                    float4* d = reinterpret_cast<float4*>(st + (u & 1) * kSlabB + (uint32_t)px * 128u + ((uint32_t)(unit ^ (px & 3)) << 5));
                    d[0] = tf32_rn4(v[u][0]); d[1] = tf32_rn4(v[u][1]);
                }
            } else {
#pragma unroll
            for (int j = 0; j < PIX / 32; ++j) {
                const int px = r0 + 32 * j;
                const uint32_t src = (uint32_t)px * 128u + ((uint32_t)unit << 5);
                const uint32_t dst = (uint32_t)px * 128u + ((uint32_t)(unit ^ (px & 3)) << 5);
                const int pp = p0 + px;
                const float* grow = (p.gate && pp < p.P) ? p.gate + (long long)(pp / gate_ppi) * g.Cin + unit * 8 : nullptr;
                for (int gi = 0; gi < ng; ++gi) {
                    float4* q = reinterpret_cast<float4*>(st + gi * kSlabB + src);
                    float4 v0 = q[0], v1 = q[1];
                    if (grow) {
                        const int cg = (g0 + gi) % p.cgroups;
                        const float4 a0 = __ldg(reinterpret_cast<const float4*>(grow + cg * 32));
                        const float4 a1 = __ldg(reinterpret_cast<const float4*>(grow + cg * 32) + 1);
                        v0.x *= a0.x; v0.y *= a0.y; v0.z *= a0.z; v0.w *= a0.w;
                        v1.x *= a1.x; v1.y *= a1.y; v1.z *= a1.z; v1.w *= a1.w;
                    }
                    __syncwarp();
                    float4* d = reinterpret_cast<float4*>(st + gi * kSlabB + dst);
                    d[0] = tf32_rn4(v0); d[1] = tf32_rn4(v1);
                }
#pragma unroll
                for (int bi = 0; bi < NBMAX; ++bi) {
                    if (bi < nb) {
                        float4* q = reinterpret_cast<float4*>(st + (ng + bi) * kSlabB + src);
                        const float4 v0 = q[0], v1 = q[1];
                        if (do_bias) {
                            bsum[bi][0] += v0.x; bsum[bi][1] += v0.y; bsum[bi][2] += v0.z; bsum[bi][3] += v0.w;
                            bsum[bi][4] += v1.x; bsum[bi][5] += v1.y; bsum[bi][6] += v1.z; bsum[bi][7] += v1.w;
                        }
                        __syncwarp();
                        float4* d = reinterpret_cast<float4*>(st + (ng + bi) * kSlabB + dst);
                        d[0] = tf32_rn4(v0); d[1] = tf32_rn4(v1);
                    }
                }
            }
            }
            fence_proxy_async();
            mbar_arrive(full_bar(s));
            if (threadIdx.x == 0) tc2::trace_ev(tlog, 7, c);
        }
        if (do_bias) {
            // lanes with the same 32-byte unit (lane & 3) hold partial sums of the same 8 channels: butterfly over the 8
            // pixel lanes first, then ONE shared-memory atomic per (warp, unit, channel) -- 128 threads hammering 32
            // addresses with float atomics took 6 us, longer than the rest of the kernel
#pragma unroll
            for (int bi = 0; bi < NBMAX; ++bi)
                if (bi < nb)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = bsum[bi][j];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (lane < 4) bias_red[warp * N + bi * 32 + unit * 8 + j] = v;     // plain store: [warp][channel]
                    }
            if (threadIdx.x == 0) tc2::trace_ev(tlog, 9, 0);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x; i < N; i += kProducerThreads)
                atomicAdd(p.dbias + i, (bias_red[i] + bias_red[N + i]) + (bias_red[2 * N + i] + bias_red[3 * N + i]));
        }
        if (threadIdx.x == 0) tc2::trace_ev(tlog, 10, 0);
        if (threadIdx.x == 96) tc2::trace_ev(tlog, 10, 3);
    } else if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % stages;
                const uint32_t ph = (uint32_t)((c / stages) & 1);
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                tc2::trace_ev(tlog, 8, c);
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                const uint32_t b_addr = a_addr + ng * kSlabB;
                for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
                    for (int k = 0; k < PIX / 8; ++k) {
                        const uint64_t da = make_desc(a_addr + mt * 4 * kSlabB + 1024u * k, ng == 1 ? 0u : (uint32_t)kSlabB, 512u, 1u);
                        const uint64_t db = make_desc(b_addr + 1024u * k, kSlabB, 512u, 1u);
                        umma_tf32(tmem_base + (uint32_t)(mt * N), da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(tfull_bar);
            tc2::trace_ev(tlog, 3, 0);
        }
    } else {
        // ---------------- epilogue: TMEM -> red.global.add ----------------
        const int q = warp - 4;
        mbar_wait(tfull_bar, 0u);
        tc_fence_after();
        if (threadIdx.x == 128) tc2::trace_ev(tlog, 4, 0);
        for (int mt = 0; mt < mtiles; ++mt) {
            const int gi = mt * 4 + q;
            const bool ok = gi < ng;
            const long long row = (long long)(g0 + gi) * 32 + lane;
            for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N + n0), rr);
                if (ok) {
                    float* dst = p.dw + row * N + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)        // 16-byte vector reductions: a quarter of the L2 atomic requests
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(rr[j])),
                                     "f"(__uint_as_float(rr[j + 1])), "f"(__uint_as_float(rr[j + 2])),
                                     "f"(__uint_as_float(rr[j + 3])) : "memory");
                }
            }
        }
    }

    if (threadIdx.x == 128) tc2::trace_ev(tlog, 5, 0);
    if (threadIdx.x == 160 || threadIdx.x == 224) tc2::trace_ev(tlog, 5, threadIdx.x);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && tlog) {
        tc2::trace_ev(tlog, 6, -1);
        const unsigned int n = min(trace_log.n, (unsigned)tc2::kTraceMax);
        p.trace[0] = n;
        for (unsigned int i = 0; i < n; ++i) { p.trace[1 + 3 * i] = trace_log.ev[i][0]; p.trace[2 + 3 * i] = trace_log.ev[i][1]; p.trace[3 + 3 * i] = trace_log.ev[i][2]; }
    }
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// PIX consecutive output pixels as a tb x th x tw box (same rule as the forward tiles)
static bool pix_geometry(int OH, int OW, int PIX, bool flat, int& tw, int& th, int& tb) {
    if (OW >= PIX) { if ((OW % PIX) && !flat) return false; tw = PIX; th = 1; tb = 1; return true; }
    if (PIX % OW) return false;
    tw = OW;
    const int rows = PIX / OW;
    if (OH >= rows) { if (OH % rows) return false; th = rows; tb = 1; return true; }
    if (rows % OH) return false;
    th = OH; tb = rows / OH;
    return true;
}

template <int PIX, int NBMAX>
static int plan(Params& p, CUtensorMap& mx, CUtensorMap& mdy, const float* x, const float* dy, int msp, int& psplits,
                size_t& smem) {
    const ConvGeom& g = p.g;
    const int nb = p.N / 32;
    const int OW = p.flat ? p.P : g.Wo, OH = p.flat ? 1 : g.Ho;
    if (!pix_geometry(OH, OW, PIX, p.flat != 0, p.tw, p.th, p.tb)) return MVAE_ERR_UNSUPPORTED;
    if (p.tw * g.sw > 256 || p.th * g.sh > 256 || p.tb > 256) return MVAE_ERR_UNSUPPORTED;
    {
        unsigned long long dims[4];
        if (p.flat) { dims[0] = g.Cin; dims[1] = (unsigned long long)p.P; dims[2] = 1; dims[3] = 1; }
        else { dims[0] = g.Cin; dims[1] = g.W; dims[2] = g.H; dims[3] = g.B; }
        const unsigned int box[4] = {32u, (unsigned)(p.tw * g.sw), (unsigned)(p.th * g.sh), (unsigned)p.tb};
        const unsigned int es[4] = {1u, (unsigned)g.sw, (unsigned)g.sh, 1u};
        if (!tma::encode_f32(&mx, x, 4, dims, box, es)) return MVAE_ERR_UNSUPPORTED;
    }
    {
        const unsigned long long dims[2] = {(unsigned long long)p.N, (unsigned long long)p.P};
        const unsigned int box[2] = {32u, (unsigned)PIX};
        if (!tma::encode_f32(&mdy, dy, 2, dims, box)) return MVAE_ERR_UNSUPPORTED;
    }
    const int stage_bytes = (p.groups_per_cta + nb) * PIX * 128;
    const int slack = (p.groups_per_cta == 1) ? 0 : 3 * PIX * 128;
    const int budget = (NBMAX == 1 && stage_bytes * 3 + slack <= 104 * 1024) ? 104 * 1024 : 200 * 1024;   // 2 CTAs/SM when it fits
    int stages = (budget - slack) / stage_bytes;
    if (stages > env_int("MVAE_WG_STAGES", 4)) stages = env_int("MVAE_WG_STAGES", 4);
    if (stages < 2) return MVAE_ERR_UNSUPPORTED;
    p.stages = stages;
    const int per_sm = budget == 104 * 1024 ? 2 : 1;
    // weight gradients run on side streams next to the dgrad chain (nothing waits for them until the optimiser): they take a
    // share of the SMs only, so the critical chain's CTAs always find free slots (MVAE_WGRAD_SMS overrides; default 32:
    // measured 1.75 ms vs 1.79 ms at 64 on cfg2, batched launches of deferred gradients get the whole GPU regardless)
    static int wg_sms = 0;
    if (!wg_sms) { const char* e = getenv("MVAE_WGRAD_SMS"); wg_sms = e ? atoi(e) : 32; if (wg_sms < 1 || wg_sms > kNumSMs) wg_sms = kNumSMs; }
    const int share = (g_wgrad_sms > 0 && g_wgrad_sms <= kNumSMs) ? g_wgrad_sms : wg_sms;
    psplits = share * per_sm / msp;
    if (psplits < 1) psplits = 1;
    const int maxs = ceil_div(p.P, 2 * PIX);          // at least two stages of work per CTA
    if (psplits > maxs) psplits = maxs;
    if (psplits < 1) psplits = 1;
    p.pix_per_cta = ceil_div(ceil_div(p.P, psplits), PIX) * PIX;
    psplits = ceil_div(p.P, p.pix_per_cta);
    smem = (size_t)stages * stage_bytes + slack + (3 * stages + 2) * 8 + p.N * 16 + 64 + 1024;
    return MVAE_OK;
}

template <int PIX, int NBMAX>
static int launch_batch(BatchW& bt, const int* psplits, int msp, size_t smem, cudaStream_t s) {
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel<PIX, NBMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    }
    // several problems share the launch (deferred weight gradients of one pyramid level, or one layer of every level): the
    // launch owns the whole GPU -- one wave of CTAs split over the problems in proportion to their pixel counts
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const long long budget = (long long)kNumSMs * per_sm / msp > 0 ? (long long)kNumSMs * per_sm / msp : 1;
    long long total_pix = 0;
    for (int l = 0; l < bt.n; ++l) total_pix += bt.p[l].P;
    bt.cta_begin[0] = 0;
    for (int l = 0; l < bt.n; ++l) {
        int ps = psplits[l];
        if (bt.n > 1) {
            Params& p = bt.p[l];
            ps = (int)(budget * p.P / (total_pix > 0 ? total_pix : 1));
            const int maxs = ceil_div(p.P, 2 * PIX);
            if (ps > maxs) ps = maxs;
            if (ps < 1) ps = 1;
            p.pix_per_cta = ceil_div(ceil_div(p.P, ps), PIX) * PIX;
            ps = ceil_div(p.P, p.pix_per_cta);
        }
        bt.cta_begin[l + 1] = bt.cta_begin[l] + ps;
    }
    dim3 grid(bt.cta_begin[bt.n], msp);
    MVAE_CUDA(launch_pdl(wgrad_tma_kernel<PIX, NBMAX>, grid, dim3(tc2::kThreads2), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

template <int PIX, int NBMAX>
static int launch(Params& p, const float* x, const float* dy, int msp, cudaStream_t s) {
    BatchW bt;
    int psplits;
    size_t smem;
    if (int e = plan<PIX, NBMAX>(p, bt.mx[0], bt.mdy[0], x, dy, msp, psplits, smem)) return e;
    bt.p[0] = p;
    bt.n = 1;
    return launch_batch<PIX, NBMAX>(bt, &psplits, msp, smem, s);
}

// 3x3 filters over 32 channels: the nine taps split over grid.y in groups of MVAE_WG3_GROUPS (0 = all nine in one CTA).  Two
// taps per CTA fit the 128-pixel stages of the 1x1 kernels (2 x-slabs + 1 dy slab = 48 KB a stage) -- a quarter of the
// barrier round trips per pixel of the 32-pixel stages -- at the price of reading dy once per tap group.
static int wgrad_few_groups(const ConvGeom& g, int N) {
    static int v = -1;
    if (v < 0) v = env_int("MVAE_WG3_GROUPS", 0);
    return (g.kh * g.kw > 1 && g.Cin == 32 && N == 32) ? v : 0;
}

// groups / slab-group splits of one problem (shared by the single and the batched entry)
static void wgrad_groups(const ConvGeom& g, int N, Params& q, int& msp) {
    q.cgroups = g.Cin / 32;
    q.groups = g.kh * g.kw * q.cgroups;
    int gmax = 4 * (512 / N);                    // TMEM: mtiles * N <= 512 columns, at most 16 slabs of shared memory
    if (gmax > 16) gmax = 16;
    const int few = wgrad_few_groups(g, N);
    if (few > 0) {
        q.groups_per_cta = few;
    } else {
        const int msplits = ceil_div(q.groups, gmax);
        q.groups_per_cta = ceil_div(ceil_div(q.groups, msplits), 4) * 4;
    }
    if (q.groups_per_cta > q.groups) q.groups_per_cta = q.groups;
    msp = ceil_div(q.groups, q.groups_per_cta);
    q.flat = (g.kh == 1 && g.kw == 1 && g.sh == 1 && g.sw == 1) ? 1 : 0;
}

template <int PIX, int NBMAX>
static int batched(int n, const ConvGeom* g, const float* const* x, const float* const* gate, const float* const* dy,
                   float* const* dw, float* const* dbias, cudaStream_t s) {
    BatchW bt;
    int psplits[kMaxBatch], msp0 = 0;
    size_t smem0 = 0;
    for (int l = 0; l < n; ++l) {
        Params& q = bt.p[l];
        int msp;
        q.g = g[l]; q.gate = gate ? gate[l] : nullptr; q.dw = dw[l]; q.dbias = dbias ? dbias[l] : nullptr; q.trace = g_trace;
        q.P = g[l].B * g[l].Ho * g[l].Wo; q.N = g[l].Cout;
        wgrad_groups(g[l], q.N, q, msp);
        size_t smem;
        if (int e = plan<PIX, NBMAX>(q, bt.mx[l], bt.mdy[l], x[l], dy[l], msp, psplits[l], smem)) return e;
        if (l == 0) { msp0 = msp; smem0 = smem; }
        else if (msp != msp0 || smem != smem0 || q.N != bt.p[0].N || q.groups_per_cta != bt.p[0].groups_per_cta ||
                 q.groups != bt.p[0].groups || q.stages != bt.p[0].stages)
            return MVAE_ERR_UNSUPPORTED;
    }
    bt.n = n;
    return launch_batch<PIX, NBMAX>(bt, psplits, msp0, smem0, s);
}

}  // namespace tcw2

int conv_wgrad_tc(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                  cudaStream_t s) {
    const int P = g.B * g.Ho * g.Wo, N = g.Cout;
    if (g.coord != 0 || (g.Cin % 32) != 0 || (N % 32) != 0 || N > 256 || P < 256) return MVAE_ERR_UNSUPPORTED;
    // dw too: the epilogue adds the accumulators with 16-byte red.global.add.v4.f32 (a 4-byte aligned sub-view of a packed
    // buffer takes the scalar CUDA-core kernel instead)
    if (!(tc::al16(x) && tc::al16(gate) && tc::al16(dy) && tc::al16(dw))) return MVAE_ERR_UNSUPPORTED;
    tcw::Params p;
    p.g = g; p.x = x; p.gate = gate; p.dy = dy; p.dw = dw; p.dbias = dbias; p.P = P; p.N = N;
    p.cgroups = g.Cin / 32;
    p.groups = g.kh * g.kw * p.cgroups;
    // TMEM: mtiles * N <= 512 columns  ->  groups per CTA <= 4 * (512 / N), and at most 16 slabs of shared memory
    int gmax = 4 * (512 / N);
    if (gmax > 16) gmax = 16;
    const int msplits = ceil_div(p.groups, gmax);
    p.groups_per_cta = ceil_div(ceil_div(p.groups, msplits), 4) * 4;
    if (tcw2::wgrad_few_groups(g, N) > 0) p.groups_per_cta = tcw2::wgrad_few_groups(g, N);
    if (p.groups_per_cta > p.groups) p.groups_per_cta = p.groups;
    const int msp = ceil_div(p.groups, p.groups_per_cta);
    {
        tcw2::Params q;
        q.g = g; q.gate = gate; q.dw = dw; q.dbias = dbias; q.P = P; q.N = N; q.trace = g_trace;
        q.cgroups = p.cgroups; q.groups = p.groups; q.groups_per_cta = p.groups_per_cta;
        q.flat = (g.kh == 1 && g.kw == 1 && g.sh == 1 && g.sw == 1) ? 1 : 0;
        const int slabs = p.groups_per_cta + N / 32;
        int r = MVAE_ERR_UNSUPPORTED;
        if (N == 32) {
            if (slabs <= 3) r = tcw2::launch<128, 1>(q, x, dy, msp, s);
            if (r == MVAE_ERR_UNSUPPORTED) r = tcw2::launch<32, 1>(q, x, dy, msp, s);
        } else {
            if (slabs <= 3) r = tcw2::launch<128, 8>(q, x, dy, msp, s);
            if (r == MVAE_ERR_UNSUPPORTED) r = tcw2::launch<32, 8>(q, x, dy, msp, s);
        }
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    p.a_slabs = ceil_div(p.groups_per_cta, 4) * 4;
    const int stage_bytes = (p.a_slabs + N / 32) * tcw::kSlab;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 4) p.stages = 4;
    if (p.stages < 2) return MVAE_ERR_UNSUPPORTED;
    int psplits = kNumSMs / msp;
    if (psplits < 1) psplits = 1;
    const int maxs = ceil_div(P, 4 * tcw::kPix);
    if (psplits > maxs) psplits = maxs;
    p.pix_per_cta = ceil_div(ceil_div(P, psplits), tcw::kPix) * tcw::kPix;
    psplits = ceil_div(P, p.pix_per_cta);
    const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 2) * 8 + N * 4 + 64 + 1024;
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(tcw::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MVAE_CUDA(cudaFuncSetAttribute(tcw::wgrad_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    dim3 grid(psplits, msp);
    MVAE_CUDA(launch_pdl(tcw::wgrad_tc_kernel, dim3(grid), dim3(tc::kThreads), smem, s, p));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

int conv_wgrad_tc_batched(int n, const ConvGeom* g, const float* const* x, const float* const* gate, const float* const* dy,
                          float* const* dw, float* const* dbias, cudaStream_t s) {
    if (n < 1 || n > tcw2::kMaxBatch) return MVAE_ERR_UNSUPPORTED;
    for (int l = 0; l < n; ++l) {
        const ConvGeom& gl = g[l];
        if (gl.coord != 0 || (gl.Cin % 32) != 0 || (gl.Cout % 32) != 0 || gl.Cout > 256) return MVAE_ERR_UNSUPPORTED;
        if (!(tc::al16(x[l]) && tc::al16(gate ? gate[l] : nullptr) && tc::al16(dy[l]) && tc::al16(dw[l]))) return MVAE_ERR_UNSUPPORTED;
    }
    const int N = g[0].Cout;
    const int few = tcw2::wgrad_few_groups(g[0], N);
    const int slabs = (few > 0 ? few : (g[0].kh * g[0].kw * (g[0].Cin / 32) > 16 ? 16 : g[0].kh * g[0].kw * (g[0].Cin / 32))) + N / 32;
    int r = MVAE_ERR_UNSUPPORTED;
    if (N == 32) {
        if (slabs <= 3) r = tcw2::batched<128, 1>(n, g, x, gate, dy, dw, dbias, s);
        if (r == MVAE_ERR_UNSUPPORTED) r = tcw2::batched<32, 1>(n, g, x, gate, dy, dw, dbias, s);
    } else {
        if (slabs <= 3) r = tcw2::batched<128, 8>(n, g, x, gate, dy, dw, dbias, s);
        if (r == MVAE_ERR_UNSUPPORTED) r = tcw2::batched<32, 8>(n, g, x, gate, dy, dw, dbias, s);
    }
    return r;
}

}  // namespace mvae

extern "C" long long mvae_tc_launch_count(void) { return mvae::g_tc_launches; }
extern "C" int mvae_set_wgrad_sm_share(int sms) { const int prev = mvae::g_wgrad_sms; mvae::g_wgrad_sms = sms; return prev; }
extern "C" int mvae_debug_trace(long long* buf) { mvae::g_trace = buf; return MVAE_OK; }
