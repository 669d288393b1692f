// Dense layers on the tensor cores (MVAE_PREC_TF32): the fused mu||logvar head of every encoder level
// (multiscale_vae.py:359-370) and the Dense that opens every decoder level (:402-406), forward and input gradient.
//
// The batch is the only row dimension (M = B <= a few hundred rows) while one of K / N is the flattened feature map
// (thousands to millions of columns), so the GEMMs are skinny and dominated by streaming the weight matrix once:
//     forward  y[M,N]  = x[M,K]  W[K,N]      W rows are reduction rows -> B operand MN-major (SWIZZLE_128B_BASE32B)
//     dgrad    dx[M,K] = dy[M,N] W[K,N]^T    W rows are output columns -> B operand K-major  (SWIZZLE_128B)
// so the Keras (in, out) kernel is consumed as stored by both passes.  Work is cut along the OUTPUT columns (tiles of NT)
// when there are many, and along the REDUCTION (split-K) when the output is small; split-K partial sums go to a caller
// workspace and a second small kernel adds them in a fixed order (deterministic), with bias / activation / activation
// gradient applied there.  The weight gradient stays on the convolution wgrad kernel (conv_tc.cu), which is this GEMM with
// both operands MN-major.
//
//   warps 0-7   producers: global -> registers (round-to-nearest TF32) -> shared memory in the canonical UMMA layouts
//   warps 8-11  epilogue : tcgen05.ld -> bias / activation / mask -> global (or raw partial sums -> workspace)
//   warp  12    MMA issue: one lane, tcgen05.mma.kind::tf32 M=128 x NT x 8, four per 32-wide reduction chunk
#include "common.cuh"
#include "umma.cuh"

namespace mvae {

extern long long g_tc_launches;

namespace dtc {

using namespace tc;

constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;
constexpr int kProducers = 256;
constexpr int kEpilogue = 128;
constexpr int kThreads = kProducers + kEpilogue + 32;

struct Params {
    const float* a;          // [M, lda]: x (forward) / dy (dgrad)
    const float* w;          // Keras (in, out) kernel, leading dimension ldw
    const float* bias;       // [NOUT] or null (applied when ksplits == 1)
    const float* act_out;    // dgrad: activation output of the producer of x, [M, NOUT], or null
    float* out;              // [M, NOUT] (ksplits == 1) or the workspace [ksplits, M, NOUT]
    int lda, ldw;
    int M, NOUT;
    int act, gact;
    int chunks, chunks_per_split, ksplits;       // reduction in chunks of 32
    int mtiles, ntiles, tiles;
};

// MODE 0 forward (B MN-major), MODE 1 dgrad (B K-major); NT = output columns per tile
template <int MODE, int NT>
__global__ void __launch_bounds__(kThreads, 1) dense_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int bbytes = NT * 128;
    constexpr int stage_bytes = kABytes + bbytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStages + 2 + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t ncols = 2 * NT < 32 ? 32 : 2 * NT;           // two accumulator stages (NT is a power of two)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), kProducers); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpilogue); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp < 8) {
        // ================================================ producers ==============================================
        const int t = threadIdx.x;
        const int r = t >> 1, half = t & 1;                         // A: row of the tile, which 64 bytes of its 128
        const uint32_t sw = (uint32_t)(r & 7);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int mt = tile % p.mtiles, rest = tile / p.mtiles;
            const int nt = rest % p.ntiles, ks = rest / p.ntiles;
            const int n0 = nt * NT;
            const int c0 = ks * p.chunks_per_split;
            const int c1 = min(p.chunks, c0 + p.chunks_per_split);
            const int m = mt * kTileM + r;
            const float* arow = m < p.M ? p.a + (long long)m * p.lda + half * 16 : nullptr;
            for (int c = c0; c < c1; ++c, ++it) {
                float4 av[4];
                if (arow) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) av[q] = __ldg(reinterpret_cast<const float4*>(arow + c * 32) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) av[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                constexpr int kPieces = NT * 8 / kProducers;        // 16-byte pieces of the B chunk per thread
                float4 bv[kPieces];
#pragma unroll
                for (int i = 0; i < kPieces; ++i) {
                    const int idx = t + i * kProducers;
                    if (MODE == 0) {
                        constexpr int per_row = NT >> 2;
                        const int kr = idx / per_row, c16 = idx - kr * per_row;
                        bv[i] = __ldg(reinterpret_cast<const float4*>(p.w + (long long)(c * 32 + kr) * p.ldw + n0 + c16 * 4));
                    } else {
                        const int n = idx >> 3, cc = idx & 7;
                        bv[i] = __ldg(reinterpret_cast<const float4*>(p.w + (long long)(n0 + n) * p.ldw + c * 32 + cc * 4));
                    }
                }
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)((it / kStages) & 1);
                mbar_wait(empty_bar(s), ph ^ 1u);
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + kABytes;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(sa + (uint32_t)r * 128u + (((uint32_t)(half * 4 + q) ^ sw) << 4)) = tf32_rn4(av[q]);
#pragma unroll
                for (int i = 0; i < kPieces; ++i) {
                    const int idx = t + i * kProducers;
                    uint32_t off;
                    if (MODE == 0) {
                        constexpr int per_row = NT >> 2;
                        const int kr = idx / per_row, c16 = idx - kr * per_row;
                        off = (uint32_t)(c16 >> 3) * 4096u + (uint32_t)kr * 128u +
                              (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
                    } else {
                        const int n = idx >> 3, cc = idx & 7;
                        off = (uint32_t)n * 128u + ((((uint32_t)cc) ^ ((uint32_t)n & 7u)) << 4);
                    }
                    *reinterpret_cast<float4*>(sb + off) = tf32_rn4(bv[i]);
                }
                fence_proxy_async();
                mbar_arrive(full_bar(s));
            }
        }
    } else if (warp == 12) {
        // ================================================ MMA issue ==============================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 0 ? 1u : 0u) << 16) | ((uint32_t)(NT >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tl) {
                const int ks = tile / (p.mtiles * p.ntiles);
                const int c0 = ks * p.chunks_per_split;
                const int c1 = min(p.chunks, c0 + p.chunks_per_split);
                const int as = tl & 1;
                const uint32_t aph = (uint32_t)((tl >> 1) & 1);
                mbar_wait(tempty_bar(as), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * NT);
                for (int c = c0; c < c1; ++c, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)((it / kStages) & 1);
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_desc(a_addr + 32u * k, 16u, 1024u);
                        const uint64_t db = (MODE == 0) ? make_desc(b_addr + 1024u * k, 4096u, 512u, 1u)
                                                        : make_desc(b_addr + 32u * k, 16u, 1024u);
                        umma_tf32(d_tmem, da, db, idesc, (c > c0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(s));
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else {
        // ================================================ epilogue ===============================================
        const int q = warp - 8;
        int tl = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tl) {
            const int mt = tile % p.mtiles, rest = tile / p.mtiles;
            const int nt = rest % p.ntiles, ks = rest / p.ntiles;
            const int as = tl & 1;
            const uint32_t aph = (uint32_t)((tl >> 1) & 1);
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const int m = mt * kTileM + q * 32 + lane;
            const bool ok = m < p.M;
            const bool fin = p.ksplits == 1;
#pragma unroll 1
            for (int j0 = 0; j0 < NT; j0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * NT + j0), rr);
                if (ok) {
                    const int n = nt * NT + j0;
                    const long long o = ((long long)ks * p.M + m) * p.NOUT + n;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]),
                                               __uint_as_float(rr[j + 3]));
                        if (fin) {
                            if (p.bias) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            }
                            if (p.act != MVAE_ACT_NONE) {
                                v.x = act_apply(v.x, p.act); v.y = act_apply(v.y, p.act);
                                v.z = act_apply(v.z, p.act); v.w = act_apply(v.w, p.act);
                            }
                            if (p.act_out) {
                                const float4 ov = __ldg(reinterpret_cast<const float4*>(p.act_out + o + j));
                                v.x *= act_grad_from_out(ov.x, p.gact); v.y *= act_grad_from_out(ov.y, p.gact);
                                v.z *= act_grad_from_out(ov.z, p.gact); v.w *= act_grad_from_out(ov.w, p.gact);
                            }
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
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// out[m,n] = epilogue( sum_s ws[s,m,n] ), the splits added in index order
__global__ void __launch_bounds__(256) dense_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                                           const float* __restrict__ act_out, float* __restrict__ out,
                                                           long long mn4, int nout, int ksplits, int act, int gact) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn4) return;
    const float4* w4 = reinterpret_cast<const float4*>(ws);
    // the splits are added in index order, eight loads in flight at a time (one load per trip made the kernel ksplits global
    // latencies long: 10.9 us for the 256 x 256 head of cfg2 level 0)
    float4 v = w4[i];
    int s = 1;
    for (; s + 8 <= ksplits; s += 8) {
        float4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = w4[(long long)(s + j) * mn4 + i];
#pragma unroll
        for (int j = 0; j < 8; ++j) { v.x += u[j].x; v.y += u[j].y; v.z += u[j].z; v.w += u[j].w; }
    }
    for (; s < ksplits; ++s) {
        const float4 u = w4[(long long)s * mn4 + i];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + (int)((i * 4) % nout)));
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    }
    if (act != MVAE_ACT_NONE) {
        v.x = act_apply(v.x, act); v.y = act_apply(v.y, act); v.z = act_apply(v.z, act); v.w = act_apply(v.w, act);
    }
    if (act_out) {
        const float4 ov = __ldg(reinterpret_cast<const float4*>(act_out) + i);
        v.x *= act_grad_from_out(ov.x, gact); v.y *= act_grad_from_out(ov.y, gact);
        v.z *= act_grad_from_out(ov.z, gact); v.w *= act_grad_from_out(ov.w, gact);
    }
    reinterpret_cast<float4*>(out)[i] = v;
}

struct Plan { int NT, mtiles, ntiles, ksplits, cps; };

// red: reduction length, nout: output columns.  Returns false when the shape is left to the SIMT path.
static bool make_plan(int M, int red, int nout, Plan& pl) {
    if (M < 1 || (red % 32) != 0 || (nout % 32) != 0) return false;
    pl.mtiles = ceil_div(M, kTileM);
    int NT = 256;
    while (NT > 32 && (nout % NT) != 0) NT >>= 1;
    // many output columns: column tiles fill the GPU (narrower tiles until there are ~100 of them)
    if (nout > 256)
        while (NT > 64 && (long long)pl.mtiles * (nout / NT) < 96) NT >>= 1;
    pl.NT = NT;
    pl.ntiles = nout / NT;
    const int chunks = red / 32;
    const int base = pl.mtiles * pl.ntiles;
    // few output tiles: split the reduction, at least two chunks per split
    int ks = 1;
    if (base <= kNumSMs / 4) {
        ks = kNumSMs / base;
        if (ks > chunks / 2) ks = chunks / 2;
        if (ks < 1) ks = 1;
    }
    pl.cps = ceil_div(chunks, ks);
    pl.ksplits = ceil_div(chunks, pl.cps);
    return true;
}

template <int MODE, int NT>
static int launch_nt(const Params& p, cudaStream_t s) {
    const size_t smem = (size_t)kStages * (kABytes + NT * 128) + 256 + 1024;
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(dense_tc_kernel<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int grid = p.tiles < kNumSMs ? p.tiles : kNumSMs;
    MVAE_CUDA(launch_pdl(dense_tc_kernel<MODE, NT>, dim3(grid), dim3(kThreads), smem, s, p));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

template <int MODE>
static int run(int M, int red, int nout, const float* a, int lda, const float* w, int ldw, const float* bias,
               const float* act_out, int act, int gact, float* out, float* ws, size_t ws_bytes, cudaStream_t s) {
    Plan pl;
    if (!make_plan(M, red, nout, pl)) return MVAE_ERR_UNSUPPORTED;
    if (pl.ksplits > 1 && (!ws || ws_bytes < (size_t)pl.ksplits * M * nout * sizeof(float))) return MVAE_ERR_UNSUPPORTED;
    Params p;
    p.a = a; p.w = w; p.lda = lda; p.ldw = ldw; p.M = M; p.NOUT = nout;
    p.bias = bias; p.act_out = act_out; p.act = act; p.gact = gact;
    p.chunks = red / 32; p.chunks_per_split = pl.cps; p.ksplits = pl.ksplits;
    p.mtiles = pl.mtiles; p.ntiles = pl.ntiles; p.tiles = pl.mtiles * pl.ntiles * pl.ksplits;
    p.out = pl.ksplits > 1 ? ws : out;
    int e;
    switch (pl.NT) {
        case 256: e = launch_nt<MODE, 256>(p, s); break;
        case 128: e = launch_nt<MODE, 128>(p, s); break;
        case 64: e = launch_nt<MODE, 64>(p, s); break;
        default: e = launch_nt<MODE, 32>(p, s); break;
    }
    if (e) return e;
    if (pl.ksplits > 1) {
        const long long mn4 = (long long)M * nout / 4;
        MVAE_CUDA(launch_pdl(dense_reduce_kernel, dim3((unsigned)ceil_div(mn4, 256)), dim3(256), 0, s, (const float*)ws, bias,
                             act_out, out, mn4, nout, pl.ksplits, act, gact));
        MVAE_LAUNCH_CHECK();
    }
    return MVAE_OK;
}

}  // namespace dtc

size_t dense_tc_workspace_bytes(int M, int K, int N) {
    size_t need = 0;
    dtc::Plan pl;
    if (dtc::make_plan(M, K, N, pl) && pl.ksplits > 1) need = (size_t)pl.ksplits * M * N * sizeof(float);            // forward
    if (dtc::make_plan(M, N, K, pl) && pl.ksplits > 1) {                                                             // dgrad
        const size_t b = (size_t)pl.ksplits * M * K * sizeof(float);
        if (b > need) need = b;
    }
    return need;
}

// y[M,N] = act(x[M,K] W[K,N] + bias)
int dense_fwd_tc(int M, int K, int N, const float* x, const float* w, const float* bias, int act, float* y, float* ws,
                 size_t ws_bytes, cudaStream_t s) {
    return dtc::run<0>(M, K, N, x, K, w, N, bias, nullptr, act, MVAE_ACT_NONE, y, ws, ws_bytes, s);
}

// dx[M,K] = (dy[M,N] W[K,N]^T) * act'(act_out)
int dense_dgrad_tc(int M, int K, int N, const float* dy, const float* w, const float* act_out, int gact, float* dx, float* ws,
                   size_t ws_bytes, cudaStream_t s) {
    return dtc::run<1>(M, N, K, dy, N, w, N, nullptr, act_out, MVAE_ACT_NONE, gact, dx, ws, ws_bytes, s);
}

}  // namespace mvae
