// Fused mobilenetV3 block kernels (layer_blocks.py:556-648) for the 32-channel configurations, TF32 tensor cores.
//
// A mobilenetV3 block is  a = relu(x W0 + b0);  u = relu(dw3x3(a) + bd);  gate = SE(mean_hw u);  y = (u * gate) W2 + b2 + x.
// The squeeze-excite gate needs the BatchNorm statistics of the whole batch, so a block has exactly one batch-wide
// dependency in the forward pass (gap -> gate) and one in the backward pass (dgate -> dgap).  Everything between two such
// points is local to a tile of pixels, and these kernels run it in ONE launch with the intermediates in shared memory /
// tensor memory instead of one launch (and one HBM round trip) per layer:
//
//   forward  kernel = [F2 of block j-1: y = (u*gate) W2 + b2 + x]  then  [F1 of block j: a, u = relu(dw(a)+bd), gap sums]
//   backward kernel = [B2 of block j+1: dv = dy W2^T, d_pre = (dv*gate + dgap)(u>0), da = dw^T(d_pre)(a>0), dwd, dbd,
//                                       dx = da W0^T + dy]         then  [B1 of block j: dgate = sum_hw (dx W2^T) * u]
//
// so a chain of n blocks is n+1 launches forward and n+1 backward (plus the tiny squeeze-excite gate kernels in between)
// instead of 4n and 5n.  A tile is 256 pixels: whole images when H*W <= 256 (16x16: one image ... 1x1: 256 images), else a
// strip of 256/W rows plus one halo row above and below whose 1x1 convolutions are recomputed (the halo of the depthwise
// window).  Tiles land in shared memory by TMA (SWIZZLE_128B: a pixel is one 128-byte row = the K-major UMMA operand
// layout), are rounded to TF32 in place, multiplied on the tensor core (tcgen05.mma.kind::tf32, M = 128 per block of rows,
// N = K = 32, accumulators in tensor memory), read back with tcgen05.ld (one pixel row per thread), and leave by TMA store.
// The 1x1 weight gradients are not computed here: they are off the critical path and go out as the engine's deferred,
// batched mvae_conv2d_wgrad launches on the x / u / da / dy tensors these kernels write.
#include <string.h>
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mvae {

extern long long g_tc_launches;
extern long long* g_trace;        // debug timeline buffer (mvae_debug_trace)

namespace mb {
using namespace tc;

constexpr int kThreads = 256;
constexpr int kC = 32;                 // channels (Cin == filters == 32)
constexpr int kBlk = 128 * 128;        // bytes of one 128-row block of a tile

struct Geom {
    int B, H, W;
    int nb;        // images per tile
    int R;         // main rows (per image) of a tile
    int halo;      // 0: whole images, 1: strips with one recomputed row above and below
    int TH;        // R + 2 * halo rows per image in the tile
    int rows;      // nb * TH * W rows landed by the TMA box
    int nm;        // 128-row blocks: ceil(rows / 128)
    int strips;    // tiles per image (strip mode) or 1
    int tiles;
    int main_off;  // first main row of the tile buffer (halo * W)
    int lgW;       // log2(W)            (every supported W is a power of two)
    int lgPpi;     // log2(R * W)        main pixels per image of a tile
    int main_px;   // nb * R * W main pixels of a tile (<= 256)
    int seg;       // rows of one depthwise work item: min(R, 8)
    int lgNseg;    // log2(R / seg)
};

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// geometry of the 256-pixel tile; false when the shape is not supported (the caller then uses the per-layer kernels)
static bool make_geom(int B, int H, int W, Geom& g) {
    if (B <= 0 || H <= 0 || W <= 0) return false;
    g.B = B; g.H = H; g.W = W;
    const int hw = H * W;
    if (hw <= 256) {
        if (256 % hw) return false;
        // at most 8 images per tile: small images then spread over more CTAs (the per-image squeeze-excite work of a tile
        // stays small) at the price of partly filled 128-row blocks
        g.nb = 256 / hw < 8 ? 256 / hw : 8; g.R = H; g.halo = 0; g.strips = 1;
        g.tiles = (B + g.nb - 1) / g.nb;
    } else {
        if (W > 128 || (256 % W) || (W % 8)) return false;
        // main pixels per strip: 128 for rows of up to 32 pixels (4 rows + 2 halo rows = 192 tile rows: two CTAs per SM;
        // measured 1.37 vs 1.40 ms per cfg2 step against 256-pixel strips, which fit one CTA per SM), else 256.
        // MVAE_MBV3_STRIP_PX overrides.
        int strip_px = env_int("MVAE_MBV3_STRIP_PX", W <= 32 ? 128 : 256);
        if ((strip_px != 128 && strip_px != 256) || strip_px < W) strip_px = 256;
        g.R = strip_px / W;
        if (H % g.R) return false;
        g.nb = 1; g.halo = 1; g.strips = H / g.R;
        g.tiles = B * g.strips;
    }
    g.TH = g.R + 2 * g.halo;
    g.rows = g.nb * g.TH * W;
    g.nm = (g.rows + 127) / 128;
    g.main_off = g.halo * W;
    if ((W & (W - 1)) || (g.R & (g.R - 1))) return false;
    g.lgW = ilog2(W); g.lgPpi = ilog2(g.R * W);
    g.main_px = g.nb * g.R * W;
    g.seg = g.R < 8 ? g.R : 8;
    g.lgNseg = ilog2(g.R / g.seg);
    return g.nm <= 4;
}

// Squeeze-excite folded into the tile kernels (whole-image tiles only: a CTA then owns complete images).  The gate
//     gate = hard_sigmoid( BN_batch( relu(gap W0 + b0) ) W1 + b1 )                     (layer_blocks.py:418-462)
// splits at its one batch-wide step, the BatchNorm statistics: F1 of block j computes h = relu(gap W0 + b0) of ITS images
// and writes it to ws; the next launch (F2 of block j) has every CTA reduce h over the whole batch (B x 32 floats, L2) for the
// statistics and finish the gate of its own images.  Backward alike: B1 of block j writes ds = dgate * hsig'(.) and
// dhn = ds W1^T of its images, B2 of block j reduces the two BatchNorm-backward sums over the batch and finishes dgap of its
// images.  ws is the scratch of mvae_se_gate_fwd / _bwd (floats, n = B*32):
//     gap[n] h1[n] dhn[n] s[n] ds[n] (n unused) mean[32] rstd[32]
// so the squeeze-excite WEIGHT gradients still come from mvae_se_gate_bwd (on a side stream, off the critical path).
// The batch-wide sums themselves (sum h, sum h^2; sum dhn*xh, sum dhn: 2 x 32 numbers each) are accumulated by the producing
// launch with one double-precision atomic per CTA and channel into a zeroed 64-double buffer, so the consuming launch reads
// 512 bytes instead of re-reducing B x 32 floats in every CTA (double: E[h^2] - mean^2 without cancellation trouble, and the
// result does not depend on the order of the atomics once rounded to float).
struct SeFwd {
    const float* w0; const float* b0; float* ws; double* stat; float inv_hw;            // F1 (block j)
    const float* gamma; const float* beta; const float* w1; const float* b1; float* mm; float* mv;   // F2 (block j-1)
    float* ws_prev; const double* stat_prev; float* gate_out; float eps, momentum; int training;
    int fold_f1, fold_f2;
};
struct SeBwd {
    const float* w1_prev; float* ws_prev; double* bstat_prev; int fold_b1;              // B1 (block j)
    const float* w0; const float* gamma; const float* ws; const double* bstat; float inv_hw; int fold_b2;    // B2 (block j+1)
};

struct FwdParams {
    Geom g;
    const float* gate;     // F2: (B, 32) gate of block j-1 (read when the gate is not folded in)
    const float* w2; const float* b2;
    const float* w0; const float* b0; const float* wd; const float* bd;
    float* gap;            // F1: (B, 32) zeroed GAP sums of block j (when the gate is not folded in)
    int has_f2, has_f1, store_a;
    SeFwd se;
    long long* trace;
    int pdl;               // launched as a programmatic dependent: weights first, griddepcontrol.wait before the first tile
};
struct FwdMaps { CUtensorMap u_in, x_in, y_out, a_out, u_out; };

struct BwdParams {
    Geom g;
    // B2 (block j+1)
    const float* dy;       // its output gradient (global copy of the tile the TMA lands, for the residual term)
    const float* gate; const float* dgap;
    const float* w2; const float* wd; const float* w0;
    float* dwd; float* dbd;
    // B1 (block j)
    const float* w2p;      // conv2 kernel of block j
    const float* up;       // u of block j
    float* dgate;          // (B, 32) zeroed (folded gate: plainly stored)
    int has_b2, has_b1;
    SeBwd se;
    long long* trace;
    int pdl;
};
struct BwdMaps { CUtensorMap dy_in, u_in, a_in, da_out, dx_out, up_in; };

// One launch can serve the same step of several chains (the pyramid levels: same layer, different image size and weights):
// CTAs [cta_begin[l], cta_begin[l+1]) work on problem l with its own geometry, tensor maps and pointers.  The coarse levels
// then share a handful of launches instead of queueing dozens of tiny ones each behind level 0's.
constexpr int kMaxBatch = 8;          // (kernel parameters: 8 x (6 tensor maps + ~300 bytes), well inside the 32 KB limit)
struct FwdBatch { FwdMaps mp[kMaxBatch]; FwdParams p[kMaxBatch]; int cta_begin[kMaxBatch + 1]; int n; };
struct BwdBatch { BwdMaps mp[kMaxBatch]; BwdParams p[kMaxBatch]; int cta_begin[kMaxBatch + 1]; int n; };

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
    return (uint32_t)row * 128u + ((((uint32_t)chunk) ^ ((uint32_t)row & 7u)) << 4);
}
// element (row, channel) of a swizzled tile
__device__ __forceinline__ uint32_t sw_el(int row, int c) { return sw_off(row, c >> 2) + (((uint32_t)c & 3u) << 2); }

// explicit shared-state-space accesses on 32-bit addresses (pointer arithmetic through the carved struct otherwise compiles
// to generic LD.E / ST.E)
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds1u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// a 32x32 Keras kernel W[ci][co] -> UMMA B operand of the forward product x W (K = ci): MN-major, SWIZZLE_128B_BASE32B
// (the global load and the shared store are separate calls: a kernel issues ALL its weight / vector loads first and only
// then the stores, so that the prologue pays one global round trip instead of one per matrix)
__device__ __forceinline__ float4 w_piece(const float* __restrict__ w) {          // this thread's 16 bytes of a 32x32 matrix
    return __ldg(reinterpret_cast<const float4*>(w) + threadIdx.x);
}
__device__ __forceinline__ void put_w_fwd(uint32_t dst, float4 v) {
    const int idx = threadIdx.x;                       // 256 pieces of 16 bytes
    const int kr = idx >> 3, c16 = idx & 7;
    const uint32_t off = (uint32_t)kr * 128u + (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
    sts4(dst + off, tf32_rn4(v));
}
__device__ __forceinline__ void put_w_dgrad(uint32_t dst, float4 v) {
    const int idx = threadIdx.x;
    sts4(dst + sw_off(idx >> 3, idx & 7), tf32_rn4(v));
}
__device__ __forceinline__ void stage_w_fwd(uint8_t* dst, const float* __restrict__ w) {
    const int idx = threadIdx.x;                       // 256 pieces of 16 bytes
    const int kr = idx >> 3, c16 = idx & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w + kr * kC + c16 * 4));
    const uint32_t off = (uint32_t)kr * 128u + (((((uint32_t)c16 >> 1) & 3u) ^ ((uint32_t)kr & 3u)) << 5) + (((uint32_t)c16 & 1u) << 4);
    *reinterpret_cast<float4*>(dst + off) = tf32_rn4(v);
}
// the same kernel as the B operand of the transposed product dy W^T (K = co): K-major, SWIZZLE_128B
__device__ __forceinline__ void stage_w_dgrad(uint8_t* dst, const float* __restrict__ w) {
    const int idx = threadIdx.x;
    const int n = idx >> 3, cc = idx & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w + n * kC + cc * 4));
    *reinterpret_cast<float4*>(dst + sw_off(n, cc)) = tf32_rn4(v);
}

// registers -> tensor memory, 32 lanes x 32 columns of 32 bits (the mirror of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// D[blk] (128 x 32, TMEM columns blk*32..) = A[blk] (128 rows x 32, K-major) * B (32 x 32); one thread issues.
// accumulate: D += ... on top of what the accumulator columns already hold (the residual tile written by tmem_st32)
__device__ __forceinline__ void issue_mma(uint32_t a_addr, uint32_t b_addr, bool b_mn_major, uint32_t tmem_base, int nm, uint32_t bar,
                                          bool accumulate = false) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(kC >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    tc_fence_after();
    for (int blk = 0; blk < nm; ++blk) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(a_addr + (uint32_t)blk * kBlk + 32u * k, 16u, 1024u);
            const uint64_t db = b_mn_major ? make_desc(b_addr + 1024u * k, 4096u, 512u, 1u) : make_desc(b_addr + 32u * k, 16u, 1024u);
            umma_tf32(tmem_base + (uint32_t)(blk * kC), da, db, idesc, (k > 0 || accumulate) ? 1u : 0u);
        }
    }
    umma_commit(bar);
}

// debug timeline: thread 0 of CTA 0 appends (event, kernel tag, globaltimer ns) triples (scripts/trace_fused.py)
__device__ __forceinline__ void trace(long long* tr, int ev, int tag) {
    if (tr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const long long i = tr[0];
        if (i < 900) { tr[1 + 3 * i] = ev; tr[2 + 3 * i] = tag; tr[3 + 3 * i] = (long long)t; tr[0] = i + 1; }
    }
}

struct RowInfo { int b, iy, tx; bool valid, main; };
__device__ __forceinline__ RowInfo row_info(const Geom& g, int b0, int y0, int r) {
    RowInfo ri;
    int bi = 0, rem = r;
    if (!g.halo) { bi = r >> g.lgPpi; rem = r & ((1 << g.lgPpi) - 1); }       // whole images: TH * W == R * W
    const int ty = rem >> g.lgW;
    ri.tx = rem & (g.W - 1);
    ri.b = b0 + bi; ri.iy = y0 - g.halo + ty;
    ri.valid = r < g.rows && ri.b < g.B && ri.iy >= 0 && ri.iy < g.H;
    ri.main = r < g.rows && ty >= g.halo && ty < g.halo + g.R;
    return ri;
}

// shared memory: three tile buffers (nm * 16 KB each, 1024-byte aligned: SWIZZLE_128B atoms), three 4 KB weight operands,
// vectors, per-warp partial sums, two mbarriers, the TMEM base slot.  All as 32-bit shared addresses.
struct Smem {
    uint32_t buf[3];
    uint32_t wa, wb, wc;
    uint32_t vec;        // 8 x 32 floats: b2, b0, bd, then the folded gate's per-channel vectors
    uint32_t dww;        // 9 x 32 floats: depthwise taps
    uint32_t part;       // 8 x 32 floats
    uint32_t img;        // 2 x nb x 32 floats: per-image vectors of the folded gate (gate / sums, dgap)

    uint32_t bar_ld, bar_mma;
    uint32_t tmem_slot;
    uint8_t* base;       // generic pointer of buf[0] (scratch use after the tile loop)
};
__device__ __forceinline__ Smem carve(uint8_t* raw, int nm, int nb) {
    Smem s;
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    s.base = p;
    uint32_t a = smem_u32(p);
    for (int i = 0; i < 3; ++i) { s.buf[i] = a; a += (uint32_t)nm * kBlk; }
    s.wa = a; a += 4096; s.wb = a; a += 4096; s.wc = a; a += 4096;
    s.vec = a; a += 8 * kC * 4;
    s.dww = a; a += 9 * kC * 4;
    s.part = a; a += 8 * kC * 4;
    s.bar_ld = a; s.bar_mma = a + 8; a += 16;
    s.tmem_slot = a; a += 16;
    s.img = a;
    return s;
}
static size_t smem_bytes(int nm, int nb) {
    return (size_t)3 * nm * kBlk + 3 * 4096 + (8 + 9 + 8) * kC * 4 + 32 + (size_t)2 * nb * kC * 4 + 1024;
}

// forward: one accumulator region (nm x 32 columns); backward: a second one that starts out holding the residual tile
__device__ __forceinline__ uint32_t tmem_cols(int nm, bool two = false) { return (nm <= 2 ? 64u : 128u) << (two ? 1 : 0); }

// common prologue: barriers, tensor memory; returns the TMEM base address
__device__ __forceinline__ uint32_t setup(const Smem& s, int nm, bool two = false) {
    if (threadIdx.x == 0) {
        mbar_init(s.bar_ld, 1);
        mbar_init(s.bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s.tmem_slot), "r"(tmem_cols(nm, two))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return lds1u(s.tmem_slot);
}
__device__ __forceinline__ void teardown(uint32_t tmem_base, int nm, bool two = false) {
    // the bulk stores only have to be done READING shared memory before the CTA's allocation goes away
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols(nm, two)) : "memory");
    }
}

__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int y, int b) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(0), "r"(0), "r"(y), "r"(b) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* m, uint32_t src, int y, int b) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(0), "r"(0), "r"(y), "r"(b) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// warm L1 with a 32x32 matrix (4 KB) that a later phase of the launch walks row by row
__device__ __forceinline__ void prefetch_matrix(const float* w) {
    if (threadIdx.x < kC) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + threadIdx.x * kC) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tile_origin(const Geom& g, int tile, int& b0, int& y0) {
    if (g.halo) { b0 = tile / g.strips; y0 = (tile - b0 * g.strips) * g.R; }
    else { b0 = tile * g.nb; y0 = 0; }
}

// in-place round-to-nearest TF32 of the landed tile, optionally times the per-image gate (one 128-byte row per thread step)
__device__ __forceinline__ void round_tile(const Geom& g, uint32_t buf, const float* __restrict__ gate, int b0,
                                           uint32_t gate_s = 0u) {
    for (int r = threadIdx.x; r < g.rows; r += kThreads) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = lds4(buf + sw_off(r, q));
        if (gate_s) {
            const uint32_t gr = gate_s + (uint32_t)(r >> g.lgPpi) * 128u;      // folded gate: whole-image tiles only
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 gt = lds4(gr + q * 16);
                v[q].x *= gt.x; v[q].y *= gt.y; v[q].z *= gt.z; v[q].w *= gt.w;
            }
        } else if (gate) {
            const int bi = g.halo ? 0 : (r >> g.lgPpi);
            const float4* gr = reinterpret_cast<const float4*>(gate + (long long)min(b0 + bi, g.B - 1) * kC);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 gt = __ldg(gr + q);
                v[q].x *= gt.x; v[q].y *= gt.y; v[q].z *= gt.z; v[q].w *= gt.w;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) sts4(buf + sw_off(r, q), tf32_rn4(v[q]));
    }
}

// Per-image sums over the main pixels of a tile held in `buf` (lane = channel; a warp adds its 32 consecutive main pixels,
// images smaller than a warp's run are flushed as they end), added into out[b][c].  Contains a __syncthreads.
__device__ __forceinline__ void image_sums(const Geom& g, const Smem& s, uint32_t buf, int b0, float* __restrict__ out,
                                           uint32_t out_s = 0u) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ppi = 1 << g.lgPpi;
    const uint32_t col = (((uint32_t)lane & 3u) << 2);
    const int q = lane >> 2;
    const int m0 = warp * 32;
    if (ppi >= 32) {
        // an image spans whole warps: 32 independent loads, four accumulators
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (m0 < g.main_px) {
            const int r0 = g.main_off + m0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                a0 += lds1(buf + sw_off(r0 + i, q) + col);
                a1 += lds1(buf + sw_off(r0 + i + 1, q) + col);
                a2 += lds1(buf + sw_off(r0 + i + 2, q) + col);
                a3 += lds1(buf + sw_off(r0 + i + 3, q) + col);
            }
        }
        sts1(s.part + (uint32_t)(warp * kC + lane) * 4u, (a0 + a1) + (a2 + a3));
    } else {
        // several images inside the warp's run of 32 pixels: one short sum per image
        const int nimg = max(0, min(32, g.main_px - m0)) >> g.lgPpi;
        for (int k = 0; k < nimg; ++k) {
            const int m = m0 + (k << g.lgPpi);
            float acc = 0.f;
            for (int i = 0; i < ppi; ++i) acc += lds1(buf + sw_off(g.main_off + m + i, q) + col);
            const int bi = m >> g.lgPpi;
            if (out_s) sts1(out_s + (uint32_t)(bi * kC + lane) * 4u, acc);
            else if (b0 + bi < g.B) atomicAdd(out + (long long)(b0 + bi) * kC + lane, acc);
        }
    }
    __syncthreads();
    if (ppi >= 32) {
        const int wpi = ppi >> 5;                  // warps per image
        const int bi = warp;                       // thread (bi, c) for bi < nb
        if (bi < g.nb) {
            float t = 0.f;
            for (int w = bi * wpi; w < (bi + 1) * wpi; ++w) t += lds1(s.part + (uint32_t)(w * kC + lane) * 4u);
            if (out_s) sts1(out_s + (uint32_t)(bi * kC + lane) * 4u, t);
            else if (b0 + bi < g.B) atomicAdd(out + (long long)(b0 + bi) * kC + lane, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// folded squeeze-excite gate (whole-image tiles).  vec slots (32 floats each): 3 = mean, 4 = gamma * rstd, 5 = beta or
// sum(dhn * xh) / B, 6 = b1 or sum(dhn) / B, 7 = rstd
// ---------------------------------------------------------------------------------------------------------------------
// F2 prologue: BatchNorm statistics of h over the whole batch (training: from the sums F1 accumulated) or the moving
// statistics -> vec slots
__device__ __forceinline__ void se_fwd_stats(const Geom& g, const Smem& s, const SeFwd& se) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < kC) {
        float mean, var;
        if (se.training) {
            const double m = se.stat_prev[lane] / (double)g.B;
            const double v = se.stat_prev[kC + lane] / (double)g.B - m * m;
            mean = (float)m; var = v > 0.0 ? (float)v : 0.f;
        } else {
            mean = se.mm[lane]; var = se.mv[lane];
        }
        const float rstd = rsqrtf(var + se.eps);
        sts1(s.vec + (3 * kC + lane) * 4, mean);
        sts1(s.vec + (4 * kC + lane) * 4, __ldg(se.gamma + lane) * rstd);
        sts1(s.vec + (5 * kC + lane) * 4, __ldg(se.beta + lane));
        sts1(s.vec + (6 * kC + lane) * 4, __ldg(se.b1 + lane));
    }
    __syncthreads();
}
// CTA 0, after its tiles: the statistics go on record for the backward pass and into the moving averages (a read-modify-
// write round trip that nothing in the launch waits for)
__device__ __forceinline__ void se_fwd_record(const Geom& g, const SeFwd& se, int lbid) {
    const int lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    if (lbid == 0 && threadIdx.x < kC) {
        float mean, var;
        if (se.training) {
            const double m = se.stat_prev[lane] / (double)g.B;
            const double v = se.stat_prev[kC + lane] / (double)g.B - m * m;
            mean = (float)m; var = v > 0.0 ? (float)v : 0.f;
            se.mm[lane] = se.mm[lane] * se.momentum + mean * (1.f - se.momentum);
            se.mv[lane] = se.mv[lane] * se.momentum + var * (1.f - se.momentum);
        } else {
            mean = se.mm[lane]; var = se.mv[lane];
        }
        se.ws_prev[6 * n + lane] = mean; se.ws_prev[6 * n + kC + lane] = rsqrtf(var + se.eps);
    }
}

// The per-image parts: warp bi works on image b0 + bi of the tile (at most 8 images), lane = channel; matrix rows are read
// coalesced and the other operand travels by shuffle (broadcast) or the product is summed over the warp.

// F2, per tile: gate of the tile's images into img[0 .. nb*32), the gate buffer and ws.s
__device__ __forceinline__ void se_fwd_gate(const Geom& g, const Smem& s, const SeFwd& se, int b0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    if (warp < g.nb) {
        const int b = b0 + warp;
        float gate = 0.f;
        if (b < g.B) {
            const float hv = __ldg(se.ws_prev + n + (long long)b * kC + lane);
            const float hn = fmaf(lds1(s.vec + (4 * kC + lane) * 4), hv - lds1(s.vec + (3 * kC + lane) * 4), lds1(s.vec + (5 * kC + lane) * 4));
            float a0 = lds1(s.vec + (6 * kC + lane) * 4), a1 = 0.f;
#pragma unroll 8
            for (int j = 0; j < kC; j += 2) {
                a0 = fmaf(__shfl_sync(0xffffffffu, hn, j), __ldg(se.w1 + j * kC + lane), a0);
                a1 = fmaf(__shfl_sync(0xffffffffu, hn, j + 1), __ldg(se.w1 + (j + 1) * kC + lane), a1);
            }
            const float acc = a0 + a1;
            gate = fminf(fmaxf(fmaf(0.2f, acc, 0.5f), 0.f), 1.f);
            se.ws_prev[3 * n + (long long)b * kC + lane] = acc;
            se.gate_out[(long long)b * kC + lane] = gate;
        }
        sts1(s.img + (uint32_t)(warp * kC + lane) * 4u, gate);
    }
    __syncthreads();
}

// F1 tail: the GAP sums of the tile's images sit in img[0 .. nb*32): gap mean and h = relu(gap W0 + b0) -> ws, and this
// tile's share of the batch sums of h (one double atomic per channel and moment)
__device__ __forceinline__ void se_fwd_h(const Geom& g, const Smem& s, const SeFwd& se, int b0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    float hv = 0.f;
    if (warp < g.nb && b0 + warp < g.B) {
        const int b = b0 + warp;
        const float gm = lds1(s.img + (uint32_t)(warp * kC + lane) * 4u) * se.inv_hw;
        float a0 = __ldg(se.b0 + lane), a1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < kC; c += 2) {
            a0 = fmaf(__shfl_sync(0xffffffffu, gm, c), __ldg(se.w0 + c * kC + lane), a0);
            a1 = fmaf(__shfl_sync(0xffffffffu, gm, c + 1), __ldg(se.w0 + (c + 1) * kC + lane), a1);
        }
        hv = fmaxf(a0 + a1, 0.f);
        se.ws[(long long)b * kC + lane] = gm;
        se.ws[n + (long long)b * kC + lane] = hv;
    }
    sts1(s.part + (uint32_t)(warp * kC + lane) * 4u, hv);
    __syncthreads();
    if (threadIdx.x < 2 * kC) {
        const int sq = threadIdx.x >> 5;
        double acc = 0.0;
        for (int bi = 0; bi < g.nb; ++bi) {
            const float v = lds1(s.part + (uint32_t)(bi * kC + lane) * 4u);
            acc += sq ? (double)v * (double)v : (double)v;
        }
        atomicAdd(se.stat + sq * kC + lane, acc);
    }
}

// B1 tail: the gate-gradient sums of the tile's images sit in img[0 .. nb*32): ds = dgate * hard_sigmoid'(s) and
// dhn = ds W1^T -> ws; the sums themselves go to `dgate` for the weight-gradient kernel; this tile's share of the two
// BatchNorm-backward sums
__device__ __forceinline__ void se_bwd_ds(const Geom& g, const Smem& s, const SeBwd& se, float* __restrict__ dgate, int b0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    float dhn = 0.f, dxh = 0.f;
    if (warp < g.nb && b0 + warp < g.B) {
        const int b = b0 + warp;
        const float dg = lds1(s.img + (uint32_t)(warp * kC + lane) * 4u);
        const float hs = fmaf(0.2f, se.ws_prev[3 * n + (long long)b * kC + lane], 0.5f);
        const float ds = (hs >= 0.f && hs <= 1.f) ? 0.2f * dg : 0.f;
        const float hv = se.ws_prev[n + (long long)b * kC + lane];
        const float mean = se.ws_prev[6 * n + lane], rstd = se.ws_prev[6 * n + kC + lane];
        dgate[(long long)b * kC + lane] = dg;
        se.ws_prev[4 * n + (long long)b * kC + lane] = ds;
#pragma unroll 4
        for (int j = 0; j < kC; ++j) {
            const float r = warp_sum(ds * __ldg(se.w1_prev + j * kC + lane));      // dhn[j] = sum_c ds[c] W1[j][c]
            if (lane == j) dhn = r;
        }
        se.ws_prev[2 * n + (long long)b * kC + lane] = dhn;
        dxh = dhn * ((hv - mean) * rstd);
    }
    sts1(s.part + (uint32_t)(warp * kC + lane) * 4u, dxh);
    sts1(s.img + (uint32_t)((g.nb + warp) * kC + lane) * 4u, dhn);
    __syncthreads();
    if (threadIdx.x < 2 * kC) {
        const int which = threadIdx.x >> 5;          // 0: sum dhn * xh, 1: sum dhn
        double acc = 0.0;
        for (int bi = 0; bi < g.nb; ++bi)
            acc += (double)(which ? lds1(s.img + (uint32_t)((g.nb + bi) * kC + lane) * 4u) : lds1(s.part + (uint32_t)(bi * kC + lane) * 4u));
        atomicAdd(se.bstat_prev + which * kC + lane, acc);
    }
}

// B2 prologue: the two BatchNorm-backward sums over the whole batch (accumulated by B1)
__device__ __forceinline__ void se_bwd_stats(const Geom& g, const Smem& s, const SeBwd& se) {
    const int lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    if (threadIdx.x < kC) {
        const float mean = se.ws[6 * n + lane], rstd = se.ws[6 * n + kC + lane];
        sts1(s.vec + (3 * kC + lane) * 4, mean);
        sts1(s.vec + (4 * kC + lane) * 4, __ldg(se.gamma + lane) * rstd);
        sts1(s.vec + (5 * kC + lane) * 4, (float)(se.bstat[lane] / (double)g.B));
        sts1(s.vec + (6 * kC + lane) * 4, (float)(se.bstat[kC + lane] / (double)g.B));
        sts1(s.vec + (7 * kC + lane) * 4, rstd);
    }
    __syncthreads();
}

// B2, per tile: dgap of the tile's images into img[nb*32 .. 2*nb*32)
__device__ __forceinline__ void se_bwd_dgap(const Geom& g, const Smem& s, const SeBwd& se, int b0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)g.B * kC;
    if (warp < g.nb) {
        float dgap = 0.f;
        if (b0 + warp < g.B) {
            const int b = b0 + warp;
            const float hv = __ldg(se.ws + n + (long long)b * kC + lane);
            const float xh = (hv - lds1(s.vec + (3 * kC + lane) * 4)) * lds1(s.vec + (7 * kC + lane) * 4);
            float d = lds1(s.vec + (4 * kC + lane) * 4) *
                      (__ldg(se.ws + 2 * n + (long long)b * kC + lane) - lds1(s.vec + (6 * kC + lane) * 4) - xh * lds1(s.vec + (5 * kC + lane) * 4));
            d = hv > 0.f ? d : 0.f;                       // dp0[j], lane = j
#pragma unroll 4
            for (int c = 0; c < kC; ++c) {
                const float r = warp_sum(d * __ldg(se.w0 + c * kC + lane));     // sum_j dp0[j] W0[c][j]
                if (lane == c) dgap = r * se.inv_hw;
            }
        }
        sts1(s.img + (uint32_t)((g.nb + warp) * kC + lane) * 4u, dgap);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// depthwise stages.  A work item is (image, pixel column, channel group, run of SEG rows): the thread walks down the rows
// with a rolling 3x3 window in registers (three shared loads per output); the loop is unrolled over the run, so the window
// rotates by renaming and the loads of the next rows issue ahead of the arithmetic.
// ---------------------------------------------------------------------------------------------------------------------
template <int SEG>
__device__ __forceinline__ void dw_fwd_stage(const Geom& g, const Smem& s, uint32_t bufA, uint32_t bufU) {
    const int q = threadIdx.x & 7;                    // 16-byte channel group, the same for every item of this thread
    float4 w4[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w4[k] = lds4(s.dww + (uint32_t)(k * kC + q * 4) * 4u);
    const float4 b4 = lds4(s.vec + (uint32_t)(2 * kC + q * 4) * 4u);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nitems = (g.nb << (g.lgW + 3 + g.lgNseg));
    for (int item = threadIdx.x; item < nitems; item += kThreads) {
        int t = item >> 3;
        const int tx = t & (g.W - 1); t >>= g.lgW;
        const int sg = t & ((1 << g.lgNseg) - 1);
        const int bi = t >> g.lgNseg;
        const int ty0 = g.halo + sg * SEG;            // first output row (tile row index within the image)
        const int rbase = bi * g.TH;
        const bool hasl = tx > 0, hasr = tx + 1 < g.W;
        float4 win[3][3];
        auto load_row = [&](int ty, float4 (&dst)[3]) {
            if (ty >= 0 && ty < g.TH) {
                const int r = ((rbase + ty) << g.lgW) + tx;
                dst[0] = hasl ? lds4(bufA + sw_off(r - 1, q)) : z4;
                dst[1] = lds4(bufA + sw_off(r, q));
                dst[2] = hasr ? lds4(bufA + sw_off(r + 1, q)) : z4;
            } else {
                dst[0] = z4; dst[1] = z4; dst[2] = z4;
            }
        };
        load_row(ty0 - 1, win[0]);
        load_row(ty0, win[1]);
#pragma unroll
        for (int rr = 0; rr < SEG; ++rr) {
            load_row(ty0 + rr + 1, win[(rr + 2) % 3]);
            float4 acc = b4;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 wv = w4[ky * 3 + kx], av = win[(rr + ky) % 3][kx];
                    acc.x = fmaf(av.x, wv.x, acc.x); acc.y = fmaf(av.y, wv.y, acc.y);
                    acc.z = fmaf(av.z, wv.z, acc.z); acc.w = fmaf(av.w, wv.w, acc.w);
                }
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
            sts4(bufU + sw_off(((rbase + ty0 + rr) << g.lgW) + tx, q), acc);
        }
    }
}

// d_pre (bufU) -> da = dw^T(d_pre) * (a > 0): fp32 in place over a (bufA), TF32 into bufD; channel PAIRS per thread so that
// the nine weight-gradient accumulators stay in registers
template <int SEG>
__device__ __forceinline__ void dw_bwd_stage(const Geom& g, const Smem& s, uint32_t bufU, uint32_t bufA, uint32_t bufD, int cp,
                                             float2 (&dwd)[9], float2& dbd) {
    float2 w2[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w2[k] = lds2(s.dww + (uint32_t)(k * kC + cp * 2) * 4u);
    const uint32_t coff = ((uint32_t)cp & 1u) << 3;          // byte offset of the pair inside its 16-byte chunk
    const int q = cp >> 1;
    const float2 z2 = make_float2(0.f, 0.f);
    const int nitems = (g.nb << (g.lgW + 4 + g.lgNseg));
    for (int item = threadIdx.x; item < nitems; item += kThreads) {
        int t = item >> 4;
        const int tx = t & (g.W - 1); t >>= g.lgW;
        const int sg = t & ((1 << g.lgNseg) - 1);
        const int bi = t >> g.lgNseg;
        const int ty0 = g.halo + sg * SEG;
        const int rbase = bi * g.TH;
        const bool hasl = tx > 0, hasr = tx + 1 < g.W;
        float2 win[3][3];
        auto load_row = [&](int ty, float2 (&dst)[3]) {
            if (ty >= 0 && ty < g.TH) {
                const int r = ((rbase + ty) << g.lgW) + tx;
                dst[0] = hasl ? lds2(bufU + sw_off(r - 1, q) + coff) : z2;
                dst[1] = lds2(bufU + sw_off(r, q) + coff);
                dst[2] = hasr ? lds2(bufU + sw_off(r + 1, q) + coff) : z2;
            } else {
                dst[0] = z2; dst[1] = z2; dst[2] = z2;
            }
        };
        load_row(ty0 - 1, win[0]);
        load_row(ty0, win[1]);
#pragma unroll
        for (int rr = 0; rr < SEG; ++rr) {
            load_row(ty0 + rr + 1, win[(rr + 2) % 3]);
            const uint32_t ce = sw_off(((rbase + ty0 + rr) << g.lgW) + tx, q) + coff;
            const float2 av = lds2(bufA + ce);
            float2 acc = z2;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float2 d = win[(rr + 2 - ky) % 3][2 - kx];       // d_pre at (y - (ky-1), x - (kx-1))
                    const int k = ky * 3 + kx;
                    acc.x = fmaf(w2[k].x, d.x, acc.x); acc.y = fmaf(w2[k].y, d.y, acc.y);
                    dwd[k].x = fmaf(av.x, d.x, dwd[k].x); dwd[k].y = fmaf(av.y, d.y, dwd[k].y);
                }
            const float2 ctr = win[(rr + 1) % 3][1];
            dbd.x += ctr.x; dbd.y += ctr.y;
            const float2 da = make_float2(av.x > 0.f ? acc.x : 0.f, av.y > 0.f ? acc.y : 0.f);
            sts2(bufA + ce, da);
            sts2(bufD + ce, make_float2(tf32_rn(da.x), tf32_rn(da.y)));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) mbv3_fwd_kernel(const __grid_constant__ FwdBatch bt) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.cta_begin[lvl + 1]) ++lvl;
    const FwdParams& p = bt.p[lvl];
    const FwdMaps& mp = bt.mp[lvl];
    const int lbid = (int)blockIdx.x - bt.cta_begin[lvl], lgrid = bt.cta_begin[lvl + 1] - bt.cta_begin[lvl];
    const Geom& g = p.g;
    const Smem s = carve(smem_raw, g.nm, g.nb);
    const uint32_t bufU = s.buf[0], bufX = s.buf[1], bufA = s.buf[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long* const tr = p.trace;
    const int ttag = p.has_f2 * 2 + p.has_f1;
    trace(tr, 0, ttag);
    if (p.has_f2 && p.se.fold_f2) prefetch_matrix(p.se.w1);
    if (p.has_f1 && p.se.fold_f1) prefetch_matrix(p.se.w0);
    if (tid == 0) {
        prefetch_map(&mp.x_in);
        if (p.has_f2) { prefetch_map(&mp.u_in); prefetch_map(&mp.y_out); }
        if (p.has_f1) { prefetch_map(&mp.u_out); if (p.store_a) prefetch_map(&mp.a_out); }
    }
    if (p.pdl) pdl_launch();             // the next kernel of the chain may start its own prologue
    const uint32_t tmem_base = setup(s, g.nm);
    if (!p.pdl) pdl_sync();
    trace(tr, 1, ttag);
    const uint32_t tile_bytes = (uint32_t)g.rows * 128u;
    auto issue_loads = [&](int tile) {
        int b0, y0;
        tile_origin(g, tile, b0, y0);
        if (p.has_f2) {
            tma::mbar_expect_tx(s.bar_ld, 2 * tile_bytes);
            tma_load_tile(bufU, &mp.u_in, s.bar_ld, y0 - g.halo, b0);
            tma_load_tile(bufX, &mp.x_in, s.bar_ld, y0 - g.halo, b0);
        } else {
            tma::mbar_expect_tx(s.bar_ld, tile_bytes);
            tma_load_tile(bufU, &mp.x_in, s.bar_ld, y0 - g.halo, b0);
        }
    };
    // the first tile is on its way while the weights are staged (as a programmatic dependent the order is reversed: the
    // weights do not depend on the predecessor, the tile does)
    if (!p.pdl && tid == 0 && lbid < g.tiles) issue_loads(lbid);

    // weights (TF32, UMMA layouts), bias vectors and depthwise taps once per CTA: every global load is issued before the first
    // shared store (one round trip for the whole prologue)
    {
        float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f), r0 = r2;
        float v2 = 0.f, v0 = 0.f, vd = 0.f, d0 = 0.f, d1 = 0.f;
        if (p.has_f2) { r2 = w_piece(p.w2); if (tid < kC) v2 = __ldg(p.b2 + tid); }
        if (p.has_f1) {
            r0 = w_piece(p.w0);
            if (tid < kC) { v0 = __ldg(p.b0 + tid); vd = __ldg(p.bd + tid); d1 = __ldg(p.wd + kThreads + tid); }
            d0 = __ldg(p.wd + tid);
        }
        if (p.has_f2) { put_w_fwd(s.wa, r2); if (tid < kC) sts1(s.vec + tid * 4, v2); }
        if (p.has_f1) {
            put_w_fwd(s.wb, r0);
            if (tid < kC) { sts1(s.vec + (kC + tid) * 4, v0); sts1(s.vec + (2 * kC + tid) * 4, vd); sts1(s.dww + (kThreads + tid) * 4, d1); }
            sts1(s.dww + tid * 4, d0);
        }
    }
    fence_proxy_async();
    if (p.pdl) {
        pdl_wait();                                    // everything the predecessor wrote is visible from here on
        if (tid == 0 && lbid < g.tiles) issue_loads(lbid);
    }
    __syncthreads();
    const bool fold2 = p.has_f2 && p.se.fold_f2, fold1 = p.has_f1 && p.se.fold_f1;
    trace(tr, 2, ttag);
    if (fold2) se_fwd_stats(g, s, p.se);               // BatchNorm statistics of the whole batch, while the first tile lands
    trace(tr, 3, ttag);

    uint32_t ph_ld = 0, ph_mma = 0;
    bool first = true;
    for (int tile = lbid; tile < g.tiles; tile += lgrid) {
        int b0, y0;
        tile_origin(g, tile, b0, y0);
        if (!first) {
            // the TMA stores of the previous tile must have read their staging buffers before these are overwritten
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            if (tid == 0) issue_loads(tile);
        }
        first = false;
        if (fold2) se_fwd_gate(g, s, p.se, b0);        // gate of this tile's images (the tile is still in flight)
        trace(tr, 4, ttag);
        mbar_wait(s.bar_ld, ph_ld);
        ph_ld ^= 1u;
        trace(tr, 5, ttag);

        // ---- A operand of the first product: tf32(u * gate) (conv2 of block j-1) or tf32(x) (conv0 of the chain's first block)
        round_tile(g, bufU, (p.has_f2 && !fold2) ? p.gate : nullptr, b0, fold2 ? s.img : 0u);
        fence_proxy_async();
        __syncthreads();
        trace(tr, 6, ttag);

        if (p.has_f2) {
            if (tid == 0) issue_mma(bufU, s.wa, true, tmem_base, g.nm, s.bar_mma);
            mbar_wait(s.bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            trace(tr, 7, ttag);
            // ---- y = D + b2 + x: fp32 into bufX (in place, staging of the TMA store), TF32 into bufU (A operand of conv0)
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
                if (r < g.rows) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 xv = lds4(bufX + sw_off(r, q));
                        const float4 bv = lds4(s.vec + q * 16);
                        float4 y;
                        y.x = __uint_as_float(rr[4 * q]) + bv.x + xv.x;
                        y.y = __uint_as_float(rr[4 * q + 1]) + bv.y + xv.y;
                        y.z = __uint_as_float(rr[4 * q + 2]) + bv.z + xv.z;
                        y.w = __uint_as_float(rr[4 * q + 3]) + bv.w + xv.w;
                        sts4(bufX + sw_off(r, q), y);
                        if (p.has_f1) sts4(bufU + sw_off(r, q), tf32_rn4(y));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) tma_store_tile(&mp.y_out, bufX + (uint32_t)g.main_off * 128u, y0, b0);
            trace(tr, 8, ttag);
        }

        if (p.has_f1) {
            if (tid == 0) issue_mma(bufU, s.wb, true, tmem_base, g.nm, s.bar_mma);
            mbar_wait(s.bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            trace(tr, 9, ttag);
            // ---- a = relu(D + b0), zero outside the image (the zero padding of the depthwise convolution)
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
                if (r < g.rows) {
                    const RowInfo ri = row_info(g, b0, y0, r);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bv = lds4(s.vec + kC * 4 + q * 16);
                        float4 a;
                        a.x = ri.valid ? fmaxf(__uint_as_float(rr[4 * q]) + bv.x, 0.f) : 0.f;
                        a.y = ri.valid ? fmaxf(__uint_as_float(rr[4 * q + 1]) + bv.y, 0.f) : 0.f;
                        a.z = ri.valid ? fmaxf(__uint_as_float(rr[4 * q + 2]) + bv.z, 0.f) : 0.f;
                        a.w = ri.valid ? fmaxf(__uint_as_float(rr[4 * q + 3]) + bv.w, 0.f) : 0.f;
                        sts4(bufA + sw_off(r, q), a);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (tid == 0 && p.store_a) tma_store_tile(&mp.a_out, bufA + (uint32_t)g.main_off * 128u, y0, b0);
            trace(tr, 10, ttag);
            // ---- depthwise 3x3 + bias + relu over the main pixels, u into bufU (conv0 has finished reading it)
            switch (g.seg) {
                case 8: dw_fwd_stage<8>(g, s, bufA, bufU); break;
                case 4: dw_fwd_stage<4>(g, s, bufA, bufU); break;
                case 2: dw_fwd_stage<2>(g, s, bufA, bufU); break;
                default: dw_fwd_stage<1>(g, s, bufA, bufU); break;
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) tma_store_tile(&mp.u_out, bufU + (uint32_t)g.main_off * 128u, y0, b0);
            trace(tr, 11, ttag);
            if (fold1) {
                image_sums(g, s, bufU, b0, nullptr, s.img);
                __syncthreads();
                se_fwd_h(g, s, p.se, b0);
            } else {
                image_sums(g, s, bufU, b0, p.gap);
            }
            trace(tr, 12, ttag);
        }
    }
    if (fold2) se_fwd_record(g, p.se, lbid);
    teardown(tmem_base, g.nm);
    trace(tr, 13, ttag);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) mbv3_bwd_kernel(const __grid_constant__ BwdBatch bt) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.cta_begin[lvl + 1]) ++lvl;
    const BwdParams& p = bt.p[lvl];
    const BwdMaps& mp = bt.mp[lvl];
    const int lbid = (int)blockIdx.x - bt.cta_begin[lvl], lgrid = bt.cta_begin[lvl + 1] - bt.cta_begin[lvl];
    const Geom& g = p.g;
    const Smem s = carve(smem_raw, g.nm, g.nb);
    const uint32_t bufD = s.buf[0], bufU = s.buf[1], bufA = s.buf[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long* const tr = p.trace;
    const int ttag = 10 + p.has_b2 * 2 + p.has_b1;
    trace(tr, 0, ttag);
    if (p.has_b2 && p.se.fold_b2) prefetch_matrix(p.se.w0);
    if (p.has_b1 && p.se.fold_b1) prefetch_matrix(p.se.w1_prev);
    if (tid == 0) {
        prefetch_map(&mp.dy_in);
        if (p.has_b2) { prefetch_map(&mp.u_in); prefetch_map(&mp.a_in); prefetch_map(&mp.da_out); prefetch_map(&mp.dx_out); }
        if (p.has_b1) prefetch_map(&mp.up_in);
    }
    if (p.pdl) pdl_launch();
    const uint32_t tmem_base = setup(s, g.nm, true);
    const uint32_t tmem_res = tmem_base + (g.nm <= 2 ? 64u : 128u);      // second accumulator region: residual + conv0^T
    if (!p.pdl) pdl_sync();
    trace(tr, 1, ttag);
    const uint32_t tile_bytes = (uint32_t)g.rows * 128u;
    auto issue_loads = [&](int tile) {
        int b0, y0;
        tile_origin(g, tile, b0, y0);
        if (p.has_b2) {
            tma::mbar_expect_tx(s.bar_ld, 3 * tile_bytes);
            tma_load_tile(bufD, &mp.dy_in, s.bar_ld, y0 - g.halo, b0);
            tma_load_tile(bufU, &mp.u_in, s.bar_ld, y0 - g.halo, b0);
            tma_load_tile(bufA, &mp.a_in, s.bar_ld, y0 - g.halo, b0);
        } else {
            // first launch of a backward chain: the gradient tile and u of the block whose gate gradient is wanted
            tma::mbar_expect_tx(s.bar_ld, 2 * tile_bytes);
            tma_load_tile(bufD, &mp.dy_in, s.bar_ld, y0 - g.halo, b0);
            tma_load_tile(bufA, &mp.up_in, s.bar_ld, y0 - g.halo, b0);
        }
    };
    if (!p.pdl && tid == 0 && lbid < g.tiles) issue_loads(lbid);

    // this thread's channel pair of the depthwise gradients (work items of the transposed depthwise stage keep it fixed)
    const int cp = tid & 15;
    float2 dwd[9], dbd = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 9; ++k) dwd[k] = make_float2(0.f, 0.f);
    {
        // every global load of the prologue before the first shared store (one round trip)
        float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f), r0 = r2, rp = r2;
        float d0 = 0.f, d1 = 0.f;
        if (p.has_b2) { r2 = w_piece(p.w2); r0 = w_piece(p.w0); d0 = __ldg(p.wd + tid); if (tid < kC) d1 = __ldg(p.wd + kThreads + tid); }
        if (p.has_b1) rp = w_piece(p.w2p);
        if (p.has_b2) {
            put_w_dgrad(s.wa, r2); put_w_dgrad(s.wb, r0);
            sts1(s.dww + tid * 4, d0);
            if (tid < kC) sts1(s.dww + (kThreads + tid) * 4, d1);
        }
        if (p.has_b1) put_w_dgrad(s.wc, rp);
    }
    fence_proxy_async();
    if (p.pdl) {
        pdl_wait();
        if (tid == 0 && lbid < g.tiles) issue_loads(lbid);
    }
    __syncthreads();
    const bool fold2 = p.has_b2 && p.se.fold_b2, fold1 = p.has_b1 && p.se.fold_b1;
    trace(tr, 2, ttag);
    if (fold2) se_bwd_stats(g, s, p.se);               // BatchNorm-backward sums of the whole batch, while the first tile lands
    trace(tr, 3, ttag);

    uint32_t ph_ld = 0, ph_mma = 0;
    bool first = true;
    for (int tile = lbid; tile < g.tiles; tile += lgrid) {
        int b0, y0;
        tile_origin(g, tile, b0, y0);
        if (!first) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            if (tid == 0) issue_loads(tile);
        }
        first = false;
        if (fold2) se_bwd_dgap(g, s, p.se, b0);        // dgap of this tile's images (the tile is still in flight)
        if (p.has_b2) {
            // gate (and dgap, unless the folded gate just produced it) of the tile's images: shared copies for the epilogue
            for (int i = tid; i < g.nb * kC; i += kThreads) {
                const int b = min(b0 + (i >> 5), g.B - 1);
                sts1(s.img + (uint32_t)i * 4u, __ldg(p.gate + (long long)b * kC + (i & 31)));
                if (!fold2) sts1(s.img + (uint32_t)(g.nb * kC + i) * 4u, __ldg(p.dgap + (long long)b * kC + (i & 31)));
            }
        }
        trace(tr, 4, ttag);
        mbar_wait(s.bar_ld, ph_ld);
        ph_ld ^= 1u;
        trace(tr, 5, ttag);
        // ---- the landed gradient tile: its fp32 values go into the second accumulator region (the residual term of
        //      dx = da W0^T + dy is then added by the tensor core itself), its TF32 rounding stays in place as the A operand
        if (p.has_b2) {
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = lds4(bufD + sw_off(r, q));
                    rr[4 * q] = __float_as_uint(v.x); rr[4 * q + 1] = __float_as_uint(v.y);
                    rr[4 * q + 2] = __float_as_uint(v.z); rr[4 * q + 3] = __float_as_uint(v.w);
                    sts4(bufD + sw_off(r, q), tf32_rn4(v));
                }
                tmem_st32(tmem_res + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
            }
            tc_fence_before();
        } else {
            round_tile(g, bufD, nullptr, b0);
        }
        fence_proxy_async();
        __syncthreads();
        trace(tr, 6, ttag);

        if (p.has_b2) {
            // ---- dv = dy W2^T  ->  d_pre = (dv * gate + dgap) * (u > 0), zero outside the image; in place over u
            if (tid == 0) issue_mma(bufD, s.wa, false, tmem_base, g.nm, s.bar_mma);
            mbar_wait(s.bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            trace(tr, 7, ttag);
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
                if (r < g.rows) {
                    const RowInfo ri = row_info(g, b0, y0, r);
                    const int bi = g.halo ? 0 : (r >> g.lgPpi);
                    const uint32_t gs_ = s.img + (uint32_t)(bi * kC) * 4u, ds_ = s.img + (uint32_t)((g.nb + bi) * kC) * 4u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 gt = lds4(gs_ + q * 16);
                        const float4 dg = lds4(ds_ + q * 16);
                        const float4 uv = lds4(bufU + sw_off(r, q));
                        float4 d;
                        d.x = (ri.valid && uv.x > 0.f) ? fmaf(gt.x, __uint_as_float(rr[4 * q]), dg.x) : 0.f;
                        d.y = (ri.valid && uv.y > 0.f) ? fmaf(gt.y, __uint_as_float(rr[4 * q + 1]), dg.y) : 0.f;
                        d.z = (ri.valid && uv.z > 0.f) ? fmaf(gt.z, __uint_as_float(rr[4 * q + 2]), dg.z) : 0.f;
                        d.w = (ri.valid && uv.w > 0.f) ? fmaf(gt.w, __uint_as_float(rr[4 * q + 3]), dg.w) : 0.f;
                        sts4(bufU + sw_off(r, q), d);
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
            trace(tr, 8, ttag);
            // ---- da = dw^T(d_pre) * (a > 0) over the main pixels: in place over a (fp32, staging of the TMA store) and as
            //      TF32 into bufD (the A operand of conv0's dgrad); depthwise weight / bias gradients in registers
            switch (g.seg) {
                case 8: dw_bwd_stage<8>(g, s, bufU, bufA, bufD, cp, dwd, dbd); break;
                case 4: dw_bwd_stage<4>(g, s, bufU, bufA, bufD, cp, dwd, dbd); break;
                case 2: dw_bwd_stage<2>(g, s, bufU, bufA, bufD, cp, dwd, dbd); break;
                default: dw_bwd_stage<1>(g, s, bufU, bufA, bufD, cp, dwd, dbd); break;
            }
            fence_proxy_async();
            __syncthreads();
            trace(tr, 9, ttag);
            if (tid == 0) {
                tma_store_tile(&mp.da_out, bufA + (uint32_t)g.main_off * 128u, y0, b0);
                issue_mma(bufD, s.wb, false, tmem_res, g.nm, s.bar_mma, true);        // += on top of the residual tile
                if (p.has_b1) {
                    // u of the next block in the chain follows da through the same buffer once the store has read it
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    tma::mbar_expect_tx(s.bar_ld, tile_bytes);
                    tma_load_tile(bufA, &mp.up_in, s.bar_ld, y0 - g.halo, b0);
                }
            }
            mbar_wait(s.bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            trace(tr, 10, ttag);
            // ---- dx = da W0^T + dy, complete in the accumulator (main pixels): fp32 into bufU (d_pre is dead), TF32 into
            //      bufD for the next product
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
                tmem_ld32(tmem_res + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
                const RowInfo ri = row_info(g, b0, y0, r);
                if (ri.main && ri.valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 d = make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]),
                                                     __uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3]));
                        sts4(bufU + sw_off(r, q), d);
                        if (p.has_b1) sts4(bufD + sw_off(r, q), tf32_rn4(d));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) tma_store_tile(&mp.dx_out, bufU + (uint32_t)g.main_off * 128u, y0, b0);
            trace(tr, 11, ttag);
        }

        if (p.has_b1) {
            // ---- dgate[b][c] += sum over the image of (dy W2^T) * u       (dy = the dx just computed, or the landed tile)
            if (tid == 0) issue_mma(bufD, s.wc, false, tmem_base, g.nm, s.bar_mma);
            if (p.has_b2) {                      // u of this block was sent for after the da store (see above)
                mbar_wait(s.bar_ld, ph_ld);
                ph_ld ^= 1u;
            }
            mbar_wait(s.bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            trace(tr, 12, ttag);
            for (int blk = warp >> 2; blk < g.nm; blk += 2) {
                const int r = blk * 128 + (warp & 3) * 32 + lane;
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(blk * kC), rr);
                const RowInfo ri = row_info(g, b0, y0, r);
                if (ri.main) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        // (rows of images past the batch were zero-filled by the TMA load: their products vanish)
                        const float4 uv = lds4(bufA + sw_off(r, q));
                        float4 d;
                        d.x = __uint_as_float(rr[4 * q]) * uv.x;
                        d.y = __uint_as_float(rr[4 * q + 1]) * uv.y;
                        d.z = __uint_as_float(rr[4 * q + 2]) * uv.z;
                        d.w = __uint_as_float(rr[4 * q + 3]) * uv.w;
                        sts4(bufD + sw_off(r, q), d);
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
            trace(tr, 13, ttag);
            if (fold1) {
                image_sums(g, s, bufD, b0, nullptr, s.img);
                __syncthreads();
                se_bwd_ds(g, s, p.se, p.dgate, b0);
            } else {
                image_sums(g, s, bufD, b0, p.dgate);
            }
            trace(tr, 14, ttag);
        }
    }
    teardown(tmem_base, g.nm, true);
    trace(tr, 15, ttag);
    if (p.has_b2 && p.dwd) {
        // depthwise weight / bias gradients: threads -> CTA through shared memory (every TMA store has read its buffer), one
        // atomic per (tap, channel) and CTA.  Thread t holds channel pair t & 15: 16 threads per pair.
        float* red = reinterpret_cast<float*>(s.base);          // [16 groups][10][32]
        const int grp = tid >> 4;
#pragma unroll
        for (int k = 0; k < 9; ++k) *reinterpret_cast<float2*>(red + (grp * 10 + k) * kC + cp * 2) = dwd[k];
        *reinterpret_cast<float2*>(red + (grp * 10 + 9) * kC + cp * 2) = dbd;
        __syncthreads();
        for (int t = tid; t < 10 * kC; t += kThreads) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 16; ++w) acc += red[w * 10 * kC + t];
            if (t < 9 * kC) atomicAdd(p.dwd + t, acc);
            else atomicAdd(p.dbd + (t - 9 * kC), acc);
        }
    }
    trace(tr, 16, ttag);
}

// tensor map of an NHWC (B, H, W, 32) tensor with a (32, W, rows, nb) box
static bool encode_tile_map(CUtensorMap* m, const float* base, const Geom& g, int rows) {
    const unsigned long long dims[4] = {(unsigned long long)kC, (unsigned long long)g.W, (unsigned long long)g.H, (unsigned long long)g.B};
    const unsigned int box[4] = {(unsigned)kC, (unsigned)g.W, (unsigned)rows, (unsigned)g.nb};
    return tma::encode_f32(m, base, 4, dims, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int configure(const void* fn, size_t smem) {
    static bool done[2][64];
    int dev = 0;
    MVAE_CUDA(cudaGetDevice(&dev));
    const int which = fn == reinterpret_cast<const void*>(mbv3_fwd_kernel) ? 0 : 1;
    if (dev < 0 || dev >= 64 || !done[which][dev]) {
        MVAE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        if (dev >= 0 && dev < 64) done[which][dev] = true;
    }
    (void)smem;
    return MVAE_OK;
}

static int grid_for(const Geom& g, size_t smem) {
    // 228 KB of shared memory per SM, 1 KB reserved per CTA
    const int per_sm = 2 * (smem + 1024) <= 228 * 1024 ? 2 : 1;
    int cap = kNumSMs * per_sm;
    const int e = env_int("MVAE_MBV3_CTAS", 0);
    if (e > 0) cap = e;
    return g.tiles < cap ? g.tiles : cap;
}

}  // namespace mb
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_mbv3_fused_supported(int B, int H, int W, int Cin, int filters) {
    mb::Geom g;
    return (Cin == mb::kC && filters == mb::kC && mb::make_geom(B, H, W, g)) ? 1 : 0;
}

// fills one problem of a batch from the C-ABI arguments
static int build_fwd(const mvae_mbv3_fwd_args* a, mb::FwdMaps& mp, mb::FwdParams& p) {
    MVAE_REQUIRE(a, "mbv3_fused_fwd: null arguments");
    mb::Geom g;
    if (a->C != mb::kC || !mb::make_geom(a->B, a->H, a->W, g)) {
        set_error("mbv3_fused_fwd: unsupported shape B%d %dx%dx%d", a->B, a->H, a->W, a->C);
        return MVAE_ERR_UNSUPPORTED;
    }
    const bool f2 = a->w2 != nullptr, f1 = a->w0 != nullptr;
    MVAE_REQUIRE(f1 || f2, "mbv3_fused_fwd: neither phase given");
    if (f2) MVAE_REQUIRE(a->u_prev && a->x_prev && (a->gate_prev || a->se_w1_prev) && a->b2 && a->y, "mbv3_fused_fwd: F2 operands missing");
    if (f1) MVAE_REQUIRE(a->b0 && a->wd && a->bd && a->u && (a->gap_sum || a->se_w0) && (f2 || a->x), "mbv3_fused_fwd: F1 operands missing");
    memset(&mp, 0, sizeof(mp));
    p.g = g; p.gate = a->gate_prev; p.w2 = a->w2; p.b2 = a->b2; p.w0 = a->w0; p.b0 = a->b0; p.wd = a->wd; p.bd = a->bd;
    p.gap = a->gap_sum; p.has_f2 = f2; p.has_f1 = f1; p.store_a = (f1 && a->a) ? 1 : 0;
    p.trace = g_trace;
    p.pdl = pdl_chain_enabled() ? 1 : 0;
    memset(&p.se, 0, sizeof(p.se));
    if (f1 && a->se_w0) {
        MVAE_REQUIRE(!g.halo && a->se_b0 && a->se_ws, "mbv3_fused_fwd: folded gate (F1) needs whole-image tiles, se_b0, se_ws");
        MVAE_REQUIRE(a->se_stat, "mbv3_fused_fwd: folded gate (F1) needs se_stat");
        p.se.fold_f1 = 1; p.se.w0 = a->se_w0; p.se.b0 = a->se_b0; p.se.ws = a->se_ws; p.se.inv_hw = 1.f / (float)(a->H * a->W);
        p.se.stat = a->se_stat;
    }
    if (f2 && a->se_w1_prev) {
        MVAE_REQUIRE(!g.halo && a->se_gamma_prev && a->se_beta_prev && a->se_b1_prev && a->se_mm_prev && a->se_mv_prev &&
                     a->se_ws_prev && a->gate_out_prev, "mbv3_fused_fwd: folded gate (F2) operands missing");
        p.se.fold_f2 = 1; p.se.gamma = a->se_gamma_prev; p.se.beta = a->se_beta_prev; p.se.w1 = a->se_w1_prev;
        p.se.b1 = a->se_b1_prev; p.se.mm = a->se_mm_prev; p.se.mv = a->se_mv_prev; p.se.ws_prev = a->se_ws_prev;
        p.se.gate_out = a->gate_out_prev; p.se.eps = a->bn_eps; p.se.momentum = a->bn_momentum; p.se.training = a->training;
        p.se.stat_prev = a->se_stat_prev;
        MVAE_REQUIRE(!a->training || a->se_stat_prev, "mbv3_fused_fwd: folded gate (F2) needs se_stat_prev when training");
    }
    bool ok = true;
    if (f2) {
        ok = ok && mb::encode_tile_map(&mp.u_in, a->u_prev, g, g.TH) && mb::encode_tile_map(&mp.x_in, a->x_prev, g, g.TH) &&
             mb::encode_tile_map(&mp.y_out, a->y, g, g.R);
    } else {
        ok = ok && mb::encode_tile_map(&mp.x_in, a->x, g, g.TH);
    }
    if (f1) {
        ok = ok && mb::encode_tile_map(&mp.u_out, a->u, g, g.R);
        if (p.store_a) ok = ok && mb::encode_tile_map(&mp.a_out, a->a, g, g.R);
    }
    if (!ok) { set_error("mbv3_fused_fwd: tensor map encoding failed"); return MVAE_ERR_UNSUPPORTED; }
    return MVAE_OK;
}

static int build_bwd(const mvae_mbv3_bwd_args* a, mb::BwdMaps& mp, mb::BwdParams& p) {
    MVAE_REQUIRE(a, "mbv3_fused_bwd: null arguments");
    mb::Geom g;
    if (a->C != mb::kC || !mb::make_geom(a->B, a->H, a->W, g)) {
        set_error("mbv3_fused_bwd: unsupported shape B%d %dx%dx%d", a->B, a->H, a->W, a->C);
        return MVAE_ERR_UNSUPPORTED;
    }
    const bool b2 = a->w0 != nullptr, b1 = a->w2_prev != nullptr;
    MVAE_REQUIRE(b1 || b2, "mbv3_fused_bwd: neither phase given");
    MVAE_REQUIRE(a->dy, "mbv3_fused_bwd: dy missing");
    if (b2) MVAE_REQUIRE(a->u && a->a && a->gate && (a->dgap || a->se_w0) && a->w2 && a->wd && a->da && a->dx && a->dwd && a->dbd,
                         "mbv3_fused_bwd: B2 operands missing");
    if (b1) MVAE_REQUIRE(a->u_prev && a->dgate_prev, "mbv3_fused_bwd: B1 operands missing");
    memset(&mp, 0, sizeof(mp));
    p.g = g; p.dy = a->dy; p.gate = a->gate; p.dgap = a->dgap; p.w2 = a->w2; p.wd = a->wd; p.w0 = a->w0; p.dwd = a->dwd;
    p.dbd = a->dbd;
    if (env_int("MVAE_DIAG_SKIP_DWD", 0)) p.dwd = nullptr;      // diagnostic only: what the CTA-level reduction + atomics cost
    p.w2p = a->w2_prev; p.up = a->u_prev; p.dgate = a->dgate_prev; p.has_b2 = b2; p.has_b1 = b1;
    p.trace = g_trace;
    p.pdl = pdl_chain_enabled() ? 1 : 0;
    memset(&p.se, 0, sizeof(p.se));
    if (b1 && a->se_w1_prev) {
        MVAE_REQUIRE(!g.halo && a->se_ws_prev, "mbv3_fused_bwd: folded gate (B1) needs whole-image tiles and se_ws_prev");
        MVAE_REQUIRE(a->se_bstat_prev, "mbv3_fused_bwd: folded gate (B1) needs se_bstat_prev");
        p.se.fold_b1 = 1; p.se.w1_prev = a->se_w1_prev; p.se.ws_prev = a->se_ws_prev; p.se.bstat_prev = a->se_bstat_prev;
    }
    if (b2 && a->se_w0) {
        MVAE_REQUIRE(!g.halo && a->se_gamma && a->se_ws, "mbv3_fused_bwd: folded gate (B2) operands missing");
        MVAE_REQUIRE(a->se_bstat, "mbv3_fused_bwd: folded gate (B2) needs se_bstat");
        p.se.fold_b2 = 1; p.se.w0 = a->se_w0; p.se.gamma = a->se_gamma; p.se.ws = a->se_ws; p.se.bstat = a->se_bstat;
        p.se.inv_hw = 1.f / (float)(a->H * a->W);
    }
    bool ok = mb::encode_tile_map(&mp.dy_in, a->dy, g, g.TH);
    if (b1) ok = ok && mb::encode_tile_map(&mp.up_in, a->u_prev, g, g.TH);
    if (b2) {
        ok = ok && mb::encode_tile_map(&mp.u_in, a->u, g, g.TH) && mb::encode_tile_map(&mp.a_in, a->a, g, g.TH) &&
             mb::encode_tile_map(&mp.da_out, a->da, g, g.R) && mb::encode_tile_map(&mp.dx_out, a->dx, g, g.R);
    }
    if (!ok) { set_error("mbv3_fused_bwd: tensor map encoding failed"); return MVAE_ERR_UNSUPPORTED; }
    return MVAE_OK;
}

extern "C" int mvae_mbv3_fused_fwd_batched(int n, const mvae_mbv3_fwd_args* a, mvae_stream_t stream) {
    MVAE_REQUIRE(n >= 1 && n <= mb::kMaxBatch && a, "mbv3_fused_fwd_batched: 1 <= n <= %d problems", mb::kMaxBatch);
    static mb::FwdBatch bt;                       // (3.5 KB... built per call; static: not on the caller's stack)
    size_t smem = 0;
    bool pdl = false;
    for (int l = 0; l < n; ++l) {
        if (int e = build_fwd(a + l, bt.mp[l], bt.p[l])) return e;
        const size_t sl = mb::smem_bytes(bt.p[l].g.nm, bt.p[l].g.nb);
        if (sl > smem) smem = sl;
        pdl = bt.p[l].pdl != 0;
    }
    bt.n = n; bt.cta_begin[0] = 0;
    for (int l = 0; l < n; ++l) bt.cta_begin[l + 1] = bt.cta_begin[l] + mb::grid_for(bt.p[l].g, smem);
    if (int e = mb::configure(reinterpret_cast<const void*>(mb::mbv3_fwd_kernel), smem)) return e;
    MVAE_CUDA(launch_pdl_ex(pdl, mb::mbv3_fwd_kernel, dim3(bt.cta_begin[n]), dim3(mb::kThreads), smem, as_stream(stream), bt));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

extern "C" int mvae_mbv3_fused_bwd_batched(int n, const mvae_mbv3_bwd_args* a, mvae_stream_t stream) {
    MVAE_REQUIRE(n >= 1 && n <= mb::kMaxBatch && a, "mbv3_fused_bwd_batched: 1 <= n <= %d problems", mb::kMaxBatch);
    static mb::BwdBatch bt;
    size_t smem = 0;
    bool pdl = false;
    for (int l = 0; l < n; ++l) {
        if (int e = build_bwd(a + l, bt.mp[l], bt.p[l])) return e;
        const size_t sl = mb::smem_bytes(bt.p[l].g.nm, bt.p[l].g.nb);
        if (sl > smem) smem = sl;
        pdl = bt.p[l].pdl != 0;
    }
    bt.n = n; bt.cta_begin[0] = 0;
    for (int l = 0; l < n; ++l) bt.cta_begin[l + 1] = bt.cta_begin[l] + mb::grid_for(bt.p[l].g, smem);
    if (int e = mb::configure(reinterpret_cast<const void*>(mb::mbv3_bwd_kernel), smem)) return e;
    MVAE_CUDA(launch_pdl_ex(pdl, mb::mbv3_bwd_kernel, dim3(bt.cta_begin[n]), dim3(mb::kThreads), smem, as_stream(stream), bt));
    MVAE_LAUNCH_CHECK();
    ++g_tc_launches;
    return MVAE_OK;
}

extern "C" int mvae_mbv3_fused_fwd(const mvae_mbv3_fwd_args* a, mvae_stream_t stream) {
    return mvae_mbv3_fused_fwd_batched(1, a, stream);
}

extern "C" int mvae_mbv3_fused_bwd(const mvae_mbv3_bwd_args* a, mvae_stream_t stream) {
    return mvae_mbv3_fused_bwd_batched(1, a, stream);
}
