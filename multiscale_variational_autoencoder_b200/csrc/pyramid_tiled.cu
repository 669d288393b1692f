// Tiled, two-levels-per-launch pyramid kernels with shared-memory halo staging (the HBM-roofline path):
//   split_pair   : x_i -> band_i, band_{i+1} (or x_{i+1} itself when i+1 is the last level), x_{i+2}
//   merge_pair   : r_{k+2}, y_{k+1}, y_k -> r_k                       (bilinear x2 + add, twice)
//   adjoint_pair : d_k -> d_{k+1}, d_{k+2}                            (adjoint of bilinear x2, twice)
// A CTA owns a TH x TW pixel tile of the finer level and stages it (plus a 3-row / 4-pixel halo) in shared memory with
// 16-byte loads; the intermediate level lives only in shared memory, so DRAM sees each level-i element once.
// Reference: multiscale_vae.py:129-160,204-224,292-315 (see pyramid.cu for the per-level restatement these replace).
// Rows are handled as flat runs of W*C floats (NHWC with C = 3 has no per-pixel alignment): the horizontal neighbours of
// a float are C floats away.
#include "common.cuh"
#include "tma.cuh"

namespace mvae {
namespace pyr {

constexpr int kThreads = 256;

struct Taps9 { float t[9]; };

__device__ __forceinline__ void up2_idx(int y, int n_coarse, int& i0, int& i1, float& w1) {
    const int k = y >> 1;
    if (y & 1) { i0 = k; i1 = min(k + 1, n_coarse - 1); w1 = 0.25f; }
    else       { i0 = max(k - 1, 0); i1 = k; w1 = 0.75f; }
}

// --------------------------------------------------------------------------------------------------------------------
// split: persistent CTAs, the level-i tile (+halo) arrives by TMA (3-D box over (W*C, H, B), out-of-image elements read as
// zero) into a 2-stage ring, so the next tile is in flight while the current one is filtered.
//
// The zero fill happens on the RAW image, the reference pads the NORMALISED one (v*na + nb) with zeros.  By linearity
//     sum_k t_k * [inside_k] * (na*v_k + nb) = na * sum_k t_k v_k(zero-filled)  +  nb * T(y, x),
// T(y,x) = sum of the taps that fall inside the image = one of 9 constants (top/middle/bottom x left/middle/right).
// --------------------------------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 256;
constexpr int kRowsPerItem = 4;

struct SplitParams {
    float taps[9];
    float nbT[9];           // nb * T for [row case][col case]; all zero when the level is not affine
    float na, nb;
    int h, w, B;
    int tiles_x, tiles_y, ntiles;
    int filter_second;
};

template <int C, int TH, int TW>
struct SplitCfg {
    static constexpr int R0 = TH + 6;                 // staged rows of level i (3-row halo)
    static constexpr int RS0 = (TW + 8) * C;          // floats per staged row (4-pixel halo left and right)
    static constexpr int S0_FLOATS = ((R0 * RS0 + 31) / 32) * 32;     // 128-byte multiple
    static constexpr int T1H = TH / 2, T1W = TW / 2;
    static constexpr int R1 = T1H + 2;                // level i+1: 1-row halo, 4-pixel halo (16-byte aligned interior)
    static constexpr int RS1 = (T1W + 8) * C;
    static constexpr int kStages = 2;
    static constexpr int kSmemBytes = (kStages * S0_FLOATS + R1 * RS1) * 4 + 64 + 128;
};

enum { DN_NONE = 0, DN_SMEM = 1, DN_GLOBAL = 2 };

// 3x3 filter of a strip of ROWS rows x 4 pixels (4*C consecutive floats) of a staged level.
//   S      : row ABOVE the first output row, 4 floats left of the outputs; row stride RS
//   gout   : global band pointer at (first output row, first output float); row stride gstride floats
//   dn_ptr : destination of the filtered value at even rows / even pixels (= the next level), see DN_*; first row and
//            first pixel of the strip are even
//   AFFINE : band = (na*centre + nb) - (na*F + nbT), next level = na*F + nbT;  y0 = image row of the first output row
template <int C, int RS, int ROWS, int DN, bool AFFINE>
__device__ __forceinline__ void band_strip(const float* __restrict__ S, const SplitParams& p, float* __restrict__ gout,
                                           long long gstride, bool vec, float* __restrict__ dn_ptr, long long dn_stride,
                                           int y0, int hh, bool xfirst, bool xlast) {
    constexpr int U = 4 * C, WN = U + 8;
    float win[3][WN];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < WN / 4; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(S + j * RS + 4 * k);
            win[j][4 * k] = v.x; win[j][4 * k + 1] = v.y; win[j][4 * k + 2] = v.z; win[j][4 * k + 3] = v.w;
        }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
#pragma unroll
        for (int k = 0; k < WN / 4; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(S + (rr + 2) * RS + 4 * k);
            win[2][4 * k] = v.x; win[2][4 * k + 1] = v.y; win[2][4 * k + 2] = v.z; win[2][4 * k + 3] = v.w;
        }
        float tl = 0.f, tm = 0.f, tr = 0.f;
        if (AFFINE) {
            const int y = y0 + rr;
            const bool top = (y == 0), bot = (y == hh - 1);
            tl = top ? p.nbT[0] : (bot ? p.nbT[6] : p.nbT[3]);
            tm = top ? p.nbT[1] : (bot ? p.nbT[7] : p.nbT[4]);
            tr = top ? p.nbT[2] : (bot ? p.nbT[8] : p.nbT[5]);
        }
        float f[U], o[U];
#pragma unroll
        for (int e = 0; e < U; ++e) {
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) acc = fmaf(p.taps[ky * 3 + kx], win[ky][4 + e + (kx - 1) * C], acc);
            if (AFFINE) {
                float t = tm;
                if (e < C && xfirst) t = tl;
                if (e >= 3 * C && xlast) t = tr;
                acc = fmaf(p.na, acc, t);
                o[e] = fmaf(win[1][4 + e], p.na, p.nb) - acc;
            } else {
                o[e] = win[1][4 + e] - acc;
            }
            f[e] = acc;
        }
        float* g = gout + rr * gstride;
        if (vec) {
#pragma unroll
            for (int k = 0; k < C; ++k) *reinterpret_cast<float4*>(g + 4 * k) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < U; ++e) g[e] = o[e];
        }
        if (DN != DN_NONE && (rr & 1) == 0) {
            float* d = dn_ptr + (rr >> 1) * dn_stride;
#pragma unroll
            for (int c = 0; c < C; ++c) { d[c] = f[c]; d[C + c] = f[2 * C + c]; }     // pixels 0 and 2 of the unit
        }
#pragma unroll
        for (int k = 0; k < WN; ++k) { win[0][k] = win[1][k]; win[1][k] = win[2][k]; }
    }
}

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kSplitThreads) split_pair_kernel(const __grid_constant__ CUtensorMap map,
                                                                   float* __restrict__ band0, float* __restrict__ band1,
                                                                   float* __restrict__ down2, const SplitParams p) {
    pdl_sync();
    using K = SplitCfg<C, TH, TW>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    float* S1 = smem + K::kStages * K::S0_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(S1 + K::R1 * K::RS1 + ((K::R1 * K::RS1) & 1));
    const uint32_t bar0 = tma::smem_u32(bars);
    const int tid = threadIdx.x;
    const int h = p.h, w = p.w, wc = w * C, h1 = h >> 1, w1 = w >> 1;
    const long long w1c = (long long)w1 * C;
    constexpr uint32_t kTxBytes = K::R0 * K::RS0 * 4;

    if (tid == 0) {
        tma::prefetch_map(&map);
        for (int s = 0; s < K::kStages; ++s) tma::mbar_init(bar0 + 8 * s, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int stage) {
        const int tx = tile % p.tiles_x, t = tile / p.tiles_x;
        const int ty = t % p.tiles_y, b = t / p.tiles_y;
        tma::mbar_expect_tx(bar0 + 8 * stage, kTxBytes);
        tma::load_3d(tma::smem_u32(smem + stage * K::S0_FLOATS), &map, bar0 + 8 * stage, (tx * TW - 4) * C, ty * TH - 3, b);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < p.ntiles) issue(tile, 0);

    for (int k = 0; tile < p.ntiles; tile += gridDim.x, ++k) {
        const int stage = k & 1;
        if (tid == 0 && tile + (int)gridDim.x < p.ntiles) issue(tile + gridDim.x, stage ^ 1);
        const int txi = tile % p.tiles_x, tt = tile / p.tiles_x;
        const int tx0 = txi * TW, ty0 = (tt % p.tiles_y) * TH;
        const long long b = tt / p.tiles_y;
        const float* S0 = smem + stage * K::S0_FLOATS;
        tma::mbar_wait(bar0 + 8 * stage, (uint32_t)((k >> 1) & 1));

        // ---- band_i on the tile; the Gaussian at even rows / pixels is level i+1 (interior of S1) ----
        {
            constexpr int UW = TW / 4, NITEMS = (TH / kRowsPerItem) * UW;
            for (int item = tid; item < NITEMS; item += kSplitThreads) {
                const int strip = item / UW, u = item - strip * UW;
                const int r0 = strip * kRowsPerItem;
                const float* S = S0 + (r0 + 2) * K::RS0 + 4 * C + 4 * C * u - 4;
                float* g = band0 + (b * h + ty0 + r0) * (long long)wc + (long long)(tx0 + 4 * u) * C;
                float* dn = S1 + (r0 / 2 + 1) * K::RS1 + (4 + 2 * u) * C;
                band_strip<C, K::RS0, kRowsPerItem, DN_SMEM, true>(S, p, g, wc, true, dn, K::RS1, ty0 + r0, h,
                                                                   tx0 + 4 * u == 0, tx0 + 4 * u + 4 == w);
            }
        }
        // ---- 1-pixel ring of level i+1 around the tile (zero outside the level-(i+1) image) ----
        {
            constexpr int RING = 2 * (K::T1W + 2) + 2 * K::T1H;
            for (int i = tid; i < RING * C; i += kSplitThreads) {
                const int pix = i / C, c = i - pix * C;
                int yl, xc;                                        // S1 row, S1 pixel column
                if (pix < K::T1W + 2) { yl = 0; xc = 3 + pix; }
                else if (pix < 2 * (K::T1W + 2)) { yl = K::T1H + 1; xc = 3 + pix - (K::T1W + 2); }
                else { const int q = pix - 2 * (K::T1W + 2); yl = 1 + (q >> 1); xc = (q & 1) ? K::T1W + 4 : 3; }
                const int Y = (ty0 >> 1) - 1 + yl, X = (tx0 >> 1) - 4 + xc;
                float f = 0.f;
                if (Y >= 0 && Y < h1 && X >= 0 && X < w1) {
                    const float* q = S0 + (2 * yl) * K::RS0 + (2 * xc - 5) * C + c;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) f = fmaf(p.taps[ky * 3 + kx], q[ky * K::RS0 + kx * C], f);
                    // fine (2Y, 2X) is never the last row / column
                    const float t = (Y == 0) ? (X == 0 ? p.nbT[0] : p.nbT[1]) : (X == 0 ? p.nbT[3] : p.nbT[4]);
                    f = fmaf(p.na, f, t);
                }
                S1[yl * K::RS1 + xc * C + c] = f;
            }
        }
        __syncthreads();

        // ---- level i+1 on its tile: band_{i+1} (or x_{i+1} itself when it is the last level) and x_{i+2} ----
        {
            float* o1 = band1 + ((b * h1 + (ty0 >> 1)) * (long long)w1 + (tx0 >> 1)) * C;
            if (!p.filter_second) {
                for (int i = tid; i < K::T1H * K::T1W * C; i += kSplitThreads) {
                    const int yl = i / (K::T1W * C), rem = i - yl * (K::T1W * C);
                    o1[yl * w1c + rem] = S1[(yl + 1) * K::RS1 + 4 * C + rem];
                }
            } else {
                const int h2 = h1 >> 1, w2 = w1 >> 1;
                const long long w2c = (long long)w2 * C;
                float* o2 = down2 + ((b * h2 + (ty0 >> 2)) * (long long)w2 + (tx0 >> 2)) * C;
                const bool vec = ((w1c & 3) == 0) && (((tx0 >> 1) * C & 3) == 0) && ((reinterpret_cast<uintptr_t>(band1) & 15) == 0);
                static_assert(K::T1H % 2 == 0 && K::T1W % 4 == 0, "tile");
                // items of 2 rows x 4 pixels so that every thread of the CTA has work
                constexpr int UW = K::T1W / 4, NITEMS = (K::T1H / 2) * UW;
                for (int item = tid; item < NITEMS; item += kSplitThreads) {
                    const int strip = item / UW, u = item - strip * UW;
                    const int r0 = strip * 2;
                    const float* S = S1 + r0 * K::RS1 + 4 * C + 4 * C * u - 4;
                    float* g = o1 + r0 * w1c + 4 * C * u;
                    float* dn = o2 + (r0 / 2) * w2c + 2 * u * C;
                    band_strip<C, K::RS1, 2, DN_GLOBAL, false>(S, p, g, w1c, vec, dn, w2c, 0, 0, false, false);
                }
            }
        }
        __syncthreads();     // S1 and the stage read above are free for the next iteration / the TMA after next
    }
}

// --------------------------------------------------------------------------------------------------------------------
// merge (forward): r_k = y_k + up2( y_{k+1} + up2(r_{k+2}) )
// Staged levels carry a halo filled with EDGE-REPLICATED values, so the half-pixel bilinear taps are always
// (0.25, 0.75) / (0.75, 0.25) with no border cases: up2(c)[2j] = .25 c[j-1] + .75 c[j], up2(c)[2j+1] = .75 c[j] + .25 c[j+1].
// --------------------------------------------------------------------------------------------------------------------
template <int C, int TH, int TW>
struct MergeCfg {
    static constexpr int RSY = TW * C;                           // y_k tile, no halo
    // windows start 4 pixels left of the tile: TMA needs the first float of a box row on a 16-byte boundary
    static constexpr int H1 = TH / 2 + 2, W1 = TW / 2 + 8;       // level k+1 window, local origin (ty0/2 - 1, tx0/2 - 4)
    static constexpr int H2 = TH / 4 + 4, W2 = TW / 4 + 8;       // level k+2 window, local origin (ty0/4 - 2, tx0/4 - 4)
    static constexpr int RS1 = W1 * C, RS2 = W2 * C;
    static constexpr int OFF1 = ((TH * RSY + 31) / 32) * 32;
    static constexpr int OFF2 = OFF1 + ((H1 * RS1 + 31) / 32) * 32;
    static constexpr int STAGE_FLOATS = OFF2 + ((H2 * RS2 + 31) / 32) * 32;
    static constexpr int kSmemBytes = 2 * STAGE_FLOATS * 4 + 64 + 128;
    static constexpr bool kBoxOk = RSY <= 256 && RS1 <= 256 && RS2 <= 256 && ((TW / 2) * C) % 4 == 0 && ((TW / 4) * C) % 4 == 0 && (TW % 8) == 0;
};

struct MergeParams {
    int h, w, B;
    int tiles_x, tiles_y, ntiles;
    int has_r2;
};

// out-of-image slots of a staged window := the nearest in-image slot (edge replicate); origin = global coordinate of slot 0
template <int C, int RS, int NH, int NW>
__device__ __forceinline__ void replicate_fix(float* S, int oy, int ox, int hh, int ww, int tid, int nthreads) {
    for (int i = tid; i < NH * NW * C; i += nthreads) {
        const int yl = i / (NW * C), rem = i - yl * (NW * C);
        const int xl = rem / C, c = rem - xl * C;
        const int Y = oy + yl, X = ox + xl;
        if (Y < 0 || Y >= hh || X < 0 || X >= ww) {
            const int ys = min(max(Y, 0), hh - 1) - oy, xs = min(max(X, 0), ww - 1) - ox;
            if (ys >= 0 && ys < NH && xs >= 0 && xs < NW) S[yl * RS + xl * C + c] = S[ys * RS + xs * C + c];
        }
    }
}

// 2 fine rows x 4 fine pixels of bilinear x2 from 3 coarse rows x 4 coarse pixels (edge-replicated halo):
// fine rows (2j, 2j+1) <- coarse rows (j-1, j, j+1); fine pixels 4u..4u+3 <- coarse pixels 2u-1..2u+2.  sp = coarse (j-1, 2u-1).
template <int C, int RS>
__device__ __forceinline__ void up2_block(const float* __restrict__ sp, float (&e)[2][4 * C]) {
    float hx[3][4 * C];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float cv[4 * C];
#pragma unroll
        for (int q = 0; q < 4 * C; ++q) cv[q] = sp[r * RS + q];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            hx[r][0 * C + c] = cv[c] + (cv[C + c] - cv[c]) * 0.75f;                         // (.25, .75) of coarse -1, 0
            hx[r][1 * C + c] = cv[C + c] + (cv[2 * C + c] - cv[C + c]) * 0.25f;             // (.75, .25) of coarse 0, 1
            hx[r][2 * C + c] = cv[C + c] + (cv[2 * C + c] - cv[C + c]) * 0.75f;             // (.25, .75) of coarse 0, 1
            hx[r][3 * C + c] = cv[2 * C + c] + (cv[3 * C + c] - cv[2 * C + c]) * 0.25f;     // (.75, .25) of coarse 1, 2
        }
    }
#pragma unroll
    for (int q = 0; q < 4 * C; ++q) {
        e[0][q] = hx[0][q] + (hx[1][q] - hx[0][q]) * 0.75f;
        e[1][q] = hx[1][q] + (hx[2][q] - hx[1][q]) * 0.25f;
    }
}

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kThreads) merge_pair_kernel(const __grid_constant__ CUtensorMap map0,
                                                              const __grid_constant__ CUtensorMap map1,
                                                              const __grid_constant__ CUtensorMap map2,
                                                              const __grid_constant__ CUtensorMap map_out,
                                                              const MergeParams p) {
    pdl_sync();
    using K = MergeCfg<C, TH, TW>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * K::STAGE_FLOATS);
    const uint32_t bar0 = tma::smem_u32(bars);
    const int tid = threadIdx.x;
    const int h = p.h, w = p.w;
    const int h1 = h >> 1, w1 = w >> 1, h2 = h >> 2, w2 = w >> 2;
    const uint32_t tx_bytes = (TH * K::RSY + K::H1 * K::RS1 + (p.has_r2 ? K::H2 * K::RS2 : 0)) * 4;

    if (tid == 0) {
        tma::prefetch_map(&map0); tma::prefetch_map(&map1); tma::prefetch_map(&map_out);
        if (p.has_r2) tma::prefetch_map(&map2);
        tma::mbar_init(bar0, 1);
        tma::mbar_init(bar0 + 8, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int stage) {
        const int tx = tile % p.tiles_x, t = tile / p.tiles_x;
        const int ty = t % p.tiles_y, b = t / p.tiles_y;
        const uint32_t bar = bar0 + 8 * stage;
        float* st = smem + stage * K::STAGE_FLOATS;
        tma::mbar_expect_tx(bar, tx_bytes);
        tma::load_3d(tma::smem_u32(st), &map0, bar, tx * TW * C, ty * TH, b);
        tma::load_3d(tma::smem_u32(st + K::OFF1), &map1, bar, (tx * TW / 2 - 4) * C, ty * TH / 2 - 1, b);
        if (p.has_r2) tma::load_3d(tma::smem_u32(st + K::OFF2), &map2, bar, (tx * TW / 4 - 4) * C, ty * TH / 4 - 2, b);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < p.ntiles) issue(tile, 0);

    for (int k = 0; tile < p.ntiles; tile += gridDim.x, ++k) {
        const int stage = k & 1;
        if (tid == 0 && tile + (int)gridDim.x < p.ntiles) {
            tma::store_wait_read<0>();          // the TMA store of the previous tile has finished reading stage^1
            issue(tile + gridDim.x, stage ^ 1);
        }
        const int txi = tile % p.tiles_x, tt = tile / p.tiles_x;
        const int tx0 = txi * TW, ty0 = (tt % p.tiles_y) * TH;
        const int b = tt / p.tiles_y;
        float* SY = smem + stage * K::STAGE_FLOATS;
        float* S1 = SY + K::OFF1;
        float* S2 = SY + K::OFF2;
        const int o1y = (ty0 >> 1) - 1, o1x = (tx0 >> 1) - 4, o2y = (ty0 >> 2) - 2, o2x = (tx0 >> 2) - 4;
        const bool edge = tx0 == 0 || ty0 == 0 || tx0 + TW == w || ty0 + TH == h;
        tma::mbar_wait(bar0 + 8 * stage, (uint32_t)((k >> 1) & 1));

        if (p.has_r2) {
            if (edge) {
                replicate_fix<C, K::RS2, K::H2, K::W2>(S2, o2y, o2x, h2, w2, tid, kThreads);
                __syncthreads();
            }
            // r_{k+1} = y_{k+1} + up2(r_{k+2}) in place: items of 2 rows x 4 pixels aligned to even level-(k+1) coordinates
            // (window rows 2jp-1, 2jp; window pixel columns 4u..4u+3).  Slots outside the image get values too; the
            // replicate pass below overwrites the ones that are read.
            constexpr int NP = TH / 4 + 2, NU = K::W1 / 4;
            for (int item = tid; item < NP * NU; item += kThreads) {
                const int jp = item / NU, u = item - jp * NU;
                float e[2][4 * C];
                up2_block<C, K::RS2>(S2 + jp * K::RS2 + (2 * u + 1) * C, e);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int yl = 2 * jp - 1 + rr;
                    if (yl >= 0 && yl < K::H1) {
                        float4* d = reinterpret_cast<float4*>(S1 + yl * K::RS1 + 4 * u * C);
#pragma unroll
                        for (int q = 0; q < C; ++q) {
                            float4 v = d[q];
                            v.x += e[rr][4 * q]; v.y += e[rr][4 * q + 1]; v.z += e[rr][4 * q + 2]; v.w += e[rr][4 * q + 3];
                            d[q] = v;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (edge) {
            replicate_fix<C, K::RS1, K::H1, K::W1>(S1, o1y, o1x, h1, w1, tid, kThreads);
            __syncthreads();
        }
        {
            // fine rows 2j, 2j+1 x 4 pixels per thread; coarse rows j-1..j+1 -> S1 rows j..j+2; coarse pixels 2u-1..2u+2 ->
            // S1 pixel columns 2u+3..2u+6
            constexpr int UW = TW / 4, NITEMS = (TH / 2) * UW;
            for (int item = tid; item < NITEMS; item += kThreads) {
                const int ij = item / UW, iu = item - ij * UW;
                float e[2][4 * C];
                up2_block<C, K::RS1>(S1 + ij * K::RS1 + (2 * iu + 3) * C, e);
                float* yio = SY + (2 * ij) * K::RSY + 4 * iu * C;          // r_k = y_k + up2(...) in place, stored by TMA below
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
                    for (int q = 0; q < C; ++q) {
                        float4* d = reinterpret_cast<float4*>(yio + rr * K::RSY + 4 * q);
                        const float4 v = *d;
                        *d = make_float4(v.x + e[rr][4 * q], v.y + e[rr][4 * q + 1], v.z + e[rr][4 * q + 2], v.w + e[rr][4 * q + 3]);
                    }
                }
            }
        }
        tma::fence_proxy_async();    // generic-proxy writes (tile, S1) before the TMA store / the next TMA refill
        __syncthreads();
        if (tid == 0) {
            tma::store_3d(&map_out, tma::smem_u32(SY), tx0 * C, ty0, b);
            tma::store_commit();
        }
    }
    if (tid == 0) tma::store_wait_all<0>();
}

// --------------------------------------------------------------------------------------------------------------------
// merge adjoint: d_{k+1} = up2^T(d_k), d_{k+2} = up2^T(d_{k+1})
// With a ONE-pixel edge-replicated halo the adjoint of the edge-clamped bilinear x2 is the plain 4x4 window
// (.25,.75,.75,.25)^2: the border pixel's extra 0.25 (both clamped taps land on it) comes from its replica.
// Halo slots further out are only ever read by slots that are themselves recomputed at clamped coordinates.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kAdjThreads = 256;
// depth of the TMA ring.  Measured at 512x512x3, batch 128 (scripts/pyr_probe.py): two stages at three CTAs per SM 158.7 us,
// three stages at two CTAs per SM 172.0 us -- the kernel is bound by the work per tile of its 24 warps per SM, not by the
// loads in flight -- so two it stays.
constexpr int kAdjStages = 2;

template <int C, int TH, int TW>
struct AdjCfg {
    static constexpr int R0 = TH + 6;
    static constexpr int RS0 = (TW + 8) * C;
    static constexpr int T1H = TH / 2, T1W = TW / 2;
    static constexpr int H1 = T1H + 2, W1 = T1W + 2;
    static constexpr int RS1 = W1 * C;
    static constexpr int kSmemFloats = R0 * RS0 + H1 * RS1;
    // register-blocked interior items: one coarse row x 4 coarse pixels, reading 4 fine rows x 10 fine pixels
    static constexpr int A0 = (3 * C) & ~3;                         // aligned start (floats) of the fine window inside a unit
    static constexpr int LEN4 = (3 * C + 10 * C - A0 + 3) / 4;      // float4s per fine row
    static constexpr int OFF = 3 * C - A0;                          // first needed float inside the loaded window
};

template <int C, int RS>
__device__ __forceinline__ float adj_window(const float* __restrict__ p) {
    const float wv[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) r = fmaf(wv[k], p[j * RS + k * C], r);
        acc = fmaf(wv[j], r, acc);
    }
    return acc;
}

struct AdjParams {
    int h, w, B;
    int tiles_x, tiles_y, ntiles;
};

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kAdjThreads) adjoint_pair_kernel(const __grid_constant__ CUtensorMap map,
                                                                   float* __restrict__ d1, float* __restrict__ d2,
                                                                   const AdjParams p) {
    pdl_sync();
    using K = AdjCfg<C, TH, TW>;
    constexpr int S0_FLOATS = ((K::R0 * K::RS0 + 31) / 32) * 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    float* S1 = smem + kAdjStages * S0_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(S1 + K::H1 * K::RS1 + ((K::H1 * K::RS1) & 1));
    const uint32_t bar0 = tma::smem_u32(bars);
    const int tid = threadIdx.x;
    const int h = p.h, w = p.w;
    const int h1 = h >> 1, w1 = w >> 1;
    const long long w1c = (long long)w1 * C;
    constexpr uint32_t kTxBytes = K::R0 * K::RS0 * 4;

    if (tid == 0) {
        tma::prefetch_map(&map);
        for (int st = 0; st < kAdjStages; ++st) tma::mbar_init(bar0 + 8 * st, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int stage) {
        const int tx = tile % p.tiles_x, t = tile / p.tiles_x;
        const int ty = t % p.tiles_y, b = t / p.tiles_y;
        tma::mbar_expect_tx(bar0 + 8 * stage, kTxBytes);
        tma::load_3d(tma::smem_u32(smem + stage * S0_FLOATS), &map, bar0 + 8 * stage, (tx * TW - 4) * C, ty * TH - 3, b);
    };
    int tile = blockIdx.x;
    if (tid == 0)
        for (int st = 0; st < kAdjStages - 1; ++st)
            if (tile + st * (int)gridDim.x < p.ntiles) issue(tile + st * (int)gridDim.x, st);

    for (int k = 0; tile < p.ntiles; tile += gridDim.x, ++k) {
        const int stage = k % kAdjStages;
        // the stage refilled here is the one the previous iteration finished with (its closing barrier is behind us)
        if (tid == 0 && tile + (kAdjStages - 1) * (int)gridDim.x < p.ntiles)
            issue(tile + (kAdjStages - 1) * (int)gridDim.x, (k + kAdjStages - 1) % kAdjStages);
        const int txi = tile % p.tiles_x, tt = tile / p.tiles_x;
        const int tx0 = txi * TW, ty0 = (tt % p.tiles_y) * TH;
        const long long b = tt / p.tiles_y;
        float* S0 = smem + stage * S0_FLOATS;
        tma::mbar_wait(bar0 + 8 * stage, (uint32_t)((k / kAdjStages) & 1));

        // ---- edge tiles: replicate the image border one pixel outwards (TMA zero-filled it) ----
        const bool eL = tx0 == 0, eR = tx0 + TW == w, eT = ty0 == 0, eB = ty0 + TH == h;
        if (eL || eR || eT || eB) {
            if (eL || eR) {
                for (int i = tid; i < K::R0 * C; i += kAdjThreads) {
                    const int r = i / C, c = i - r * C;
                    float* row = S0 + r * K::RS0;
                    if (eL) row[3 * C + c] = row[4 * C + c];
                    if (eR) row[(TW + 4) * C + c] = row[(TW + 3) * C + c];
                }
                __syncthreads();
            }
            if (eT || eB) {
                for (int i = tid; i < K::RS0; i += kAdjThreads) {
                    if (eT) S0[2 * K::RS0 + i] = S0[3 * K::RS0 + i];
                    if (eB) S0[(TH + 3) * K::RS0 + i] = S0[(TH + 2) * K::RS0 + i];
                }
            }
            tma::fence_proxy_async();        // generic-proxy writes precede the TMA that will refill this stage
            __syncthreads();
        }

        float* o1 = d1 + ((b * h1 + (ty0 >> 1)) * (long long)w1 + (tx0 >> 1)) * C;
        static_assert(K::T1W % 4 == 0, "tile");
        {
            // ---- interior of d_{k+1}: one coarse row x 4 coarse pixels per item.  A tile has fewer items than the CTA has
            //      threads (32x64: 128 of 256), and the ring of d_{k+1} around the tile -- recomputed from the staged d_k at
            //      edge-clamped coordinates, needed by the second adjoint only -- reads nothing the items write: the idle
            //      threads do the ring meanwhile (one barrier and one half-empty phase less per tile)
            constexpr int UW = K::T1W / 4, NITEMS = K::T1H * UW;
            static_assert(NITEMS < kAdjThreads, "the ring needs idle threads");
            const bool vec = ((w1c & 3) == 0) && (((tx0 >> 1) * C & 3) == 0) && ((reinterpret_cast<uintptr_t>(d1) & 15) == 0);
            if (tid >= NITEMS && d2) {
                constexpr int RING = 2 * K::W1 + 2 * K::T1H;
                for (int i = tid - NITEMS; i < RING * C; i += kAdjThreads - NITEMS) {
                    const int pix = i / C, c = i - pix * C;
                    int yl, xl;
                    if (pix < K::W1) { yl = 0; xl = pix; }
                    else if (pix < 2 * K::W1) { yl = K::H1 - 1; xl = pix - K::W1; }
                    else { const int q = pix - 2 * K::W1; yl = 1 + (q >> 1); xl = (q & 1) ? K::W1 - 1 : 0; }
                    const int Yc = min(max((ty0 >> 1) - 1 + yl, 0), h1 - 1) - ((ty0 >> 1) - 1);
                    const int Xc = min(max((tx0 >> 1) - 1 + xl, 0), w1 - 1) - ((tx0 >> 1) - 1);
                    S1[yl * K::RS1 + xl * C + c] = adj_window<C, K::RS0>(S0 + (2 * Yc) * K::RS0 + (2 * Xc + 1) * C + c);
                }
            }
            for (int item = tid; item < NITEMS; item += kAdjThreads) {
                const int yl = item / UW, u = item - yl * UW;
                const float* S = S0 + (2 * yl + 2) * K::RS0 + 8 * u * C + K::A0;
                float acc[4 * C];
#pragma unroll
                for (int q = 0; q < 4 * C; ++q) acc[q] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float row[K::LEN4 * 4];
#pragma unroll
                    for (int q = 0; q < K::LEN4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(S + j * K::RS0 + 4 * q);
                        row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
                    }
                    const float wy = (j == 0 || j == 3) ? 0.25f : 0.75f;
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float* f = row + K::OFF + 2 * m * C + c;
                            const float hs = fmaf(0.25f, f[3 * C], fmaf(0.75f, f[2 * C], fmaf(0.75f, f[C], 0.25f * f[0])));
                            acc[m * C + c] = fmaf(wy, hs, acc[m * C + c]);
                        }
                }
                float* g = o1 + yl * w1c + 4 * u * C;
                if (vec) {
#pragma unroll
                    for (int q = 0; q < C; ++q)
                        *reinterpret_cast<float4*>(g + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4 * C; ++q) g[q] = acc[q];
                }
                float* sd = S1 + (yl + 1) * K::RS1 + (1 + 4 * u) * C;
#pragma unroll
                for (int q = 0; q < 4 * C; ++q) sd[q] = acc[q];
            }
        }
        if (d2) {
            __syncthreads();             // interior (threads < NITEMS) and ring (the others) of d_{k+1} are in S1
            const int h2 = h1 >> 1, w2 = w1 >> 1;
            constexpr int T2H = TH / 4, T2W = TW / 4;
            float* o2 = d2 + ((b * h2 + (ty0 >> 2)) * (long long)w2 + (tx0 >> 2)) * C;
            for (int i = tid; i < T2H * T2W * C; i += kAdjThreads) {
                const int yl = i / (T2W * C), rem = i - yl * (T2W * C);
                const int xl = rem / C, c = rem - xl * C;
                o2[(long long)yl * w2 * C + rem] = adj_window<C, K::RS1>(S1 + (2 * yl) * K::RS1 + (2 * xl) * C + c);
            }
        }
        __syncthreads();     // stage and S1 are free again
    }
}

// --------------------------------------------------------------------------------------------------------------------
// dispatch
// --------------------------------------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int C, int TH, int TW>
static int launch_split(const float* src, float* band0, float* band1, float* down2, int B, int h, int w, const float* taps9,
                        float na, float nb, int filter_second, cudaStream_t s) {
    using K = SplitCfg<C, TH, TW>;
    if ((TW + 8) * C > 256) return MVAE_ERR_UNSUPPORTED;          // TMA box dimension limit
    CUtensorMap map;
    const unsigned long long dims[3] = {(unsigned long long)w * C, (unsigned long long)h, (unsigned long long)B};
    const unsigned int box[3] = {(unsigned)K::RS0, (unsigned)K::R0, 1u};
    if (!tma::encode_f32(&map, src, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    SplitParams p;
    // T[row case][col case] = sum of the taps inside the image (row case 0: no row above, 2: no row below; same for cols)
    for (int ry = 0; ry < 3; ++ry)
        for (int rx = 0; rx < 3; ++rx) {
            double t = 0.0;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    if ((ry == 0 && ky == 0) || (ry == 2 && ky == 2) || (rx == 0 && kx == 0) || (rx == 2 && kx == 2)) continue;
                    t += (double)taps9[ky * 3 + kx];
                }
            p.nbT[ry * 3 + rx] = (float)((double)nb * t);
        }
    for (int i = 0; i < 9; ++i) p.taps[i] = taps9[i];
    p.na = na; p.nb = nb; p.h = h; p.w = w; p.B = B;
    p.tiles_x = w / TW; p.tiles_y = h / TH; p.ntiles = p.tiles_x * p.tiles_y * B;
    p.filter_second = filter_second;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        MVAE_CUDA(cudaFuncSetAttribute(split_pair_kernel<C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmemBytes));
        MVAE_CUDA(cudaFuncSetAttribute(split_pair_kernel<C, TH, TW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        MVAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, split_pair_kernel<C, TH, TW>, kSplitThreads,
                                                                K::kSmemBytes));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int grid = kNumSMs * ctas_per_sm;
    if (grid > p.ntiles) grid = p.ntiles;
    MVAE_CUDA(launch_pdl(split_pair_kernel<C, TH, TW>, dim3(grid), dim3(kSplitThreads), K::kSmemBytes, s, map, band0, band1, down2, p));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

template <int C, int TH, int TW>
static int launch_merge(const float* y0, const float* y1, const float* r2, float* out, int B, int h, int w, cudaStream_t s) {
    using K = MergeCfg<C, TH, TW>;
    if (!K::kBoxOk) return MVAE_ERR_UNSUPPORTED;
    if (((w / 2) * C) % 4 != 0 || (r2 && ((w / 4) * C) % 4 != 0)) return MVAE_ERR_UNSUPPORTED;     // TMA row strides: 16 bytes
    if ((reinterpret_cast<uintptr_t>(y1) & 15) || (reinterpret_cast<uintptr_t>(r2) & 15)) return MVAE_ERR_UNSUPPORTED;
    CUtensorMap m0, m1, m2;
    {
        const unsigned long long dims[3] = {(unsigned long long)w * C, (unsigned long long)h, (unsigned long long)B};
        const unsigned int box[3] = {(unsigned)K::RSY, (unsigned)TH, 1u};
        if (!tma::encode_f32(&m0, y0, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    }
    {
        const unsigned long long dims[3] = {(unsigned long long)(w / 2) * C, (unsigned long long)(h / 2), (unsigned long long)B};
        const unsigned int box[3] = {(unsigned)K::RS1, (unsigned)K::H1, 1u};
        if (!tma::encode_f32(&m1, y1, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    }
    CUtensorMap mo;
    {
        const unsigned long long dims[3] = {(unsigned long long)w * C, (unsigned long long)h, (unsigned long long)B};
        const unsigned int box[3] = {(unsigned)K::RSY, (unsigned)TH, 1u};
        if (!tma::encode_f32(&mo, out, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    }
    m2 = m1;
    if (r2) {
        const unsigned long long dims[3] = {(unsigned long long)(w / 4) * C, (unsigned long long)(h / 4), (unsigned long long)B};
        const unsigned int box[3] = {(unsigned)K::RS2, (unsigned)K::H2, 1u};
        if (!tma::encode_f32(&m2, r2, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    }
    MergeParams p;
    p.h = h; p.w = w; p.B = B;
    p.tiles_x = w / TW; p.tiles_y = h / TH; p.ntiles = p.tiles_x * p.tiles_y * B;
    p.has_r2 = r2 ? 1 : 0;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        MVAE_CUDA(cudaFuncSetAttribute(merge_pair_kernel<C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmemBytes));
        MVAE_CUDA(cudaFuncSetAttribute(merge_pair_kernel<C, TH, TW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        MVAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, merge_pair_kernel<C, TH, TW>, kThreads,
                                                                K::kSmemBytes));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int grid = kNumSMs * ctas_per_sm;
    if (grid > p.ntiles) grid = p.ntiles;
    MVAE_CUDA(launch_pdl(merge_pair_kernel<C, TH, TW>, dim3(grid), dim3(kThreads), K::kSmemBytes, s, m0, m1, m2, mo, p));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

template <int C, int TH, int TW>
static int launch_adjoint(const float* d0, float* d1, float* d2, int B, int h, int w, cudaStream_t s) {
    using K = AdjCfg<C, TH, TW>;
    if ((TW + 8) * C > 256) return MVAE_ERR_UNSUPPORTED;          // TMA box dimension limit
    constexpr int S0_FLOATS = ((K::R0 * K::RS0 + 31) / 32) * 32;
    constexpr int kSmemBytes = (kAdjStages * S0_FLOATS + K::H1 * K::RS1) * 4 + 64 + 128;
    CUtensorMap map;
    const unsigned long long dims[3] = {(unsigned long long)w * C, (unsigned long long)h, (unsigned long long)B};
    const unsigned int box[3] = {(unsigned)K::RS0, (unsigned)K::R0, 1u};
    if (!tma::encode_f32(&map, d0, 3, dims, box)) return MVAE_ERR_UNSUPPORTED;
    AdjParams p;
    p.h = h; p.w = w; p.B = B;
    p.tiles_x = w / TW; p.tiles_y = h / TH; p.ntiles = p.tiles_x * p.tiles_y * B;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        MVAE_CUDA(cudaFuncSetAttribute(adjoint_pair_kernel<C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        MVAE_CUDA(cudaFuncSetAttribute(adjoint_pair_kernel<C, TH, TW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        MVAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, adjoint_pair_kernel<C, TH, TW>, kAdjThreads,
                                                                kSmemBytes));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int grid = kNumSMs * ctas_per_sm;
    if (grid > p.ntiles) grid = p.ntiles;
    MVAE_CUDA(launch_pdl(adjoint_pair_kernel<C, TH, TW>, dim3(grid), dim3(kAdjThreads), kSmemBytes, s, map, d1, d2, p));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

// tile shapes: 0 = 32x64, 1 = 32x32, 2 = 16x16, 3 = 8x8 (the finer level must tile exactly; B <= 65535)
static int pick_tile(int B, int h, int w) {
    if (B > 65535 || (h & 3) || (w & 3)) return -1;
    if (h % 32 == 0 && w % 64 == 0) return 0;
    if (h % 32 == 0 && w % 32 == 0) return 1;
    if (h % 16 == 0 && w % 16 == 0) return 2;
    if (h % 8 == 0 && w % 8 == 0) return 3;
    return -1;
}

#define MVAE_PYR_DISPATCH(FN, C, tile, ...)                                   \
    do {                                                                      \
        switch (tile) {                                                       \
            case 0: return FN<C, 32, 64>(__VA_ARGS__);                        \
            case 1: return FN<C, 32, 32>(__VA_ARGS__);                        \
            case 2: return FN<C, 16, 16>(__VA_ARGS__);                        \
            default: return FN<C, 8, 8>(__VA_ARGS__);                         \
        }                                                                     \
    } while (0)

}  // namespace pyr

// Each returns MVAE_ERR_UNSUPPORTED when the shape is not covered (the caller falls back to the per-level kernels).
int pyr_split_pair(const float* src, float* band0, float* band1, float* down2, int B, int h, int w, int C,
                   const float* taps9, float na, float nb, int filter_second, cudaStream_t s) {
    const int tile = pyr::pick_tile(B, h, w);
    if (tile < 0 || !pyr::al16(src) || !pyr::al16(band0)) return MVAE_ERR_UNSUPPORTED;
    const float* t = taps9;
    if (C == 3) MVAE_PYR_DISPATCH(pyr::launch_split, 3, tile, src, band0, band1, down2, B, h, w, t, na, nb, filter_second, s);
    if (C == 1) MVAE_PYR_DISPATCH(pyr::launch_split, 1, tile, src, band0, band1, down2, B, h, w, t, na, nb, filter_second, s);
    if (C == 4) MVAE_PYR_DISPATCH(pyr::launch_split, 4, tile, src, band0, band1, down2, B, h, w, t, na, nb, filter_second, s);
    return MVAE_ERR_UNSUPPORTED;
}

int pyr_merge_pair(const float* y0, const float* y1, const float* r2, float* out, int B, int h, int w, int C,
                   cudaStream_t s) {
    const int tile = pyr::pick_tile(B, h, w);
    if (tile < 0 || !pyr::al16(y0) || !pyr::al16(out)) return MVAE_ERR_UNSUPPORTED;
    if (C == 3) MVAE_PYR_DISPATCH(pyr::launch_merge, 3, tile, y0, y1, r2, out, B, h, w, s);
    if (C == 1) MVAE_PYR_DISPATCH(pyr::launch_merge, 1, tile, y0, y1, r2, out, B, h, w, s);
    if (C == 4) MVAE_PYR_DISPATCH(pyr::launch_merge, 4, tile, y0, y1, r2, out, B, h, w, s);
    return MVAE_ERR_UNSUPPORTED;
}

int pyr_adjoint_pair(const float* d0, float* d1, float* d2, int B, int h, int w, int C, cudaStream_t s) {
    const int tile = pyr::pick_tile(B, h, w);
    if (tile < 0 || !pyr::al16(d0)) return MVAE_ERR_UNSUPPORTED;
    if (C == 3) MVAE_PYR_DISPATCH(pyr::launch_adjoint, 3, tile, d0, d1, d2, B, h, w, s);
    if (C == 1) MVAE_PYR_DISPATCH(pyr::launch_adjoint, 1, tile, d0, d1, d2, B, h, w, s);
    if (C == 4) MVAE_PYR_DISPATCH(pyr::launch_adjoint, 4, tile, d0, d1, d2, B, h, w, s);
    return MVAE_ERR_UNSUPPORTED;
}

}  // namespace mvae
