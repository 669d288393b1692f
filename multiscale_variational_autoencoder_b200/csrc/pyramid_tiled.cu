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
// split
// --------------------------------------------------------------------------------------------------------------------
template <int C, int TH, int TW>
struct SplitCfg {
    static constexpr int R0 = TH + 6;                 // staged rows of level i
    static constexpr int RS0 = (TW + 8) * C;          // floats per staged row (4-pixel halo left and right)
    static constexpr int H1 = TH / 2 + 2, W1 = TW / 2 + 2;
    static constexpr int RS1 = W1 * C;
    static constexpr int kSmemFloats = R0 * RS0 + H1 * RS1;
    static constexpr int ROWS_PER_THREAD = 8;
};

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kThreads) split_pair_kernel(const float* __restrict__ src, float* __restrict__ band0,
                                                              float* __restrict__ band1, float* __restrict__ down2,
                                                              int h, int w, Taps9 taps, float na, float nb, int filter_second) {
    using K = SplitCfg<C, TH, TW>;
    extern __shared__ __align__(16) float smem[];
    float* S0 = smem;
    float* S1 = smem + K::R0 * K::RS0;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const long long b = blockIdx.z;
    const int wc = w * C;
    const float* img = src + b * (long long)h * wc;

    // ---- stage level i (normalised; zero outside the image == SAME zero padding in the normalised domain) ----
    {
        constexpr int Q = K::RS0 / 4;
        const int col0 = (tx0 - 4) * C;                       // global float column of staged column 0 (multiple of 4)
        for (int i = tid; i < K::R0 * Q; i += kThreads) {
            const int r = i / Q, q = i - r * Q;
            const int y = ty0 - 3 + r, gc = col0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < h && gc >= 0 && gc < wc) {
                v = __ldg(reinterpret_cast<const float4*>(img + (long long)y * wc + gc));
                v.x = fmaf(v.x, na, nb); v.y = fmaf(v.y, na, nb); v.z = fmaf(v.z, na, nb); v.w = fmaf(v.w, na, nb);
            }
            *reinterpret_cast<float4*>(S0 + r * K::RS0 + 4 * q) = v;
        }
    }
    __syncthreads();

    // ---- level i+1 (with a 1-pixel halo) = Gaussian of level i at even rows/cols; zero outside the level-(i+1) image ----
    {
        const int h1 = h >> 1, w1 = w >> 1;
        for (int i = tid; i < K::H1 * K::RS1; i += kThreads) {
            const int yl = i / K::RS1, rem = i - yl * K::RS1;
            const int xl = rem / C, c = rem - xl * C;
            const int Y = (ty0 >> 1) - 1 + yl, X = (tx0 >> 1) - 1 + xl;
            float f = 0.f;
            if (Y >= 0 && Y < h1 && X >= 0 && X < w1) {
                const float* p = S0 + (2 * yl) * K::RS0 + (2 * xl + 1) * C + c;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) f = fmaf(taps.t[ky * 3 + kx], p[ky * K::RS0 + kx * C], f);
            }
            S1[i] = f;
        }
    }

    // ---- band_i = x_i - G*x_i on the tile: each thread owns 4 consecutive floats x ROWS_PER_THREAD rows, rolling 3 rows ----
    {
        constexpr int R = K::ROWS_PER_THREAD;
        constexpr int QW = TW * C / 4;
        constexpr int NITEMS = (TH / R) * QW;
        float* out = band0 + (b * h + ty0) * (long long)wc + (long long)tx0 * C;
        for (int item = tid; item < NITEMS; item += kThreads) {
            const int strip = item / QW, q = item - strip * QW;
            const int r0 = strip * R;                         // first tile row of the strip
            const float* base = S0 + (r0 + 2) * K::RS0 + 4 * C + 4 * q - 4;   // row above, 4 floats left of the outputs
            float win[3][12];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(base + j * K::RS0 + 4 * k);
                    win[j][4 * k] = v.x; win[j][4 * k + 1] = v.y; win[j][4 * k + 2] = v.z; win[j][4 * k + 3] = v.w;
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(base + (rr + 2) * K::RS0 + 4 * k);
                    win[2][4 * k] = v.x; win[2][4 * k + 1] = v.y; win[2][4 * k + 2] = v.z; win[2][4 * k + 3] = v.w;
                }
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float f = 0.f;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) f = fmaf(taps.t[ky * 3 + kx], win[ky][4 + e + (kx - 1) * C], f);
                    o[e] = win[1][4 + e] - f;
                }
                *reinterpret_cast<float4*>(out + (long long)(r0 + rr) * wc + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
                for (int k = 0; k < 12; ++k) { win[0][k] = win[1][k]; win[1][k] = win[2][k]; }
            }
        }
    }
    __syncthreads();

    // ---- level i+1 on its tile: band_{i+1} (or x_{i+1} itself when it is the last level) and x_{i+2} ----
    {
        const int h1 = h >> 1, w1 = w >> 1;
        constexpr int T1H = TH / 2, T1W = TW / 2;
        float* o1 = band1 + ((b * h1 + (ty0 >> 1)) * (long long)w1 + (tx0 >> 1)) * C;
        for (int i = tid; i < T1H * T1W * C; i += kThreads) {
            const int yl = i / (T1W * C), rem = i - yl * (T1W * C);
            const float* p = S1 + (yl + 1) * K::RS1 + C + rem;
            const float centre = *p;
            if (!filter_second) { o1[(long long)yl * w1 * C + rem] = centre; continue; }
            float f = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) f = fmaf(taps.t[ky * 3 + kx], p[(ky - 1) * K::RS1 + (kx - 1) * C], f);
            o1[(long long)yl * w1 * C + rem] = centre - f;
            const int xl = rem / C, c = rem - xl * C;
            if (!(yl & 1) && !(xl & 1)) {
                const int h2 = h1 >> 1, w2 = w1 >> 1;
                down2[((b * h2 + (ty0 >> 2) + (yl >> 1)) * (long long)w2 + (tx0 >> 2) + (xl >> 1)) * C + c] = f;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// merge (forward): r_k = y_k + up2( y_{k+1} + up2(r_{k+2}) )
// --------------------------------------------------------------------------------------------------------------------
template <int C, int TH, int TW>
struct MergeCfg {
    static constexpr int H1 = TH / 2 + 2, W1 = TW / 2 + 2;
    static constexpr int H2 = TH / 4 + 4, W2 = TW / 4 + 4;
    static constexpr int kSmemFloats = H1 * W1 * C + H2 * W2 * C;
};

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kThreads) merge_pair_kernel(const float* __restrict__ y0, const float* __restrict__ y1,
                                                              const float* __restrict__ r2, float* __restrict__ out,
                                                              int h, int w) {
    using K = MergeCfg<C, TH, TW>;
    extern __shared__ __align__(16) float smem[];
    float* S1 = smem;                         // r_{k+1} on [A1y,B1y) x [A1x,B1x)
    float* S2 = smem + K::H1 * K::W1 * C;     // r_{k+2} on [A2y,B2y) x [A2x,B2x)
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const long long b = blockIdx.z;
    const int h1 = h >> 1, w1 = w >> 1, h2 = h >> 2, w2 = w >> 2;
    const int A1y = max((ty0 >> 1) - 1, 0), B1y = min((ty0 >> 1) + TH / 2 + 1, h1);
    const int A1x = max((tx0 >> 1) - 1, 0), B1x = min((tx0 >> 1) + TW / 2 + 1, w1);
    const int n1y = B1y - A1y, n1x = B1x - A1x;
    int A2y = 0, A2x = 0, n2x = 0;
    if (r2) {
        A2y = max((A1y >> 1) - 1, 0); A2x = max((A1x >> 1) - 1, 0);
        const int B2y = min(((B1y - 1) >> 1) + 2, h2), B2x = min(((B1x - 1) >> 1) + 2, w2);
        const int n2y = B2y - A2y;
        n2x = B2x - A2x;
        const float* g2 = r2 + b * (long long)h2 * w2 * C;
        for (int i = tid; i < n2y * n2x * C; i += kThreads) {
            const int yl = i / (n2x * C), rem = i - yl * (n2x * C);
            S2[i] = __ldg(g2 + ((long long)(A2y + yl) * w2 + A2x) * C + rem);
        }
        __syncthreads();
    }
    {
        const float* g1 = y1 + b * (long long)h1 * w1 * C;
        for (int i = tid; i < n1y * n1x * C; i += kThreads) {
            const int yl = i / (n1x * C), rem = i - yl * (n1x * C);
            const int xl = rem / C, c = rem - xl * C;
            float v = __ldg(g1 + ((long long)(A1y + yl) * w1 + A1x) * C + rem);
            if (r2) {
                int ya, yb, xa, xb; float wy, wx;
                up2_idx(A1y + yl, h2, ya, yb, wy);
                up2_idx(A1x + xl, w2, xa, xb, wx);
                const float* s = S2 + c;
                const float c00 = s[((ya - A2y) * n2x + (xa - A2x)) * C], c01 = s[((ya - A2y) * n2x + (xb - A2x)) * C];
                const float c10 = s[((yb - A2y) * n2x + (xa - A2x)) * C], c11 = s[((yb - A2y) * n2x + (xb - A2x)) * C];
                const float top = c00 + (c01 - c00) * wx, bot = c10 + (c11 - c10) * wx;
                v += top + (bot - top) * wy;
            }
            S1[i] = v;
        }
    }
    __syncthreads();
    {
        constexpr int QW = TW * C / 4;
        const int wc = w * C;
        const float* g0 = y0 + (b * h + ty0) * (long long)wc + (long long)tx0 * C;
        float* o = out + (b * h + ty0) * (long long)wc + (long long)tx0 * C;
        for (int i = tid; i < TH * QW; i += kThreads) {
            const int r = i / QW, q = i - r * QW;
            int ya, yb; float wy;
            up2_idx(ty0 + r, h1, ya, yb, wy);
            const float* sa = S1 + (ya - A1y) * n1x * C;
            const float* sb = S1 + (yb - A1y) * n1x * C;
            const float4 v = __ldg(reinterpret_cast<const float4*>(g0 + (long long)r * wc + 4 * q));
            float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = 4 * q + k, xl = j / C, c = j - xl * C;
                int xa, xb; float wx;
                up2_idx(tx0 + xl, w1, xa, xb, wx);
                const int ia = (xa - A1x) * C + c, ib = (xb - A1x) * C + c;
                const float top = sa[ia] + (sa[ib] - sa[ia]) * wx, bot = sb[ia] + (sb[ib] - sb[ia]) * wx;
                e[k] += top + (bot - top) * wy;
            }
            *reinterpret_cast<float4*>(o + (long long)r * wc + 4 * q) = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// merge adjoint: d_{k+1} = up2^T(d_k), d_{k+2} = up2^T(d_{k+1})
// --------------------------------------------------------------------------------------------------------------------
template <int C, int TH, int TW>
struct AdjCfg {
    static constexpr int R0 = TH + 6;
    static constexpr int RS0 = (TW + 8) * C;
    static constexpr int H1 = TH / 2 + 2, W1 = TW / 2 + 2;
    static constexpr int RS1 = W1 * C;
    static constexpr int kSmemFloats = R0 * RS0 + H1 * RS1;
};

__device__ __forceinline__ void adj_weights(int Y, int n_coarse, float (&wv)[4]) {
    wv[0] = 0.25f; wv[1] = 0.75f; wv[2] = 0.75f; wv[3] = 0.25f;
    if (Y == 0) { wv[0] = 0.f; wv[1] = 1.f; }
    if (Y == n_coarse - 1) { wv[3] = 0.f; wv[2] = 1.f; }
}

template <int C, int TH, int TW>
__global__ void __launch_bounds__(kThreads) adjoint_pair_kernel(const float* __restrict__ d0, float* __restrict__ d1,
                                                                float* __restrict__ d2, int h, int w) {
    using K = AdjCfg<C, TH, TW>;
    extern __shared__ __align__(16) float smem[];
    float* S0 = smem;
    float* S1 = smem + K::R0 * K::RS0;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const long long b = blockIdx.z;
    const int wc = w * C;
    const int h1 = h >> 1, w1 = w >> 1;
    {
        constexpr int Q = K::RS0 / 4;
        const float* img = d0 + b * (long long)h * wc;
        const int col0 = (tx0 - 4) * C;
        for (int i = tid; i < K::R0 * Q; i += kThreads) {
            const int r = i / Q, q = i - r * Q;
            const int y = ty0 - 3 + r, gc = col0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < h && gc >= 0 && gc < wc) v = __ldg(reinterpret_cast<const float4*>(img + (long long)y * wc + gc));
            *reinterpret_cast<float4*>(S0 + r * K::RS0 + 4 * q) = v;
        }
    }
    __syncthreads();
    {
        float* o1 = d1 + b * (long long)h1 * w1 * C;
        for (int i = tid; i < K::H1 * K::RS1; i += kThreads) {
            const int yl = i / K::RS1, rem = i - yl * K::RS1;
            const int xl = rem / C, c = rem - xl * C;
            const int Y = (ty0 >> 1) - 1 + yl, X = (tx0 >> 1) - 1 + xl;
            float acc = 0.f;
            if (Y >= 0 && Y < h1 && X >= 0 && X < w1) {
                float wy[4], wx[4];
                adj_weights(Y, h1, wy);
                adj_weights(X, w1, wx);
                const float* p = S0 + (2 * yl) * K::RS0 + (2 * xl + 1) * C + c;      // fine (2Y-1, 2X-1)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float r = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) r = fmaf(wx[k], p[j * K::RS0 + k * C], r);
                    acc = fmaf(wy[j], r, acc);
                }
                if (yl >= 1 && yl <= TH / 2 && xl >= 1 && xl <= TW / 2) o1[((long long)Y * w1 + X) * C + c] = acc;
            }
            S1[i] = acc;
        }
    }
    if (!d2) return;
    __syncthreads();
    {
        const int h2 = h1 >> 1, w2 = w1 >> 1;
        constexpr int T2H = TH / 4, T2W = TW / 4;
        float* o2 = d2 + b * (long long)h2 * w2 * C;
        for (int i = tid; i < T2H * T2W * C; i += kThreads) {
            const int yl = i / (T2W * C), rem = i - yl * (T2W * C);
            const int xl = rem / C, c = rem - xl * C;
            const int Y = (ty0 >> 2) + yl, X = (tx0 >> 2) + xl;
            float wy[4], wx[4];
            adj_weights(Y, h2, wy);
            adj_weights(X, w2, wx);
            const float* p = S1 + (2 * yl) * K::RS1 + (2 * xl) * C + c;               // level-(k+1) (2Y-1, 2X-1), local origin -1
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float r = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) r = fmaf(wx[k], p[j * K::RS1 + k * C], r);
                acc = fmaf(wy[j], r, acc);
            }
            o2[((long long)Y * w2 + X) * C + c] = acc;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// dispatch
// --------------------------------------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int C, int TH, int TW>
static int launch_split(const float* src, float* band0, float* band1, float* down2, int B, int h, int w, const Taps9& t,
                        float na, float nb, int filter_second, cudaStream_t s) {
    using K = SplitCfg<C, TH, TW>;
    static bool configured = false;
    const int smem = K::kSmemFloats * 4;
    if (!configured) {
        MVAE_CUDA(cudaFuncSetAttribute(split_pair_kernel<C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(w / TW, h / TH, B);
    split_pair_kernel<C, TH, TW><<<grid, kThreads, smem, s>>>(src, band0, band1, down2, h, w, t, na, nb, filter_second);
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

template <int C, int TH, int TW>
static int launch_merge(const float* y0, const float* y1, const float* r2, float* out, int B, int h, int w, cudaStream_t s) {
    using K = MergeCfg<C, TH, TW>;
    dim3 grid(w / TW, h / TH, B);
    merge_pair_kernel<C, TH, TW><<<grid, kThreads, K::kSmemFloats * 4, s>>>(y0, y1, r2, out, h, w);
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

template <int C, int TH, int TW>
static int launch_adjoint(const float* d0, float* d1, float* d2, int B, int h, int w, cudaStream_t s) {
    using K = AdjCfg<C, TH, TW>;
    static bool configured = false;
    const int smem = K::kSmemFloats * 4;
    if (!configured) {
        MVAE_CUDA(cudaFuncSetAttribute(adjoint_pair_kernel<C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(w / TW, h / TH, B);
    adjoint_pair_kernel<C, TH, TW><<<grid, kThreads, smem, s>>>(d0, d1, d2, h, w);
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
    pyr::Taps9 t;
    for (int i = 0; i < 9; ++i) t.t[i] = taps9[i];
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
