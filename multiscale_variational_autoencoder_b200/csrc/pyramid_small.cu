// Whole-pyramid kernels for small images (the CIFAR-shaped configurations, and the coarse tail of any pyramid): every level
// from `first` on fits shared memory together (64x64x3 and everything below it is 65 KB), so ONE launch does the whole level
// chain of an image -- split (normalise + Gaussian + decimate + band), merge (bilinear x2 + add, coarse to fine) and the
// merge adjoint -- instead of one launch per pair of levels.  Measured with cold caches (scripts/pyr_probe.py): 32x32x3,
// five levels, batch 128: merge 19.4 -> 14.3 us, adjoint 13.3 -> 12.3 us, split 15.4 -> 18.4 us (so the split keeps the
// two-level launches unless MVAE_PYR_SMALL_SPLIT=1); the coarse tail of 512x512x3 with nine levels: merge 230 -> 208 us.
// Arithmetic follows the per-level kernels of pyramid.cu expression by expression (multiscale_vae.py:129-160, 204-224,
// 292-315).
#include "common.cuh"

namespace mvae {
namespace pys {

constexpr int kThreads = 256;
constexpr int kMaxLevels = 12;
// One CTA per image: beyond 32x32 pixels the serial level chain of a single CTA loses to the tiled two-level launches
// (measured, 64x64x3, 6 levels, batch 128: split 47 us here).  32x32x4 and everything below it is 5.5 K floats.
constexpr int kMaxSmemFloats = 6 * 1024;

struct Args {
    const float* src[kMaxLevels];      // split: [0] = image / x_first; merge: ys[first + i]; adjoint: [0] = d_first
    float* dst[kMaxLevels];            // split: bands[first + i]; merge: [0] = r_first; adjoint: dys[first + i], i >= 1
    int off[kMaxLevels + 1];           // shared-memory offset of level i (floats, multiples of 4)
    int n, B, h0, w0, C;
    float na, nb;                      // split: x * na + nb normalises level 0 (1, 0 when the input is x_first already)
    float taps[9];
};

__device__ __forceinline__ void up2_taps(int y, int Hc, int& i0, int& i1, float& w1) {
    const int k = y >> 1;
    if (y & 1) { i0 = k; i1 = min(k + 1, Hc - 1); w1 = 0.25f; }
    else       { i0 = max(k - 1, 0); i1 = k; w1 = 0.75f; }
}

// global <-> shared copies of one level of one image (16-byte vectors when the level is a multiple of 4 floats: every level
// of >= 2x2 pixels with C = 3, and the per-image base offsets then stay 16-byte aligned)
__device__ __forceinline__ void load_level(float* S, const float* __restrict__ g, int n, bool vec) {
    if (vec) {
        for (int i = threadIdx.x; i < (n >> 2); i += kThreads) reinterpret_cast<float4*>(S)[i] = __ldg(reinterpret_cast<const float4*>(g) + i);
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) S[i] = __ldg(g + i);
    }
}
__device__ __forceinline__ void store_level(float* __restrict__ g, const float* S, int n, bool vec) {
    if (vec) {
        for (int i = threadIdx.x; i < (n >> 2); i += kThreads) reinterpret_cast<float4*>(g)[i] = reinterpret_cast<const float4*>(S)[i];
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) g[i] = S[i];
    }
}
__device__ __forceinline__ bool vec_ok(const void* p, int n) { return (n & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// r_first = y_first + up2(y_{first+1} + up2(... y_last))
__global__ void __launch_bounds__(kThreads) merge_small_kernel(const Args a) {
    pdl_sync();
    extern __shared__ __align__(16) float S[];
    const int C = a.C;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        for (int i = 0; i < a.n; ++i) {
            const int n = (a.h0 >> i) * (a.w0 >> i) * C;
            load_level(S + a.off[i], a.src[i] + (long long)b * n, n, vec_ok(a.src[i], n));
        }
        __syncthreads();
        for (int i = a.n - 2; i >= 0; --i) {
            const int H = a.h0 >> i, W = a.w0 >> i, Hc = H >> 1, Wc = W >> 1;
            float* fine = S + a.off[i];
            const float* cb = S + a.off[i + 1];
            for (int idx = threadIdx.x; idx < H * W * C; idx += kThreads) {
                const int c = idx % C;
                const int p = idx / C;
                const int x = p % W, y = p / W;
                int y0, y1, x0, x1; float wy, wx;
                up2_taps(y, Hc, y0, y1, wy);
                up2_taps(x, Wc, x0, x1, wx);
                const float c00 = cb[(y0 * Wc + x0) * C + c], c01 = cb[(y0 * Wc + x1) * C + c];
                const float c10 = cb[(y1 * Wc + x0) * C + c], c11 = cb[(y1 * Wc + x1) * C + c];
                const float top = c00 + (c01 - c00) * wx, bot = c10 + (c11 - c10) * wx;
                fine[idx] = fine[idx] + (top + (bot - top) * wy);
            }
            __syncthreads();
        }
        const int n0 = a.h0 * a.w0 * C;
        store_level(a.dst[0] + (long long)b * n0, S, n0, vec_ok(a.dst[0], n0));
    }
}

// d_{i} = up2^T(d_{i-1}), i = first+1 .. last, from d_first
__global__ void __launch_bounds__(kThreads) adjoint_small_kernel(const Args a) {
    pdl_sync();
    extern __shared__ __align__(16) float S[];
    const int C = a.C;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        const int n0 = a.h0 * a.w0 * C;
        load_level(S, a.src[0] + (long long)b * n0, n0, vec_ok(a.src[0], n0));
        __syncthreads();
        for (int i = 1; i < a.n; ++i) {
            const int Hc = a.h0 >> i, Wc = a.w0 >> i, H = Hc * 2, W = Wc * 2;
            const float* fb = S + a.off[i - 1];
            float* dc = S + a.off[i];
            float* gout = a.dst[i] + (long long)b * Hc * Wc * C;
            for (int idx = threadIdx.x; idx < Hc * Wc * C; idx += kThreads) {
                const int c = idx % C;
                const int p = idx / C;
                const int X = p % Wc, Y = p / Wc;
                float wy[4] = {0.25f, 0.75f, 0.75f, 0.25f}, wx[4] = {0.25f, 0.75f, 0.75f, 0.25f};
                if (Y == 0) { wy[0] = 0.f; wy[1] = 1.f; }
                if (Y == Hc - 1) { wy[3] = 0.f; wy[2] = 1.f; }
                if (X == 0) { wx[0] = 0.f; wx[1] = 1.f; }
                if (X == Wc - 1) { wx[3] = 0.f; wx[2] = 1.f; }
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int y = 2 * Y - 1 + j;
                    if (y < 0 || y >= H) continue;
                    float r = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int x = 2 * X - 1 + k;
                        if (x < 0 || x >= W) continue;
                        r = fmaf(wx[k], fb[(y * W + x) * C + c], r);
                    }
                    acc = fmaf(wy[j], r, acc);
                }
                dc[idx] = acc;
                gout[idx] = acc;
            }
            __syncthreads();
        }
    }
}

// band_i = x_i - G(x_i), x_{i+1} = G(x_i) at even pixels (SAME zero padding of the normalised level), band_last = x_last
__global__ void __launch_bounds__(kThreads) split_small_kernel(const Args a) {
    pdl_sync();
    extern __shared__ __align__(16) float S[];
    const int C = a.C;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = a.taps[k];
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        const int n0 = a.h0 * a.w0 * C;
        {
            const float* g = a.src[0] + (long long)b * n0;
            if (vec_ok(a.src[0], n0)) {
                for (int i = threadIdx.x; i < (n0 >> 2); i += kThreads) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
                    reinterpret_cast<float4*>(S)[i] = make_float4(fmaf(v.x, a.na, a.nb), fmaf(v.y, a.na, a.nb), fmaf(v.z, a.na, a.nb),
                                                                  fmaf(v.w, a.na, a.nb));
                }
            } else {
                for (int i = threadIdx.x; i < n0; i += kThreads) S[i] = fmaf(__ldg(g + i), a.na, a.nb);
            }
        }
        __syncthreads();
        for (int i = 0; i + 1 < a.n; ++i) {
            const int H = a.h0 >> i, W = a.w0 >> i, WC = W * C;
            const float* img = S + a.off[i];
            float* down = S + a.off[i + 1];
            float* band = a.dst[i] + (long long)b * H * WC;
            for (int idx = threadIdx.x; idx < H * WC; idx += kThreads) {
                const int xc = idx % WC, y = idx / WC;
                const int x = xc / C;
                float f = 0.f, centre = 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = y + ky - 1;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int xx = x + kx - 1;
                        if (xx < 0 || xx >= W) continue;
                        const float v = img[yy * WC + xc + (kx - 1) * C];
                        f = fmaf(t[ky * 3 + kx], v, f);
                        if (ky == 1 && kx == 1) centre = v;
                    }
                }
                band[idx] = centre - f;
                if (!(y & 1) && !(x & 1)) down[((y >> 1) * (W >> 1) + (x >> 1)) * C + (xc - x * C)] = f;
            }
            __syncthreads();
        }
        const int i = a.n - 1;
        const int nl = (a.h0 >> i) * (a.w0 >> i) * C;
        store_level(a.dst[i] + (long long)b * nl, S + a.off[i], nl, vec_ok(a.dst[i], nl));
    }
}

// shared-memory layout of levels first .. first+n-1; false when they do not fit
static bool layout(Args& a, int B, int h0, int w0, int C, int n) {
    if (n < 1 || n > kMaxLevels || B < 1) return false;
    if ((h0 % (1 << (n - 1))) || (w0 % (1 << (n - 1)))) return false;
    int o = 0;
    for (int i = 0; i < n; ++i) {
        a.off[i] = o;
        o += (((h0 >> i) * (w0 >> i) * C + 3) / 4) * 4;
    }
    a.off[n] = o;
    if (o > kMaxSmemFloats) return false;
    a.n = n; a.B = B; a.h0 = h0; a.w0 = w0; a.C = C;
    return true;
}

template <int WHICH>      // one `configured` flag per kernel (the three kernels share a function type)
static int launch(void (*kernel)(const Args), const Args& a, cudaStream_t s) {
    const size_t smem = (size_t)a.off[a.n] * sizeof(float);
    static DeviceOnce configured;
    if (configured.first()) MVAE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemFloats * 4));
    const int per_sm = smem > 0 ? (int)((200 * 1024) / (smem + 1024)) : 8;
    int grid = kNumSMs * (per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm);
    if (grid > a.B) grid = a.B;
    MVAE_CUDA(launch_pdl(kernel, dim3(grid), dim3(kThreads), smem, s, a));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

}  // namespace pys

// First level (even, so that the two-level launches above it still pair up) from which the rest of the pyramid fits shared
// memory; -1 when that leaves fewer than two levels.  MVAE_PYR_SMALL=0 switches the path off.
int pyr_small_first(int B, int H, int W, int C, int levels, int is_split) {
    static int enabled = -1, split_enabled = -1;
    if (enabled < 0) { enabled = env_int("MVAE_PYR_SMALL", 1); split_enabled = env_int("MVAE_PYR_SMALL_SPLIT", 0); }
    if (!enabled || (is_split && !split_enabled)) return -1;
    for (int k = 0; k + 2 <= levels; k += 2) {
        pys::Args a = {};
        if (pys::layout(a, B, H >> k, W >> k, C, levels - k)) return k;
    }
    return -1;
}

// levels first .. first+n-1 of a pyramid whose level `first` is (B, h0, w0, C).  MVAE_ERR_UNSUPPORTED: they do not fit.
int pyr_merge_small(const float* const* ys, float* out, int B, int h0, int w0, int C, int n, cudaStream_t s) {
    pys::Args a = {};
    if (!pys::layout(a, B, h0, w0, C, n)) return MVAE_ERR_UNSUPPORTED;
    for (int i = 0; i < n; ++i) a.src[i] = ys[i];
    a.dst[0] = out;
    return pys::launch<0>(pys::merge_small_kernel, a, s);
}

int pyr_adjoint_small(const float* d0, float* const* ds, int B, int h0, int w0, int C, int n, cudaStream_t s) {
    pys::Args a = {};
    if (!pys::layout(a, B, h0, w0, C, n)) return MVAE_ERR_UNSUPPORTED;
    a.src[0] = d0;
    for (int i = 1; i < n; ++i) a.dst[i] = ds[i];
    return pys::launch<1>(pys::adjoint_small_kernel, a, s);
}

int pyr_split_small(const float* x, float* const* bands, int B, int h0, int w0, int C, int n, const float* taps9, float na,
                    float nb, cudaStream_t s) {
    pys::Args a = {};
    if (!pys::layout(a, B, h0, w0, C, n)) return MVAE_ERR_UNSUPPORTED;
    a.src[0] = x;
    for (int i = 0; i < n; ++i) a.dst[i] = bands[i];
    a.na = na; a.nb = nb;
    for (int k = 0; k < 9; ++k) a.taps[k] = taps9[k];
    return pys::launch<2>(pys::split_small_kernel, a, s);
}

}  // namespace mvae
