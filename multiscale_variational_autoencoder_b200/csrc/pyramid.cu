// Multiscale pyramid: split (normalize + Gaussian + decimate + band), merge (bilinear x2 + add) and its adjoint.
// Reference: multiscale_vae.py:79-94,129-160,204-224,292-315; layer_blocks.py:23-185,980-1050.
#include "common.cuh"

namespace mvae {

struct Taps {
    float t[49];
    int kh, kw;
};

// One pyramid level.  src is the raw image when `affine` (level 0: v*na + nb normalises on the fly) else x_i.
//   f      = SAME zero-padded Gaussian of the (normalised) level
//   band   = centre - f                      (DIFF_NO_UPSAMPLE; band == nullptr skips it)
//   down   = f at even rows/cols             (down == nullptr skips it)
//   filt   = f at full resolution            (gaussian_filter_block; nullptr skips it)
__global__ void __launch_bounds__(256) split_level_kernel(const float* __restrict__ src, float* __restrict__ band,
                                                          float* __restrict__ down, float* __restrict__ filt,
                                                          int B, int H, int W, int C, Taps taps, float na, float nb) {
    pdl_sync();
    const long long total = (long long)B * H * W * C;
    const int WC = W * C;
    const int ph = (taps.kh - 1) / 2, pw = (taps.kw - 1) / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xc = (int)(idx % WC);
        const long long row = idx / WC;
        const int y = (int)(row % H);
        const int x = xc / C;
        const float* img = src + (row - y) * WC;   // start of this image
        float f = 0.f, centre = 0.f;
        for (int ky = 0; ky < taps.kh; ++ky) {
            const int yy = y + ky - ph;
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < taps.kw; ++kx) {
                const int xx = x + kx - pw;
                if (xx < 0 || xx >= W) continue;
                const float v = fmaf(__ldg(img + (long long)yy * WC + xc + (kx - pw) * C), na, nb);
                f = fmaf(taps.t[ky * taps.kw + kx], v, f);
                if (ky == ph && kx == pw) centre = v;
            }
        }
        if (band) band[idx] = centre - f;
        if (filt) filt[idx] = f;
        if (down && !(y & 1) && !(x & 1)) {
            const long long b = row / H;
            down[((b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * C + (xc - x * C)] = f;
        }
    }
}

__device__ __forceinline__ void up2_taps(int y, int Hc, int& i0, int& i1, float& w1) {
    const int k = y >> 1;
    if (y & 1) { i0 = k; i1 = min(k + 1, Hc - 1); w1 = 0.25f; }
    else       { i0 = max(k - 1, 0); i1 = k; w1 = 0.75f; }
}

// out = a*fine(normalised or not) + sgn * bilinear_up2(coarse).   Used for
//   Laplacian band : band = norm(x_i) - up2(x_{i+1})   (sgn = -1, fine optionally normalised)
//   merge          : r_i  = y_i + up2(r_{i+1})         (sgn = +1)
__global__ void __launch_bounds__(256) up2_combine_kernel(const float* __restrict__ fine,
                                                          const float* __restrict__ coarse, float* __restrict__ out,
                                                          int B, int H, int W, int C, float na, float nb, float sgn) {
    pdl_sync();
    const long long total = (long long)B * H * W * C;
    const int Hc = H >> 1, Wc = W >> 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long long p = idx / C;
        const int x = (int)(p % W); p /= W;
        const int y = (int)(p % H);
        const long long b = p / H;
        int y0, y1, x0, x1; float wy, wx;
        up2_taps(y, Hc, y0, y1, wy);
        up2_taps(x, Wc, x0, x1, wx);
        const float* cb = coarse + b * (long long)Hc * Wc * C + c;
        const float c00 = __ldg(cb + ((long long)y0 * Wc + x0) * C), c01 = __ldg(cb + ((long long)y0 * Wc + x1) * C);
        const float c10 = __ldg(cb + ((long long)y1 * Wc + x0) * C), c11 = __ldg(cb + ((long long)y1 * Wc + x1) * C);
        const float top = c00 + (c01 - c00) * wx, bot = c10 + (c11 - c10) * wx;
        const float up = top + (bot - top) * wy;
        out[idx] = fmaf(__ldg(fine + idx), na, nb) + sgn * up;
    }
}

// adjoint of bilinear x2: dcoarse[Y,X] = sum over the <=4x4 fine window
__global__ void __launch_bounds__(256) up2_adjoint_kernel(const float* __restrict__ dfine, float* __restrict__ dcoarse,
                                                          int B, int Hc, int Wc, int C) {
    pdl_sync();
    const long long total = (long long)B * Hc * Wc * C;
    const int H = Hc * 2, W = Wc * 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long long p = idx / C;
        const int X = (int)(p % Wc); p /= Wc;
        const int Y = (int)(p % Hc);
        const long long b = p / Hc;
        float wy[4] = {0.25f, 0.75f, 0.75f, 0.25f}, wx[4] = {0.25f, 0.75f, 0.75f, 0.25f};
        if (Y == 0) { wy[0] = 0.f; wy[1] = 1.f; }
        if (Y == Hc - 1) { wy[3] = 0.f; wy[2] = 1.f; }
        if (X == 0) { wx[0] = 0.f; wx[1] = 1.f; }
        if (X == Wc - 1) { wx[3] = 0.f; wx[2] = 1.f; }
        const float* fb = dfine + b * (long long)H * W * C + c;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = 2 * Y - 1 + j;
            if (y < 0 || y >= H) continue;
            float r = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = 2 * X - 1 + i;
                if (x < 0 || x >= W) continue;
                r = fmaf(wx[i], __ldg(fb + ((long long)y * W + x) * C), r);
            }
            acc = fmaf(wy[j], r, acc);
        }
        dcoarse[idx] = acc;
    }
}

__global__ void __launch_bounds__(256) affine_clip_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          long long n, float a, float b, float lo, float hi) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(fmaf(__ldg(in + i), a, b), lo), hi);
}

__global__ void __launch_bounds__(256) coord_channels_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                             int H, int W, int C, int extra) {
    pdl_sync();
    const int Co = C + extra;
    const long long total = (long long)B * H * W * Co;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Co);
        const long long p = idx / Co;
        const int j = (int)(p % W);
        const int i = (int)((p / W) % H);
        float v;
        if (c < C) v = __ldg(x + p * C + c);
        else {
            const float xx = (float)i / (float)(H - 1) * 2.f - 1.f;   // coord.py:117-119
            const float yy = (float)j / (float)(W - 1) * 2.f - 1.f;   // coord.py:121-123
            v = (c == C) ? xx : (c == C + 1) ? yy : sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f));
        }
        y[idx] = v;
    }
}

// pyramid_tiled.cu: two levels per launch, shared-memory halo staging; MVAE_ERR_UNSUPPORTED when the shape is not covered
int pyr_split_pair(const float* src, float* band0, float* band1, float* down2, int B, int h, int w, int C,
                   const float* taps9, float na, float nb, int filter_second, cudaStream_t s);
int pyr_merge_pair(const float* y0, const float* y1, const float* r2, float* out, int B, int h, int w, int C,
                   cudaStream_t s);
int pyr_adjoint_pair(const float* d0, float* d1, float* d2, int B, int h, int w, int C, cudaStream_t s);
// whole-pyramid kernels for the levels that fit shared memory together (pyramid_small.cu)
int pyr_small_first(int B, int H, int W, int C, int levels, int is_split);
int pyr_merge_small(const float* const* ys, float* out, int B, int h0, int w0, int C, int n, cudaStream_t s);
int pyr_adjoint_small(const float* d0, float* const* ds, int B, int h0, int w0, int C, int n, cudaStream_t s);
int pyr_split_small(const float* x, float* const* bands, int B, int h0, int w0, int C, int n, const float* taps9, float na,
                    float nb, cudaStream_t s);

// training-time corruption of the input transform (multiscale_vae.py:139-147): GaussianNoise in normalised space, then
// SpatialDropout2D (whole channels of a sample dropped, survivors scaled).  Written back in RAW units so that the pyramid
// split (which normalises) sees exactly keep*scale*(norm(x) + std*noise); the loss keeps comparing against the clean x.
__global__ void __launch_bounds__(256) input_corrupt_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                            const float* __restrict__ keep, float* __restrict__ out,
                                                            long long total, int HWC, int C, float na, float nb, float da,
                                                            float db, float noise_std, float keep_scale) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float t = fmaf(__ldg(x + i), na, nb);
        if (noise) t = fmaf(noise_std, __ldg(noise + i), t);
        if (keep) t *= __ldg(keep + (i / HWC) * C + (i % C)) * keep_scale;
        out[i] = fmaf(t, da, db);
    }
}

static inline int grid_for(long long total, int threads = 256) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int fill_taps(Taps& t, const float* taps, int kh, int kw) {
    MVAE_REQUIRE(taps && kh > 0 && kw > 0 && kh <= 7 && kw <= 7 && (kh & 1) && (kw & 1), "taps: kh,kw must be odd <= 7");
    t.kh = kh; t.kw = kw;
    for (int i = 0; i < kh * kw; ++i) t.t[i] = taps[i];
    return MVAE_OK;
}

}  // namespace mvae

using namespace mvae;

extern "C" size_t mvae_pyramid_split_workspace_bytes(int B, int H, int W, int C, int levels) {
    // x_1 .. x_{levels-2}  (x_{levels-1} is written straight into bands[levels-1])
    size_t n = 0;
    for (int i = 1; i <= levels - 2; ++i) n += (size_t)B * (H >> i) * (W >> i) * C;
    return n * sizeof(float) + 256;
}

extern "C" int mvae_pyramid_split(const float* x, float* const* bands, void* workspace, int B, int H, int W, int C,
                                  int levels, float v0, float v1, const float* taps, int kh, int kw, int diff_mode,
                                  mvae_stream_t stream) {
    MVAE_REQUIRE(x && bands && levels >= 1 && B > 0 && C > 0, "pyramid_split: bad arguments");
    MVAE_REQUIRE((H % (1 << (levels - 1))) == 0 && (W % (1 << (levels - 1))) == 0,
                 "pyramid_split: H=%d W=%d must be divisible by 2^(levels-1)=%d", H, W, 1 << (levels - 1));
    MVAE_REQUIRE(diff_mode == MVAE_DIFF_NO_UPSAMPLE || diff_mode == MVAE_DIFF_LAPLACIAN, "pyramid_split: diff_mode");
    MVAE_REQUIRE(levels <= 2 || workspace, "pyramid_split: workspace required");
    Taps t;
    if (int e = fill_taps(t, taps, kh, kw)) return e;
    cudaStream_t s = as_stream(stream);
    const float na = 2.f / (v1 - v0), nb = -2.f * v0 / (v1 - v0) - 1.f;   // 2(y-v0)/(v1-v0)-1
    if (levels == 1) {
        const long long n = (long long)B * H * W * C;
        MVAE_CUDA(launch_pdl(affine_clip_kernel, dim3(grid_for(n)), dim3(256), 0, s, x, bands[0], n, na, nb, -INFINITY, INFINITY));
        MVAE_LAUNCH_CHECK();
        return MVAE_OK;
    }
    float* ws = reinterpret_cast<float*>(workspace);
    // x_i (1 <= i <= levels-2) lives in the workspace; x_{levels-1} is bands[levels-1]
    float* xs[32];
    {
        size_t o = 0;
        for (int i = 1; i <= levels - 2; ++i) { xs[i] = ws + o; o += (size_t)B * (H >> i) * (W >> i) * C; }
        xs[levels - 1] = bands[levels - 1];
    }
    const float* src = x;
    float a = na, b = nb;
    int i = 0;
    const int small = (diff_mode == MVAE_DIFF_NO_UPSAMPLE && kh == 3 && kw == 3) ? pyr_small_first(B, H, W, C, levels, 1) : -1;
    while (i < levels - 1) {
        const int h = H >> i, w = W >> i;
        const long long n = (long long)B * h * w * C;
        if (i == small) {
            // every remaining level fits shared memory: one launch finishes the pyramid
            const int rc = pyr_split_small(src, bands + i, B, h, w, C, levels - i, taps, a, b, s);
            if (rc == MVAE_OK) break;
            if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        }
        if (diff_mode == MVAE_DIFF_NO_UPSAMPLE && kh == 3 && kw == 3) {
            // tiled fast path: levels i and i+1 in one launch
            const int second = (i + 1 <= levels - 2) ? 1 : 0;
            float* down2 = second ? xs[i + 2] : nullptr;
            const int rc = pyr_split_pair(src, bands[i], bands[i + 1], down2, B, h, w, C, taps, a, b, second, s);
            if (rc == MVAE_OK) {
                if (!second) break;
                src = xs[i + 2]; a = 1.f; b = 0.f; i += 2;
                continue;
            }
            if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        }
        float* down = xs[i + 1];
        float* band = (diff_mode == MVAE_DIFF_NO_UPSAMPLE) ? bands[i] : nullptr;
        MVAE_CUDA(launch_pdl(split_level_kernel, dim3(grid_for(n)), dim3(256), 0, s, src, band, down, nullptr, B, h, w, C, t, a, b));
        MVAE_LAUNCH_CHECK();
        if (diff_mode == MVAE_DIFF_LAPLACIAN) {
            MVAE_CUDA(launch_pdl(up2_combine_kernel, dim3(grid_for(n)), dim3(256), 0, s, src, down, bands[i], B, h, w, C, a, b, -1.f));
            MVAE_LAUNCH_CHECK();
        }
        src = down;
        a = 1.f; b = 0.f;
        ++i;
    }
    return MVAE_OK;
}

extern "C" int mvae_gaussian_filter(const float* x, float* y, int B, int H, int W, int C, const float* taps, int kh,
                                    int kw, mvae_stream_t stream) {
    MVAE_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "gaussian_filter: bad arguments");
    Taps t;
    if (int e = fill_taps(t, taps, kh, kw)) return e;
    const long long n = (long long)B * H * W * C;
    MVAE_CUDA(launch_pdl(split_level_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), x, nullptr, nullptr, y, B, H, W, C, t, 1.f, 0.f));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" size_t mvae_pyramid_merge_workspace_bytes(int B, int H, int W, int C, int levels) {
    size_t n = 0;   // r_1 .. r_{levels-2}
    for (int i = 1; i <= levels - 2; ++i) n += (size_t)B * (H >> i) * (W >> i) * C;
    return n * sizeof(float) + 256;
}

extern "C" int mvae_pyramid_merge_fwd(const float* const* ys, float* r0, void* workspace, int B, int H, int W, int C,
                                      int levels, mvae_stream_t stream) {
    MVAE_REQUIRE(ys && r0 && levels >= 2, "pyramid_merge_fwd: needs levels >= 2 (multiscale_vae.py:210-222)");
    MVAE_REQUIRE((H % (1 << (levels - 1))) == 0 && (W % (1 << (levels - 1))) == 0, "pyramid_merge_fwd: H,W");
    MVAE_REQUIRE(levels <= 2 || workspace, "pyramid_merge_fwd: workspace required");
    cudaStream_t s = as_stream(stream);
    // r_i buffers: r_0 = r0, r_i (1 <= i <= L-2) in the workspace, r_{L-1} = ys[L-1]
    float* ws = reinterpret_cast<float*>(workspace);
    size_t off[32];
    size_t o = 0;
    for (int i = 1; i <= levels - 2; ++i) { off[i] = o; o += (size_t)B * (H >> i) * (W >> i) * C; }
    const float* coarse = ys[levels - 1];
    int cur = levels - 1;                      // r_cur is available at `coarse`
    {
        // the coarse levels that fit shared memory together: one launch up to r_small (the whole merge when small == 0)
        const int small = pyr_small_first(B, H, W, C, levels, 0);
        if (small >= 0) {
            float* out = (small == 0) ? r0 : ws + off[small];
            const int rc = pyr_merge_small(ys + small, out, B, H >> small, W >> small, C, levels - small, s);
            if (rc == MVAE_OK) { coarse = out; cur = small; }
            else if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        }
    }
    while (cur > 0) {
        if (cur >= 2) {
            const int k = cur - 2;
            float* out = (k == 0) ? r0 : ws + off[k];
            const int rc = pyr_merge_pair(ys[k], ys[k + 1], coarse, out, B, H >> k, W >> k, C, s);
            if (rc == MVAE_OK) { coarse = out; cur = k; continue; }
            if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        } else {
            const int rc = pyr_merge_pair(ys[0], coarse, nullptr, r0, B, H, W, C, s);
            if (rc == MVAE_OK) break;
            if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        }
        const int i = cur - 1;
        const int h = H >> i, w = W >> i;
        const long long n = (long long)B * h * w * C;
        float* out = (i == 0) ? r0 : ws + off[i];
        MVAE_CUDA(launch_pdl(up2_combine_kernel, dim3(grid_for(n)), dim3(256), 0, s, ys[i], coarse, out, B, h, w, C, 1.f, 0.f, 1.f));
        MVAE_LAUNCH_CHECK();
        coarse = out;
        cur = i;
    }
    return MVAE_OK;
}

extern "C" int mvae_pyramid_merge_bwd(const float* dr0, float* const* dys, int B, int H, int W, int C, int levels,
                                      mvae_stream_t stream) {
    MVAE_REQUIRE(dr0 && dys && levels >= 2, "pyramid_merge_bwd: bad arguments");
    cudaStream_t s = as_stream(stream);
    if (dys[0] != dr0)
        MVAE_CUDA(cudaMemcpyAsync(dys[0], dr0, (size_t)B * H * W * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int k = 0;                                 // d_k is available in dys[k]
    const int small = pyr_small_first(B, H, W, C, levels, 0);
    while (k < levels - 1) {
        if (k == small) {
            const int rc = pyr_adjoint_small(dys[k], dys + k, B, H >> k, W >> k, C, levels - k, s);
            if (rc == MVAE_OK) break;
            if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        }
        float* d2 = (k + 2 <= levels - 1) ? dys[k + 2] : nullptr;
        const int rc = pyr_adjoint_pair(dys[k], dys[k + 1], d2, B, H >> k, W >> k, C, s);
        if (rc == MVAE_OK) { k += d2 ? 2 : 1; continue; }
        if (rc != MVAE_ERR_UNSUPPORTED) return rc;
        const int hc = H >> (k + 1), wc = W >> (k + 1);
        const long long n = (long long)B * hc * wc * C;
        MVAE_CUDA(launch_pdl(up2_adjoint_kernel, dim3(grid_for(n)), dim3(256), 0, s, dys[k], dys[k + 1], B, hc, wc, C));
        MVAE_LAUNCH_CHECK();
        ++k;
    }
    return MVAE_OK;
}

extern "C" int mvae_denormalize_clip(const float* r0, float* out, long long n, float v0, float v1,
                                     mvae_stream_t stream) {
    MVAE_REQUIRE(r0 && out && n > 0, "denormalize_clip: bad arguments");
    const float a = (v1 - v0) * 0.5f, b = (v1 - v0) * 0.5f + v0;   // (y+1)(v1-v0)/2 + v0
    MVAE_CUDA(launch_pdl(affine_clip_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), r0, out, n, a, b, v0, v1));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_coord_channels(const float* x, float* y, int B, int H, int W, int C, int use_radius,
                                   mvae_stream_t stream) {
    MVAE_REQUIRE(x && y && B > 0 && H > 1 && W > 1 && C > 0, "coord_channels: needs H,W > 1 (coord.py:118,122)");
    const int extra = use_radius ? 3 : 2;
    const long long n = (long long)B * H * W * (C + extra);
    MVAE_CUDA(launch_pdl(coord_channels_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), x, y, B, H, W, C, extra));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_input_corrupt(const float* x, const float* noise, const float* keep, float* out, int B, int HW, int C,
                                  float v0, float v1, float noise_std, float keep_scale, mvae_stream_t stream) {
    MVAE_REQUIRE(x && out && B > 0 && HW > 0 && C > 0 && v1 > v0, "input_corrupt: bad arguments");
    const long long n = (long long)B * HW * C;
    const float na = 2.f / (v1 - v0), nb = -2.f * v0 / (v1 - v0) - 1.f;
    const float da = (v1 - v0) * 0.5f, db = (v1 - v0) * 0.5f + v0;
    MVAE_CUDA(launch_pdl(input_corrupt_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), x, noise, keep, out, n, HW * C, C,
                         na, nb, da, db, noise_std, keep_scale));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}
