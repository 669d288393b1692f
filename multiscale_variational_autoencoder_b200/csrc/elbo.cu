// ELBO kernels: reparameterisation + per-scale analytic KL, reconstruction loss (forward sums / backward), finalize.
// Reference: multiscale_vae.py:372-383 (sample), 453-495 (losses); multiscale_vae_.py:29-34 (0.5*logvar form).
#include "common.cuh"

namespace mvae {

constexpr int kMaxC = 4;   // image channels supported by the loss kernels

// ---------------------------------------------------------------------------------------------------------
// z = mu + exp(s*lv) * std * eps ;  kl[b] = -0.5 * sum_k (1 + lv - mu^2 - exp(lv)).  One warp per sample.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reparam_kl_fwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                                             float* __restrict__ z, float* __restrict__ kl, int B, int zd,
                                                             float s, float sd) {
    pdl_sync();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* row = mulv + (long long)warp * 2 * zd;
    float acc = 0.f;
    for (int k = lane; k < zd; k += 32) {
        const float mu = row[k], lv = row[zd + k];
        z[(long long)warp * zd + k] = fmaf(expf(s * lv) * sd, eps[(long long)warp * zd + k], mu);
        acc += (lv - expm1f(lv)) - mu * mu;      // 1 + lv - exp(lv) without the fp32 cancellation near lv = 0
    }
    acc = warp_sum(acc);
    if (lane == 0) kl[warp] = -0.5f * acc;
}

__global__ void __launch_bounds__(256) reparam_kl_bwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                                             const float* __restrict__ dz, float* __restrict__ dmulv,
                                                             int B, int zd, float s, float sd, float kls) {
    pdl_sync();
    const long long total = (long long)B * zd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / zd;
        const int k = (int)(i - b * zd);
        const float mu = mulv[b * 2 * zd + k], lv = mulv[b * 2 * zd + zd + k];
        const float g = dz[i];
        dmulv[b * 2 * zd + k] = fmaf(kls, mu, g);
        dmulv[b * 2 * zd + zd + k] = g * s * expf(s * lv) * sd * eps[i] + kls * 0.5f * expm1f(lv);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Reconstruction loss.  yh = clip(r0*a + bb, v0, v1).  grid = (chunks, B); a thread walks whole pixels.
// sums[b] = { sum|y-yh|, sum_hw (y-yh)[c] for c<C, same over the centre crop }.
// ---------------------------------------------------------------------------------------------------------
struct Crop { int r0, r1, c0, c1; };
__device__ __forceinline__ float sgnf(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

__global__ void __launch_bounds__(256) recon_loss_fwd_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                             float* __restrict__ out, float* __restrict__ sums, int H,
                                                             int W, int C, float a, float bb, float v0, float v1,
                                                             Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int npix = H * W;
    const long long base = (long long)b * npix * C;
    float s1 = 0.f, d[kMaxC], dc[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) d[c] = dc[c] = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int h = p / W, w = p - h * W;
        const bool in = h >= crop.r0 && h < crop.r1 && w >= crop.c0 && w < crop.c1;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) {
            if (c < C) {
                const long long i = base + (long long)p * C + c;
                const float yh = fminf(fmaxf(fmaf(__ldg(r0 + i), a, bb), v0), v1);
                if (out) out[i] = yh;
                const float e = __ldg(y + i) - yh;
                s1 += fabsf(e);
                d[c] += e;
                if (in) dc[c] += e;
            }
        }
    }
    __shared__ float red[8][1 + 2 * kMaxC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s1 = warp_sum(s1);
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) { d[c] = warp_sum(d[c]); dc[c] = warp_sum(dc[c]); }
    if (lane == 0) {
        red[warp][0] = s1;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) { red[warp][1 + c] = d[c]; red[warp][1 + kMaxC + c] = dc[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 1 + 2 * kMaxC) {
        float v = 0.f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += red[wv][threadIdx.x];
        const int j = threadIdx.x;
        float* sb = sums + (long long)b * (1 + 2 * C);
        if (j == 0) atomicAdd(sb, v);
        else if (j <= kMaxC) { if (j - 1 < C) atomicAdd(sb + 1 + (j - 1), v); }
        else { if (j - 1 - kMaxC < C) atomicAdd(sb + 1 + C + (j - 1 - kMaxC), v); }
    }
}

__global__ void __launch_bounds__(256) recon_loss_bwd_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                             const float* __restrict__ sums, float* __restrict__ dr0,
                                                             int H, int W, int C, float a, float bb, float v0, float v1,
                                                             float r_scale, Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int npix = H * W;
    const long long base = (long long)b * npix * C;
    const float* sb = sums + (long long)b * (1 + 2 * C);
    const float k_px = 1.f / ((float)npix * (float)C);
    const float k_ch = 0.5f / ((float)C * (float)npix);
    const float k_cc = 0.5f / ((float)C * (float)((crop.r1 - crop.r0) * (crop.c1 - crop.c0)));
    float gch[kMaxC], gcc[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
        gch[c] = c < C ? sgnf(sb[1 + c]) * k_ch : 0.f;
        gcc[c] = c < C ? sgnf(sb[1 + C + c]) * k_cc : 0.f;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int h = p / W, w = p - h * W;
        const bool in = h >= crop.r0 && h < crop.r1 && w >= crop.c0 && w < crop.c1;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) {
            if (c < C) {
                const long long i = base + (long long)p * C + c;
                const float pre = fmaf(__ldg(r0 + i), a, bb);
                const float yh = fminf(fmaxf(pre, v0), v1);
                const float e = __ldg(y + i) - yh;
                // dL/dyh = -( sign(e)/(HWC) + 0.5 sign(D_c)/(C HW) + [crop] 0.5 sign(D'_c)/(C H'W') )
                float g = -(sgnf(e) * k_px + gch[c] + (in ? gcc[c] : 0.f));
                g = (pre >= v0 && pre <= v1) ? g * a * r_scale : 0.f;
                dr0[i] = g;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Vectorised variants (W % 4 == 0, 16-byte aligned): a thread handles groups of 4 pixels = C float4s, so the channel of
// every register is a compile-time constant and all loads / stores are 16 bytes wide.
// ---------------------------------------------------------------------------------------------------------
template <int C, bool WRITE_OUT>
__global__ void __launch_bounds__(256) recon_loss_fwd_vec_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                                 float* __restrict__ out, float* __restrict__ sums, int H,
                                                                 int W, float a, float bb, float v0, float v1, Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int ngroups = H * W / 4, gw = W / 4;
    const long long base = (long long)b * H * W * C;
    const float4* r4 = reinterpret_cast<const float4*>(r0 + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    float4* o4 = reinterpret_cast<float4*>(out + base);
    float s1 = 0.f, d[C], dc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = dc[c] = 0.f;
#pragma unroll 2
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        float rv[4 * C], yv[4 * C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float4 t = __ldg(r4 + (long long)g * C + k);
            rv[4 * k] = t.x; rv[4 * k + 1] = t.y; rv[4 * k + 2] = t.z; rv[4 * k + 3] = t.w;
            const float4 u = __ldg(y4 + (long long)g * C + k);
            yv[4 * k] = u.x; yv[4 * k + 1] = u.y; yv[4 * k + 2] = u.z; yv[4 * k + 3] = u.w;
        }
        const int h = g / gw, x0 = (g - h * gw) * 4;
        const bool rin = h >= crop.r0 && h < crop.r1;
        float yh[4 * C];
#pragma unroll
        for (int e = 0; e < 4 * C; ++e) {
            yh[e] = fminf(fmaxf(fmaf(rv[e], a, bb), v0), v1);
            const float err = yv[e] - yh[e];
            s1 += fabsf(err);
            d[e % C] += err;
            const int x = x0 + e / C;
            if (rin && x >= crop.c0 && x < crop.c1) dc[e % C] += err;
        }
        if (WRITE_OUT) {
#pragma unroll
            for (int k = 0; k < C; ++k) o4[(long long)g * C + k] = make_float4(yh[4 * k], yh[4 * k + 1], yh[4 * k + 2], yh[4 * k + 3]);
        }
    }
    __shared__ float red[8][1 + 2 * C];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s1 = warp_sum(s1);
#pragma unroll
    for (int c = 0; c < C; ++c) { d[c] = warp_sum(d[c]); dc[c] = warp_sum(dc[c]); }
    if (lane == 0) {
        red[warp][0] = s1;
#pragma unroll
        for (int c = 0; c < C; ++c) { red[warp][1 + c] = d[c]; red[warp][1 + C + c] = dc[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 1 + 2 * C) {
        float v = 0.f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += red[wv][threadIdx.x];
        atomicAdd(sums + (long long)b * (1 + 2 * C) + threadIdx.x, v);
    }
}

template <int C>
__global__ void __launch_bounds__(256) recon_loss_bwd_vec_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                                 const float* __restrict__ sums, float* __restrict__ dr0,
                                                                 int H, int W, float a, float bb, float v0, float v1,
                                                                 float r_scale, Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int npix = H * W, ngroups = npix / 4, gw = W / 4;
    const long long base = (long long)b * npix * C;
    const float* sb = sums + (long long)b * (1 + 2 * C);
    const float k_px = 1.f / ((float)npix * (float)C);
    const float k_ch = 0.5f / ((float)C * (float)npix);
    const float k_cc = 0.5f / ((float)C * (float)((crop.r1 - crop.r0) * (crop.c1 - crop.c0)));
    float gch[C], gcc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { gch[c] = sgnf(sb[1 + c]) * k_ch; gcc[c] = sgnf(sb[1 + C + c]) * k_cc; }
    const float4* r4 = reinterpret_cast<const float4*>(r0 + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    float4* o4 = reinterpret_cast<float4*>(dr0 + base);
#pragma unroll 2
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        float rv[4 * C], yv[4 * C], gv[4 * C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float4 t = __ldg(r4 + (long long)g * C + k);
            rv[4 * k] = t.x; rv[4 * k + 1] = t.y; rv[4 * k + 2] = t.z; rv[4 * k + 3] = t.w;
            const float4 u = __ldg(y4 + (long long)g * C + k);
            yv[4 * k] = u.x; yv[4 * k + 1] = u.y; yv[4 * k + 2] = u.z; yv[4 * k + 3] = u.w;
        }
        const int h = g / gw, x0 = (g - h * gw) * 4;
        const bool rin = h >= crop.r0 && h < crop.r1;
#pragma unroll
        for (int e = 0; e < 4 * C; ++e) {
            const float pre = fmaf(rv[e], a, bb);
            const float yh = fminf(fmaxf(pre, v0), v1);
            const float err = yv[e] - yh;
            const int x = x0 + e / C;
            const bool in = rin && x >= crop.c0 && x < crop.c1;
            float gg = -(sgnf(err) * k_px + gch[e % C] + (in ? gcc[e % C] : 0.f));
            gv[e] = (pre >= v0 && pre <= v1) ? gg * a * r_scale : 0.f;
        }
#pragma unroll
        for (int k = 0; k < C; ++k) o4[(long long)g * C + k] = make_float4(gv[4 * k], gv[4 * k + 1], gv[4 * k + 2], gv[4 * k + 3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Flat variants (W % 4 == 0, 16-byte aligned): consecutive threads read consecutive float4s -- every warp access is one
// contiguous 512-byte run (the per-pixel-group kernels above read 48-byte pieces per thread: three partial passes over every
// 128-byte line, 3.8 TB/s at 512x512x3 against 6.3 TB/s for a plain copy).  The channel of element j of float4 i is
// (4 i + j) mod C; blocks of 192 threads and a grid stride that is a multiple of 192 keep 4 i mod C fixed per thread, so a
// thread accumulates into slots by j mod C and maps slots to channels once, at the end.  A row is 3W/4.. float4s (W % 4 == 0):
// the four elements of a float4 share their image row, and the crop test is a range test on the float offset within the row.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFlatThreads = 192;

template <int C, bool WRITE_OUT>
__global__ void __launch_bounds__(kFlatThreads) recon_loss_fwd_flat_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                                           float* __restrict__ out, float* __restrict__ sums,
                                                                           int H, int W, float a, float bb, float v0, float v1,
                                                                           Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int RL = W * C;                                  // floats per image row
    const int n4 = H * RL / 4;
    const long long base = (long long)b * H * RL;
    const float4* r4 = reinterpret_cast<const float4*>(r0 + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    float4* o4 = reinterpret_cast<float4*>(out + base);
    const int i0 = blockIdx.x * kFlatThreads + threadIdx.x, T = gridDim.x * kFlatThreads;
    const int ph = (int)((4LL * i0) % C);                 // channel of element 0 of every float4 of this thread
    const int f0 = crop.c0 * C, f1 = crop.c1 * C;          // crop columns as a float-offset range within a row
    float s1 = 0.f, d[4], dc[4];                           // slots: element j of a float4
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = dc[j] = 0.f;
#pragma unroll 4
    for (int i = i0; i < n4; i += T) {
        const float4 rv = __ldg(r4 + i), yv = __ldg(y4 + i);
        const int e0 = 4 * i, h = e0 / RL, off = e0 - h * RL;
        const bool rin = h >= crop.r0 && h < crop.r1;
        const float r[4] = {rv.x, rv.y, rv.z, rv.w}, t[4] = {yv.x, yv.y, yv.z, yv.w};
        float yh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            yh[j] = fminf(fmaxf(fmaf(r[j], a, bb), v0), v1);
            const float err = t[j] - yh[j];
            s1 += fabsf(err);
            d[j] += err;
            if (rin && off + j >= f0 && off + j < f1) dc[j] += err;
        }
        if (WRITE_OUT) o4[i] = make_float4(yh[0], yh[1], yh[2], yh[3]);
    }
    // slots -> channels: element j belongs to channel (ph + j) mod C
    float dch[4] = {0.f, 0.f, 0.f, 0.f}, dcc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (ph + j) % C;
#pragma unroll
        for (int k = 0; k < C; ++k) { dch[k] += c == k ? d[j] : 0.f; dcc[k] += c == k ? dc[j] : 0.f; }
    }
    __shared__ float red[kFlatThreads / 32][1 + 2 * C];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s1 = warp_sum(s1);
#pragma unroll
    for (int c = 0; c < C; ++c) { dch[c] = warp_sum(dch[c]); dcc[c] = warp_sum(dcc[c]); }
    if (lane == 0) {
        red[warp][0] = s1;
#pragma unroll
        for (int c = 0; c < C; ++c) { red[warp][1 + c] = dch[c]; red[warp][1 + C + c] = dcc[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 1 + 2 * C) {
        float v = 0.f;
        for (int wv = 0; wv < kFlatThreads / 32; ++wv) v += red[wv][threadIdx.x];
        atomicAdd(sums + (long long)b * (1 + 2 * C) + threadIdx.x, v);
    }
}

template <int C>
__global__ void __launch_bounds__(kFlatThreads) recon_loss_bwd_flat_kernel(const float* __restrict__ r0, const float* __restrict__ y,
                                                                           const float* __restrict__ sums, float* __restrict__ dr0,
                                                                           int H, int W, float a, float bb, float v0, float v1,
                                                                           float r_scale, Crop crop) {
    pdl_sync();
    const int b = blockIdx.y;
    const int RL = W * C, npix = H * W;
    const int n4 = H * RL / 4;
    const long long base = (long long)b * H * RL;
    const float* sb = sums + (long long)b * (1 + 2 * C);
    const float k_px = 1.f / ((float)npix * (float)C);
    const float k_ch = 0.5f / ((float)C * (float)npix);
    const float k_cc = 0.5f / ((float)C * (float)((crop.r1 - crop.r0) * (crop.c1 - crop.c0)));
    const int i0 = blockIdx.x * kFlatThreads + threadIdx.x, T = gridDim.x * kFlatThreads;
    const int ph = (int)((4LL * i0) % C);
    float gch[4], gcc[4];                                  // per element slot j of this thread's float4s
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (ph + j) % C;
        gch[j] = sgnf(sb[1 + c]) * k_ch;
        gcc[j] = sgnf(sb[1 + C + c]) * k_cc;
    }
    const int f0 = crop.c0 * C, f1 = crop.c1 * C;
    const float4* r4 = reinterpret_cast<const float4*>(r0 + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    float4* o4 = reinterpret_cast<float4*>(dr0 + base);
#pragma unroll 4
    for (int i = i0; i < n4; i += T) {
        const float4 rv = __ldg(r4 + i), yv = __ldg(y4 + i);
        const int e0 = 4 * i, h = e0 / RL, off = e0 - h * RL;
        const bool rin = h >= crop.r0 && h < crop.r1;
        const float r[4] = {rv.x, rv.y, rv.z, rv.w}, t[4] = {yv.x, yv.y, yv.z, yv.w};
        float gv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float pre = fmaf(r[j], a, bb);
            const float yh = fminf(fmaxf(pre, v0), v1);
            const float err = t[j] - yh;
            const bool in = rin && off + j >= f0 && off + j < f1;
            const float gg = -(sgnf(err) * k_px + gch[j] + (in ? gcc[j] : 0.f));
            gv[j] = (pre >= v0 && pre <= v1) ? gg * a * r_scale : 0.f;
        }
        o4[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
}

// single block
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ kl,
                                                            int levels, float* __restrict__ per_sample,
                                                            float* __restrict__ scalars, int B, int npix, int C,
                                                            int ncrop, float rf, float kf) {
    pdl_sync();
    float t_loss = 0.f, t_r = 0.f, t_m = 0.f, t_kl = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* sb = sums + (long long)b * (1 + 2 * C);
        const float metric = sb[0] / ((float)npix * (float)C);
        float ch = 0.f, cc = 0.f;
        for (int c = 0; c < C; ++c) { ch += fabsf(sb[1 + c]) / (float)npix; cc += fabsf(sb[1 + C + c]) / (float)ncrop; }
        const float r = metric + 0.5f * (ch / (float)C + cc / (float)C);
        float k = 0.f;
        for (int l = 0; l < levels; ++l) k += kl[(long long)l * B + b];
        per_sample[b] = r; per_sample[B + b] = metric; per_sample[2 * B + b] = k;
        t_loss += r * rf + k * kf; t_r += r; t_m += metric; t_kl += k;
    }
    __shared__ float red[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    t_loss = warp_sum(t_loss); t_r = warp_sum(t_r); t_m = warp_sum(t_m); t_kl = warp_sum(t_kl);
    if (lane == 0) { red[warp][0] = t_loss; red[warp][1] = t_r; red[warp][2] = t_m; red[warp][3] = t_kl; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += red[wv][threadIdx.x];
        scalars[threadIdx.x] = v / (float)B;
    }
}

static Crop make_crop(int H, int W) {
    // multiscale_vae.py:459-460,472,475: d0=int(H/2); rows int(d0/2):int(d0*3/2)
    const int d0 = H / 2, d1 = W / 2;
    Crop c{d0 / 2, (d0 * 3) / 2, d1 / 2, (d1 * 3) / 2};
    return c;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_reparam_kl_fwd(const float* mulv, const float* eps, float* z, float* kl, int B, int zdim,
                                   float logvar_scale, float sample_std, mvae_stream_t stream) {
    MVAE_REQUIRE(mulv && eps && z && kl && B > 0 && zdim > 0, "reparam_kl_fwd: bad arguments");
    const int warps_per_block = 8;
    MVAE_CUDA(launch_pdl(reparam_kl_fwd_kernel, dim3(ceil_div(B, warps_per_block)), dim3(256), 0, as_stream(stream), mulv, eps, z, kl, B, zdim,
                                                                                      logvar_scale, sample_std));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_reparam_kl_bwd(const float* mulv, const float* eps, const float* dz, float* dmulv, int B, int zdim,
                                   float logvar_scale, float sample_std, float kl_scale, mvae_stream_t stream) {
    MVAE_REQUIRE(mulv && eps && dz && dmulv && B > 0 && zdim > 0, "reparam_kl_bwd: bad arguments");
    const long long n = (long long)B * zdim;
    int grid = ceil_div(n, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    MVAE_CUDA(launch_pdl(reparam_kl_bwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), mulv, eps, dz, dmulv, B, zdim, logvar_scale, sample_std,
                                                             kl_scale));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int loss_grid_x(int B, int npix) {
    // aim for ~4 waves of 256-thread blocks over the chip, at least one block per sample
    int gx = ceil_div((long long)kNumSMs * 8, B);
    const int maxx = ceil_div(npix, 256);
    if (gx > maxx) gx = maxx;
    return gx < 1 ? 1 : gx;
}

// blocks of 192 threads per image for the flat kernels: ~8 resident blocks per SM over the chip, never more than the image has
// float4s; any count works for the channel phase (the stride is a multiple of 192, and 4 * 192 is a multiple of every C <= 4)
static int flat_grid_x(int B, long long n4) {
    int gx = ceil_div((long long)kNumSMs * 10, B);
    const int maxx = ceil_div(n4, kFlatThreads * 4);
    if (gx > maxx) gx = maxx;
    return gx < 1 ? 1 : gx;
}

extern "C" int mvae_recon_loss_fwd(const float* r0, const float* y, float* out, float* sums, int B, int H, int W, int C,
                                   float v0, float v1, mvae_stream_t stream) {
    MVAE_REQUIRE(r0 && y && sums && B > 0 && H > 1 && W > 1, "recon_loss_fwd: bad arguments");
    MVAE_REQUIRE(C >= 1 && C <= kMaxC, "recon_loss_fwd: C=%d unsupported (max %d)", C, kMaxC);
    MVAE_REQUIRE(B <= 65535, "recon_loss_fwd: B too large");
    const float a = (v1 - v0) * 0.5f, bb = (v1 - v0) * 0.5f + v0;
    const bool vec = (W % 4) == 0 && al16(r0) && al16(y) && (out == nullptr || al16(out)) && ((long long)H * W * C) % 4 == 0;
    if (vec) {
        dim3 grid(flat_grid_x(B, (long long)H * W * C / 4), B);
        cudaStream_t s = as_stream(stream);
        const Crop cr = make_crop(H, W);
#define MVAE_FWD(CC)                                                                                                        \
        if (out) MVAE_CUDA(launch_pdl(recon_loss_fwd_flat_kernel<CC, true>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, out, sums, H, W, a, bb, v0, v1, cr));          \
        else     MVAE_CUDA(launch_pdl(recon_loss_fwd_flat_kernel<CC, false>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, out, sums, H, W, a, bb, v0, v1, cr))
        if (C == 1) { MVAE_FWD(1); } else if (C == 2) { MVAE_FWD(2); } else if (C == 3) { MVAE_FWD(3); } else { MVAE_FWD(4); }
#undef MVAE_FWD
        MVAE_LAUNCH_CHECK();
        return MVAE_OK;
    }
    dim3 grid(loss_grid_x(B, H * W), B);
    MVAE_CUDA(launch_pdl(recon_loss_fwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), r0, y, out, sums, H, W, C, a, bb, v0, v1, make_crop(H, W)));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_recon_loss_bwd(const float* r0, const float* y, const float* sums, float* dr0, int B, int H, int W,
                                   int C, float v0, float v1, float r_scale, mvae_stream_t stream) {
    MVAE_REQUIRE(r0 && y && sums && dr0 && B > 0 && H > 1 && W > 1, "recon_loss_bwd: bad arguments");
    MVAE_REQUIRE(C >= 1 && C <= kMaxC, "recon_loss_bwd: C=%d unsupported (max %d)", C, kMaxC);
    MVAE_REQUIRE(B <= 65535, "recon_loss_bwd: B too large");
    const float a = (v1 - v0) * 0.5f, bb = (v1 - v0) * 0.5f + v0;
    if ((W % 4) == 0 && al16(r0) && al16(y) && al16(dr0) && ((long long)H * W * C) % 4 == 0) {
        dim3 grid(flat_grid_x(B, (long long)H * W * C / 4), B);
        cudaStream_t s = as_stream(stream);
        const Crop cr = make_crop(H, W);
        if (C == 1)      MVAE_CUDA(launch_pdl(recon_loss_bwd_flat_kernel<1>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, sums, dr0, H, W, a, bb, v0, v1, r_scale, cr));
        else if (C == 2) MVAE_CUDA(launch_pdl(recon_loss_bwd_flat_kernel<2>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, sums, dr0, H, W, a, bb, v0, v1, r_scale, cr));
        else if (C == 3) MVAE_CUDA(launch_pdl(recon_loss_bwd_flat_kernel<3>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, sums, dr0, H, W, a, bb, v0, v1, r_scale, cr));
        else             MVAE_CUDA(launch_pdl(recon_loss_bwd_flat_kernel<4>, dim3(grid), dim3(kFlatThreads), 0, s, r0, y, sums, dr0, H, W, a, bb, v0, v1, r_scale, cr));
        MVAE_LAUNCH_CHECK();
        return MVAE_OK;
    }
    dim3 grid(loss_grid_x(B, H * W), B);
    MVAE_CUDA(launch_pdl(recon_loss_bwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), r0, y, sums, dr0, H, W, C, a, bb, v0, v1, r_scale,
                                                             make_crop(H, W)));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_loss_finalize(const float* sums, const float* kl, int levels, float* per_sample, float* scalars,
                                  int B, int H, int W, int C, float r_factor, float kl_factor, mvae_stream_t stream) {
    MVAE_REQUIRE(sums && kl && per_sample && scalars && B > 0 && levels > 0, "loss_finalize: bad arguments");
    const Crop c = make_crop(H, W);
    MVAE_CUDA(launch_pdl(loss_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), sums, kl, levels, per_sample, scalars, B, H * W, C,
                                                         (c.r1 - c.r0) * (c.c1 - c.c0), r_factor, kl_factor));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}
