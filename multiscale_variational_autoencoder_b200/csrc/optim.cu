// Optimiser: Keras kernel regularisers ('l1'/'l2', factor 0.01) folded into the gradient, per-variable clipnorm and
// Adagrad (initial accumulator 0.1 is set by the host, epsilon 1e-7).  Reference: multiscale_vae.py:497-499 and the
// kernel_regularizer arguments at multiscale_vae.py:340,362,368,406,430 / layer_blocks.py:14,900,943.
#include "common.cuh"

namespace mvae {

constexpr float kRegFactor = 0.01f;

struct Seg { long long offset, count, width, ld, reg; };

__device__ __forceinline__ long long seg_addr(const Seg& s, long long j) {
    if (s.count < (1LL << 31)) {                        // 32-bit division: the 64-bit one made these kernels compute-bound
        const unsigned int w = (unsigned int)s.width, jj = (unsigned int)j;
        const unsigned int r = jj / w;
        return s.offset + (long long)r * s.ld + (jj - r * w);
    }
    const long long r = j / s.width;
    return s.offset + r * s.ld + (j - r * s.width);
}

__global__ void __launch_bounds__(256) optim_norms_kernel(const float* __restrict__ params, float* __restrict__ grads,
                                                          const long long* __restrict__ segs,
                                                          const long long* __restrict__ chunks, int chunk_elems,
                                                          float grad_scale, float* __restrict__ partials) {
    pdl_sync();
    const long long sid = chunks[2 * blockIdx.x], start = chunks[2 * blockIdx.x + 1];
    Seg s;
    s.offset = segs[5 * sid]; s.count = segs[5 * sid + 1]; s.width = segs[5 * sid + 2]; s.ld = segs[5 * sid + 3];
    s.reg = segs[5 * sid + 4];
    const long long end = min(s.count, start + (long long)chunk_elems);
    float ss = 0.f, rl = 0.f;
    const bool dense = s.ld == s.width;                 // contiguous variable: no 64-bit division per element
    for (long long j = start + threadIdx.x; j < end; j += blockDim.x) {
        const long long a = dense ? s.offset + j : seg_addr(s, j);
        const float w = params[a];
        float g = grads[a] * grad_scale;
        if (s.reg == MVAE_REG_L1) {
            g += kRegFactor * ((float)(w > 0.f) - (float)(w < 0.f));
            rl += kRegFactor * fabsf(w);
        } else if (s.reg == MVAE_REG_L2) {
            g = fmaf(2.f * kRegFactor, w, g);
            rl = fmaf(kRegFactor * w, w, rl);
        }
        grads[a] = g;
        ss = fmaf(g, g, ss);
    }
    __shared__ float red[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ss = warp_sum(ss); rl = warp_sum(rl);
    if (lane == 0) { red[warp][0] = ss; red[warp][1] = rl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += red[i][0]; b += red[i][1]; }
        // one partial per chunk, summed in a fixed order by optim_finalize_kernel: the clip factors (and with them the
        // updated weights) do not depend on the order in which the chunks' CTAs finish, so data-parallel replicas that
        // apply the same summed gradient stay bit-identical
        partials[2 * blockIdx.x] = a;
        partials[2 * blockIdx.x + 1] = b;
    }
}

// block s < nseg: sumsq[s] = sum of the partials of segment s's chunks (contiguous in the chunk table);  block nseg: reg_loss
__global__ void __launch_bounds__(256) optim_finalize_kernel(const long long* __restrict__ chunks, int nchunk, int nseg,
                                                             const float* __restrict__ partials, float* __restrict__ sumsq,
                                                             float* __restrict__ reg_loss) {
    pdl_sync();
    const int sid = blockIdx.x;
    int lo = 0, hi = nchunk, which = 1;
    if (sid < nseg) {
        which = 0;
        // first chunk of segment sid, and of sid + 1 (chunks are sorted by segment)
        int a = 0, b = nchunk;
        while (a < b) { const int m = (a + b) >> 1; if (chunks[2 * m] < sid) a = m + 1; else b = m; }
        lo = a; b = nchunk;
        while (a < b) { const int m = (a + b) >> 1; if (chunks[2 * m] <= sid) a = m + 1; else b = m; }
        hi = a;
    }
    float v = 0.f;
    for (int c = lo + threadIdx.x; c < hi; c += 256) v += partials[2 * c + which];
    __shared__ float red[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        if (which == 0) sumsq[sid] = t; else *reg_loss += t;
    }
}

__global__ void __launch_bounds__(256) optim_adagrad_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                            float* __restrict__ acc, const long long* __restrict__ segs,
                                                            const long long* __restrict__ chunks, int chunk_elems,
                                                            const float* __restrict__ sumsq, const float* __restrict__ lr,
                                                            float clip_norm, float eps) {
    pdl_sync();
    const long long sid = chunks[2 * blockIdx.x], start = chunks[2 * blockIdx.x + 1];
    Seg s;
    s.offset = segs[5 * sid]; s.count = segs[5 * sid + 1]; s.width = segs[5 * sid + 2]; s.ld = segs[5 * sid + 3];
    s.reg = segs[5 * sid + 4];
    const long long end = min(s.count, start + (long long)chunk_elems);
    float scale = 1.f;
    if (clip_norm > 0.f) scale = clip_norm / fmaxf(sqrtf(sumsq[sid]), clip_norm);   // tf.clip_by_norm
    const float rate = *lr;
    const bool dense = s.ld == s.width;
    for (long long j = start + threadIdx.x; j < end; j += blockDim.x) {
        const long long a = dense ? s.offset + j : seg_addr(s, j);
        const float g = grads[a] * scale;
        const float ac = fmaf(g, g, acc[a]);
        acc[a] = ac;
        params[a] -= rate * g / (sqrtf(ac) + eps);
    }
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_optim_norms(const float* params, float* grads, const long long* segs, const long long* chunks,
                                int nseg, int nchunk, int chunk_elems, float grad_scale, float* partials, float* sumsq,
                                float* reg_loss, mvae_stream_t stream) {
    MVAE_REQUIRE(params && grads && segs && chunks && partials && sumsq && reg_loss && nseg > 0 && nchunk > 0 && chunk_elems > 0,
                 "optim_norms: bad arguments");
    MVAE_CUDA(launch_pdl(optim_norms_kernel, dim3(nchunk), dim3(256), 0, as_stream(stream), params, grads, segs, chunks, chunk_elems, grad_scale,
                         partials));
    MVAE_LAUNCH_CHECK();
    MVAE_CUDA(launch_pdl(optim_finalize_kernel, dim3(nseg + 1), dim3(256), 0, as_stream(stream), chunks, nchunk, nseg,
                         (const float*)partials, sumsq, reg_loss));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_optim_adagrad(float* params, const float* grads, float* acc, const long long* segs,
                                  const long long* chunks, int nchunk, int chunk_elems, const float* sumsq,
                                  const float* lr, float clip_norm, float eps, mvae_stream_t stream) {
    MVAE_REQUIRE(params && grads && acc && segs && chunks && sumsq && lr && nchunk > 0 && chunk_elems > 0,
                 "optim_adagrad: bad arguments");
    MVAE_CUDA(launch_pdl(optim_adagrad_kernel, dim3(nchunk), dim3(256), 0, as_stream(stream), params, grads, acc, segs, chunks, chunk_elems, sumsq, lr,
                                                              clip_norm, eps));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}
