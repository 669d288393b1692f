// Depthwise 3x3 (+bias +ReLU +GAP partial sums) and its backward, shared-memory tiled (layer_blocks.py:604-614).
// A CTA owns TH rows x the full width of ONE image: the tile (+1 halo row above / below, +1 zero pixel left / right) is
// staged once with 16-byte loads (NHWC rows are contiguous runs of W*C floats), then every thread owns one (pixel column,
// 4-channel group) and walks down the rows with a rolling 3x3 window in registers: 3 shared-memory loads per output instead
// of 9 global ones, no per-pixel index arithmetic, per-thread accumulators for the GAP / weight-gradient sums.
// The generic per-pixel kernels in blocks.cu remain the fallback (C % 4 != 0, tiny or very wide images).
#include "common.cuh"

namespace mvae {
namespace dwt {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 relu4(float4 a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// smem tile: (TH + 2) rows x (W + 2) pixels x C floats; rows outside the image and the two pad pixels are zero
template <bool DPRE>
__device__ __forceinline__ void stage_tile(float* S, const float* __restrict__ src, const float* __restrict__ u,
                                           const float* __restrict__ gate_b, const float* __restrict__ dgap_b, int y0,
                                           int x0, int TH, int TW, int H, int W, int C) {
    // staged pixel (r, j) == image pixel (y0 - 1 + r, x0 - 1 + j); zero outside the image
    const int cq = C / 4, rowq = (TW + 2) * cq;
    const int RS = (TW + 2) * C;
    for (int i = threadIdx.x; i < (TH + 2) * rowq; i += blockDim.x) {
        const int r = i / rowq, q = i - r * rowq;
        const int j = q / cq, c4 = (q - j * cq) * 4;
        const int y = y0 - 1 + r, x = x0 - 1 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const long long o = ((long long)y * W + x) * C + c4;
            v = __ldg(reinterpret_cast<const float4*>(src + o));
            if (DPRE) {
                // d_pre = (gate * dv + dgap) * (u > 0)
                const float4 uv = __ldg(reinterpret_cast<const float4*>(u + o));
                const float4 g = ld4(gate_b + c4), d = ld4(dgap_b + c4);
                v.x = uv.x > 0.f ? fmaf(g.x, v.x, d.x) : 0.f; v.y = uv.y > 0.f ? fmaf(g.y, v.y, d.y) : 0.f;
                v.z = uv.z > 0.f ? fmaf(g.z, v.z, d.z) : 0.f; v.w = uv.w > 0.f ? fmaf(g.w, v.w, d.w) : 0.f;
            }
        }
        *reinterpret_cast<float4*>(S + r * RS + 4 * q) = v;
    }
}

// Several independent problems (pyramid levels: same C, different H x W and weights) share a launch: grid.x ranges.
constexpr int kMaxBatch = 8;
struct FwdP { const float* a; const float* w; const float* bias; float* u; float* gap_sum; int H, W, TH, TW, ctiles, x_begin; };
struct BwdP { const float* a; const float* u; const float* dv; const float* gate; const float* dgap; const float* w;
              float* da; float* dw; float* dbias; int H, W, TH, TW, ctiles, x_begin; };
struct FwdBatch { FwdP p[kMaxBatch]; int n, C; };
struct BwdBatch { BwdP p[kMaxBatch]; int n, C; };

// grid = (sum over problems of tiles per image, B); blockDim = ncols * C/4 threads
__global__ void __launch_bounds__(512) dw_fwd_tiled_kernel(const __grid_constant__ FwdBatch bt) {
    pdl_sync();
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.p[lvl + 1].x_begin) ++lvl;
    const FwdP& pr = bt.p[lvl];
    const float* __restrict__ a = pr.a; const float* __restrict__ w = pr.w; const float* __restrict__ bias = pr.bias;
    float* __restrict__ u = pr.u; float* __restrict__ gap_sum = pr.gap_sum;
    const int H = pr.H, W = pr.W, C = bt.C, TH = pr.TH, TW = pr.TW;
    const int bx = blockIdx.x - pr.x_begin;
    extern __shared__ __align__(16) float sm[];
    const int cqn = C / 4, RS = (TW + 2) * C;
    const int rt = bx / pr.ctiles, ct = bx - rt * pr.ctiles;
    const int b = blockIdx.y, y0 = rt * TH, x0 = ct * TW;
    const int th = min(TH, H - y0), tw = min(TW, W - x0);
    const long long img = (long long)b * H * W * C;
    const int cq = threadIdx.x % cqn, col0 = threadIdx.x / cqn, ncols = blockDim.x / cqn;
    float4 wr[9], br;
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w + k * C) + cq);
    br = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    stage_tile<false>(sm, a + img, nullptr, nullptr, nullptr, y0, x0, TH, TW, H, W, C);
    __syncthreads();
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int xl = col0; xl < tw; xl += ncols) {
        const int x = x0 + xl;
        const float* base = sm + xl * C + cq * 4;         // staged pixel column xl == image column x - 1
        float4 win[3][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) win[r][k] = ld4(base + r * RS + k * C);
        float* out = u + img + ((long long)y0 * W + x) * C + cq * 4;
        for (int ry = 0; ry < th; ++ry) {
#pragma unroll
            for (int k = 0; k < 3; ++k) win[2][k] = ld4(base + (ry + 2) * RS + k * C);
            float4 acc = br;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) acc = fma4(win[ky][kx], wr[ky * 3 + kx], acc);
            acc = relu4(acc);
            gs = add4(gs, acc);
            *reinterpret_cast<float4*>(out + (long long)ry * W * C) = acc;
#pragma unroll
            for (int k = 0; k < 3; ++k) { win[0][k] = win[1][k]; win[1][k] = win[2][k]; }
        }
    }
    if (gap_sum == nullptr) return;
    __syncthreads();                                       // the tile is dead: reuse it for the column reduction
    *reinterpret_cast<float4*>(sm + col0 * C + cq * 4) = gs;
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float s = 0.f;
        for (int l = 0; l < ncols; ++l) s += sm[l * C + threadIdx.x];
        atomicAdd(gap_sum + (long long)b * C + threadIdx.x, s);
    }
}

// d_pre(q) = (gate*dv(q) + dgap) * (u(q) > 0);  da(q) = (sum_k d_pre(q - off_k) w_k) * (a(q) > 0)
// dw_k += sum_q a(q + off_k) d_pre(q)  ==  sum_q' a(q') d_pre(q' - off_k)   (zero padding: the out-of-image terms of both
// forms vanish), so the SAME 3x3 window of d_pre serves the data gradient and the weight gradient and `a` is only needed
// at the centre pixel: one staged tile (d_pre), `a` read straight from global memory one row ahead.
// dbias += sum_q d_pre(q)
__global__ void __launch_bounds__(256, 2) dw_bwd_tiled_kernel(const __grid_constant__ BwdBatch bt) {
    pdl_sync();
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.p[lvl + 1].x_begin) ++lvl;
    const BwdP& pr = bt.p[lvl];
    const float* __restrict__ a = pr.a; const float* __restrict__ u = pr.u; const float* __restrict__ dv = pr.dv;
    const float* __restrict__ gate = pr.gate; const float* __restrict__ dgap = pr.dgap; const float* __restrict__ w = pr.w;
    float* __restrict__ da = pr.da; float* __restrict__ dw = pr.dw; float* __restrict__ dbias = pr.dbias;
    const int H = pr.H, W = pr.W, C = bt.C, TH = pr.TH, TW = pr.TW;
    const int bx = blockIdx.x - pr.x_begin;
    extern __shared__ __align__(16) float sm[];
    const int cqn = C / 4, RS = (TW + 2) * C;
    float* SP = sm;                                 // d_pre tile (+1 halo row / zero pixel on every side)
    const int rt = bx / pr.ctiles, ct = bx - rt * pr.ctiles;
    const int b = blockIdx.y, y0 = rt * TH, x0 = ct * TW;
    const int th = min(TH, H - y0), tw = min(TW, W - x0);
    const long long img = (long long)b * H * W * C;
    const int cq = threadIdx.x % cqn, col0 = threadIdx.x / cqn, ncols = blockDim.x / cqn;
    float4 wr[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w + k * C) + cq);
    stage_tile<true>(SP, dv + img, u + img, gate + (long long)b * C, dgap + (long long)b * C, y0, x0, TH, TW, H, W, C);
    __syncthreads();
    float4 gw[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int xl = col0; xl < tw; xl += ncols) {
        const int x = x0 + xl;
        const float* bp = SP + xl * C + cq * 4;
        float4 wp[3][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) wp[r][k] = ld4(bp + r * RS + k * C);
        const float* ain = a + img + ((long long)y0 * W + x) * C + cq * 4;
        float* out = da + img + ((long long)y0 * W + x) * C + cq * 4;
        float4 ac = __ldg(reinterpret_cast<const float4*>(ain));
        for (int ry = 0; ry < th; ++ry) {
            float4 an = ac;
            if (ry + 1 < th) an = __ldg(reinterpret_cast<const float4*>(ain + (long long)(ry + 1) * W * C));   // next row's centre
#pragma unroll
            for (int k = 0; k < 3; ++k) wp[2][k] = ld4(bp + (ry + 2) * RS + k * C);
            // window entry [2-ky][2-kx] = d_pre(q - off_k)
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    acc = fma4(wp[2 - ky][2 - kx], wr[ky * 3 + kx], acc);
                    gw[ky * 3 + kx] = fma4(ac, wp[2 - ky][2 - kx], gw[ky * 3 + kx]);
                }
            gw[9] = add4(gw[9], wp[1][1]);
            acc.x = ac.x > 0.f ? acc.x : 0.f; acc.y = ac.y > 0.f ? acc.y : 0.f;
            acc.z = ac.z > 0.f ? acc.z : 0.f; acc.w = ac.w > 0.f ? acc.w : 0.f;
            *reinterpret_cast<float4*>(out + (long long)ry * W * C) = acc;
#pragma unroll
            for (int k = 0; k < 3; ++k) { wp[0][k] = wp[1][k]; wp[1][k] = wp[2][k]; }
            ac = an;
        }
    }
    __syncthreads();                                       // the tile is dead: reuse shared memory for the reduction
    // red[col][k][C]
#pragma unroll
    for (int k = 0; k < 10; ++k) *reinterpret_cast<float4*>(sm + ((col0 * 10 + k) * C) + cq * 4) = gw[k];
    __syncthreads();
    for (int o = threadIdx.x; o < 10 * C; o += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < ncols; ++l) s += sm[l * 10 * C + o];
        const int k = o / C, c = o - k * C;
        if (k < 9) atomicAdd(dw + k * C + c, s);
        else if (dbias) atomicAdd(dbias + c, s);
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// tile = TH rows x TW columns (+1 halo each side) within ~72 KB of shared memory (three CTAs per SM): full image rows when
// they fit with at least 4 rows, otherwise the width is halved until they do
static void plan_tile(int H, int W, int C, int& TH, int& TW) {
    int tw = W;
    while (tw > 8 && (size_t)6 * (tw + 2) * C * 4 > 72 * 1024) tw = (tw + 1) / 2;
    int th = H;
    while (th > 4 && (size_t)(th + 2) * (tw + 2) * C * 4 > 72 * 1024) th = (th + 1) / 2;
    TH = th; TW = tw;
}

// one thread count for the whole launch: as many pixel columns as the widest problem has (capped), times C/4 channel groups
static bool plan_threads(int n, const int* W, int C, int max_threads, int& threads) {
    if ((C % 4) || C > 512) return false;
    const int cqn = C / 4;
    int wmax = 0;
    for (int l = 0; l < n; ++l) wmax = W[l] > wmax ? W[l] : wmax;      // W: tile widths
    int ncols = wmax;
    while (ncols * cqn > max_threads) ncols = (ncols + 1) / 2;
    if (ncols < 1) return false;
    threads = ncols * cqn;
    if (threads < 32) threads = 32 / cqn * cqn > 0 ? ((32 + cqn - 1) / cqn) * cqn : cqn;
    return threads <= max_threads && threads >= C;      // the final reductions use one thread per channel
}

}  // namespace dwt

int dw_fwd_tiled_batched(int n, const float* const* a, const float* const* w, const float* const* bias, float* const* u,
                         float* const* gap_sum, int B, const int* H, const int* W, int C, cudaStream_t s) {
    if (n < 1 || n > dwt::kMaxBatch || B > 65535) return MVAE_ERR_UNSUPPORTED;
    int threads, THs[dwt::kMaxBatch], TWs[dwt::kMaxBatch];
    if ((C % 4) || C > 512) return MVAE_ERR_UNSUPPORTED;
    for (int l = 0; l < n; ++l) dwt::plan_tile(H[l], W[l], C, THs[l], TWs[l]);
    if (!dwt::plan_threads(n, TWs, C, 512, threads)) return MVAE_ERR_UNSUPPORTED;
    dwt::FwdBatch bt;
    bt.n = n; bt.C = C;
    size_t smem = 0;
    int gx = 0;
    for (int l = 0; l < n; ++l) {
        const float* bi = bias ? bias[l] : nullptr;
        if (!dwt::al16(a[l]) || !dwt::al16(u[l]) || !dwt::al16(w[l]) || (bi && !dwt::al16(bi))) return MVAE_ERR_UNSUPPORTED;
        const int TH = THs[l], TW = TWs[l];
        size_t sm = (size_t)(TH + 2) * (TW + 2) * C * 4;
        const size_t red = (size_t)(threads / (C / 4)) * C * 4;
        if (red > sm) sm = red;
        if (sm > smem) smem = sm;
        const int ctiles = (W[l] + TW - 1) / TW;
        bt.p[l] = dwt::FwdP{a[l], w[l], bi, u[l], gap_sum ? gap_sum[l] : nullptr, H[l], W[l], TH, TW, ctiles, gx};
        gx += ((H[l] + TH - 1) / TH) * ctiles;
    }
    if (smem > 200 * 1024) return MVAE_ERR_UNSUPPORTED;
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(dwt::dw_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    MVAE_CUDA(launch_pdl(dwt::dw_fwd_tiled_kernel, dim3(gx, B), dim3(threads), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

int dw_bwd_tiled_batched(int n, const float* const* a, const float* const* u, const float* const* dv,
                         const float* const* gate, const float* const* dgap, const float* const* w, float* const* da,
                         float* const* dw, float* const* dbias, int B, const int* H, const int* W, int C, cudaStream_t s) {
    if (n < 1 || n > dwt::kMaxBatch || B > 65535) return MVAE_ERR_UNSUPPORTED;
    int threads, THs[dwt::kMaxBatch], TWs[dwt::kMaxBatch];
    if ((C % 4) || C > 512) return MVAE_ERR_UNSUPPORTED;
    for (int l = 0; l < n; ++l) dwt::plan_tile(H[l], W[l], C, THs[l], TWs[l]);
    if (!dwt::plan_threads(n, TWs, C, 256, threads)) return MVAE_ERR_UNSUPPORTED;
    dwt::BwdBatch bt;
    bt.n = n; bt.C = C;
    size_t smem = 0;
    int gx = 0;
    for (int l = 0; l < n; ++l) {
        if (!dwt::al16(a[l]) || !dwt::al16(u[l]) || !dwt::al16(dv[l]) || !dwt::al16(da[l]) || !dwt::al16(w[l]) ||
            !dwt::al16(gate[l]) || !dwt::al16(dgap[l]))
            return MVAE_ERR_UNSUPPORTED;
        const int TH = THs[l], TW = TWs[l];
        size_t sm = (size_t)(TH + 2) * (TW + 2) * C * 4;
        const size_t red = (size_t)(threads / (C / 4)) * 10 * C * 4;
        if (red > sm) sm = red;
        if (sm > smem) smem = sm;
        const int ctiles = (W[l] + TW - 1) / TW;
        bt.p[l] = dwt::BwdP{a[l], u[l], dv[l], gate[l], dgap[l], w[l], da[l], dw[l], dbias ? dbias[l] : nullptr, H[l], W[l], TH, TW,
                            ctiles, gx};
        gx += ((H[l] + TH - 1) / TH) * ctiles;
    }
    if (smem > 200 * 1024) return MVAE_ERR_UNSUPPORTED;
    static DeviceOnce configured;
    if (configured.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(dwt::dw_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    MVAE_CUDA(launch_pdl(dwt::dw_bwd_tiled_kernel, dim3(gx, B), dim3(threads), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

int dw_fwd_tiled(const float* a, const float* w, const float* bias, float* u, float* gap_sum, int B, int H, int W, int C,
                 cudaStream_t s) {
    if (W * (C / 4) < 64) return MVAE_ERR_UNSUPPORTED;       // a lone tiny image row: the per-pixel kernel is as good
    return dw_fwd_tiled_batched(1, &a, &w, bias ? &bias : nullptr, &u, gap_sum ? &gap_sum : nullptr, B, &H, &W, C, s);
}

int dw_bwd_tiled(const float* a, const float* u, const float* dv, const float* gate, const float* dgap, const float* w,
                 float* da, float* dw, float* dbias, int B, int H, int W, int C, cudaStream_t s) {
    if (W * (C / 4) < 64) return MVAE_ERR_UNSUPPORTED;
    return dw_bwd_tiled_batched(1, &a, &u, &dv, &gate, &dgap, &w, &da, &dw, dbias ? &dbias : nullptr, B, &H, &W, C, s);
}

}  // namespace mvae
