// mobilenetV3 internals: depthwise 3x3 (+ReLU, +GAP partial sums), squeeze-excite gate (with batch-stat BatchNorm),
// and the decoder tail BatchNorm + 1x1 conv.  Reference: layer_blocks.py:418-462, 604-623; multiscale_vae.py:420-431.
#include "common.cuh"

namespace mvae {

// ---------------------------------------------------------------------------------------------------------
// Channel-group helpers: a thread owns VW consecutive channels (VW = 4 when C % 4 == 0 else 1).
// ---------------------------------------------------------------------------------------------------------
template <int VW> struct Vec;
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ static Vec<4> load(const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ static Vec<1> load(const float* p) { Vec<1> r; r.v[0] = __ldg(p); return r; }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

// grid = (chunks, B); a block walks pixels of ONE image; thread t < PPB*CQ: lp = t / CQ, cq = t % CQ
template <int VW>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ u,
                                                     float* __restrict__ gap_sum, int H, int W, int C) {
    pdl_sync();
    const int CQ = C / VW, PPB = 256 / CQ;
    const int t = threadIdx.x, b = blockIdx.y;
    const bool active = t < PPB * CQ;
    const int lp = active ? t / CQ : 0, cq = active ? t % CQ : 0;
    const int c0 = cq * VW;
    float wr[9][VW], br[VW], gs[VW];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) wr[k][v] = __ldg(w + k * C + c0 + v);
#pragma unroll
    for (int v = 0; v < VW; ++v) { br[v] = bias ? __ldg(bias + c0 + v) : 0.f; gs[v] = 0.f; }
    const long long img = (long long)b * H * W * C;
    const int npix = H * W;
    if (active) {
        for (int p = blockIdx.x * PPB + lp; p < npix; p += gridDim.x * PPB) {
            const int y = p / W, x = p - y * W;
            float acc[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v) acc[v] = br[v];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const Vec<VW> av = Vec<VW>::load(a + img + ((long long)yy * W + xx) * C + c0);
#pragma unroll
                    for (int v = 0; v < VW; ++v) acc[v] = fmaf(av.v[v], wr[ky * 3 + kx][v], acc[v]);
                }
            }
            Vec<VW> o;
#pragma unroll
            for (int v = 0; v < VW; ++v) { o.v[v] = fmaxf(acc[v], 0.f); gs[v] += o.v[v]; }
            o.store(u + img + (long long)p * C + c0);
        }
    }
    if (gap_sum == nullptr) return;
    __shared__ float red[256][VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) red[t][v] = active ? gs[v] : 0.f;
    __syncthreads();
    if (t < CQ) {
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            float s = 0.f;
            for (int l = 0; l < PPB; ++l) s += red[l * CQ + t][v];
            atomicAdd(gap_sum + (long long)b * C + t * VW + v, s);
        }
    }
}

// d_pre(q) = (gate*dv(q) + dgap) * (u(q) > 0)
// da(q)  = ( sum_k d_pre(q - off_k) * w_k ) * (a(q) > 0)
// dw_k  += sum_q a(q + off_k) * d_pre(q) ;  dbias += sum_q d_pre(q)
template <int VW>
__global__ void __launch_bounds__(256) dw_bwd_kernel(const float* __restrict__ a, const float* __restrict__ u,
                                                     const float* __restrict__ dv, const float* __restrict__ gate,
                                                     const float* __restrict__ dgap, const float* __restrict__ w,
                                                     float* __restrict__ da, float* __restrict__ dw,
                                                     float* __restrict__ dbias, int H, int W, int C) {
    pdl_sync();
    const int CQ = C / VW, PPB = 256 / CQ;
    const int t = threadIdx.x, b = blockIdx.y;
    const bool active = t < PPB * CQ;
    const int lp = active ? t / CQ : 0, cq = active ? t % CQ : 0;
    const int c0 = cq * VW;
    float wr[9][VW], gt[VW], dg[VW], gw[10][VW];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) { wr[k][v] = __ldg(w + k * C + c0 + v); gw[k][v] = 0.f; }
#pragma unroll
    for (int v = 0; v < VW; ++v) {
        gt[v] = __ldg(gate + (long long)b * C + c0 + v);
        dg[v] = __ldg(dgap + (long long)b * C + c0 + v);
        gw[9][v] = 0.f;
    }
    const long long img = (long long)b * H * W * C;
    const int npix = H * W;
    if (active) {
        for (int p = blockIdx.x * PPB + lp; p < npix; p += gridDim.x * PPB) {
            const int y = p / W, x = p - y * W;
            float acc[VW], dc[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v) acc[v] = 0.f;
            {   // centre d_pre
                const Vec<VW> uv = Vec<VW>::load(u + img + (long long)p * C + c0);
                const Vec<VW> gv = Vec<VW>::load(dv + img + (long long)p * C + c0);
#pragma unroll
                for (int v = 0; v < VW; ++v) { dc[v] = uv.v[v] > 0.f ? fmaf(gt[v], gv.v[v], dg[v]) : 0.f; gw[9][v] += dc[v]; }
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    // data gradient: neighbour q - off = (y - (ky-1), x - (kx-1))
                    const int yy = y - (ky - 1), xx = x - (kx - 1);
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                        const long long o = img + ((long long)yy * W + xx) * C + c0;
                        const Vec<VW> uv = Vec<VW>::load(u + o);
                        const Vec<VW> gv = Vec<VW>::load(dv + o);
#pragma unroll
                        for (int v = 0; v < VW; ++v) {
                            const float dp = uv.v[v] > 0.f ? fmaf(gt[v], gv.v[v], dg[v]) : 0.f;
                            acc[v] = fmaf(dp, wr[ky * 3 + kx][v], acc[v]);
                        }
                    }
                    // weight gradient: a(q + off) * d_pre(q)
                    const int ya = y + (ky - 1), xa = x + (kx - 1);
                    if (ya >= 0 && ya < H && xa >= 0 && xa < W) {
                        const Vec<VW> av = Vec<VW>::load(a + img + ((long long)ya * W + xa) * C + c0);
#pragma unroll
                        for (int v = 0; v < VW; ++v) gw[ky * 3 + kx][v] = fmaf(av.v[v], dc[v], gw[ky * 3 + kx][v]);
                    }
                }
            }
            const Vec<VW> ac = Vec<VW>::load(a + img + (long long)p * C + c0);
            Vec<VW> o;
#pragma unroll
            for (int v = 0; v < VW; ++v) o.v[v] = ac.v[v] > 0.f ? acc[v] : 0.f;
            o.store(da + img + (long long)p * C + c0);
        }
    }
    __shared__ float red[256][10 * VW + 1];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) red[t][k * VW + v] = active ? gw[k][v] : 0.f;
    __syncthreads();
    // outputs: (k, cq, v) -> 10 * C values, each summed over PPB threads
    for (int o = t; o < 10 * C; o += 256) {
        const int k = o / C, c = o - k * C;
        const int q = c / VW, v = c - q * VW;
        float s = 0.f;
        for (int l = 0; l < PPB; ++l) s += red[l * CQ + q][k * VW + v];
        if (k < 9) atomicAdd(dw + k * C + c, s);
        else if (dbias) atomicAdd(dbias + c, s);
    }
}

// several problems (pyramid levels) per launch: grid.x ranges
constexpr int kMaxBatch = 8;
struct DgP { const float* dv; const float* u; float* dg; int HW, x_begin, x_count; };
struct DgBatch { DgP p[kMaxBatch]; int n, C; };

template <int VW>
__global__ void __launch_bounds__(256) dgate_reduce_kernel(const __grid_constant__ DgBatch bt) {
    pdl_sync();
    int lvl = 0;
    while (lvl + 1 < bt.n && (int)blockIdx.x >= bt.p[lvl + 1].x_begin) ++lvl;
    const DgP& pr = bt.p[lvl];
    const float* __restrict__ dv = pr.dv; const float* __restrict__ u = pr.u; float* __restrict__ dg = pr.dg;
    const int HW = pr.HW, C = bt.C;
    const int bx = blockIdx.x - pr.x_begin, gx = pr.x_count;
    const int CQ = C / VW, PPB = 256 / CQ;
    const int t = threadIdx.x, b = blockIdx.y;
    const bool active = t < PPB * CQ;
    const int lp = active ? t / CQ : 0, cq = active ? t % CQ : 0;
    const int c0 = cq * VW;
    const long long img = (long long)b * HW * C;
    float s[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) s[v] = 0.f;
    if (active) {
        for (int p = bx * PPB + lp; p < HW; p += gx * PPB) {
            const Vec<VW> a = Vec<VW>::load(dv + img + (long long)p * C + c0);
            if (u) {
                const Vec<VW> bq = Vec<VW>::load(u + img + (long long)p * C + c0);
#pragma unroll
                for (int v = 0; v < VW; ++v) s[v] = fmaf(a.v[v], bq.v[v], s[v]);
            } else {
#pragma unroll
                for (int v = 0; v < VW; ++v) s[v] += a.v[v];
            }
        }
    }
    __shared__ float red[256][VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) red[t][v] = active ? s[v] : 0.f;
    __syncthreads();
    if (t < CQ) {
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            float r = 0.f;
            for (int l = 0; l < PPB; ++l) r += red[l * CQ + t][v];
            atomicAdd(dg + (long long)b * C + t * VW + v, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Decoder tail
// ---------------------------------------------------------------------------------------------------------
template <int VW>
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                       long long M, int C) {
    pdl_sync();
    const int CQ = C / VW, PPB = 256 / CQ;
    const int t = threadIdx.x;
    const bool active = t < PPB * CQ;
    const int lp = active ? t / CQ : 0, cq = active ? t % CQ : 0;
    const int c0 = cq * VW;
    float s[VW], q[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) s[v] = q[v] = 0.f;
    if (active) {
        // four grid-stride iterations at a time: their loads are in flight together (one load per trip left the kernel at the
        // latency of a global load per 16 bytes and thread: 18 us for 33 MB); the sums are taken in the same order as before
        const long long st = (long long)gridDim.x * PPB;
        long long p = (long long)blockIdx.x * PPB + lp;
        for (; p + 3 * st < M; p += 4 * st) {
            Vec<VW> a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = Vec<VW>::load(x + (p + u * st) * C + c0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < VW; ++v) { s[v] += a[u].v[v]; q[v] = fmaf(a[u].v[v], a[u].v[v], q[v]); }
        }
        for (; p < M; p += st) {
            const Vec<VW> a = Vec<VW>::load(x + p * C + c0);
#pragma unroll
            for (int v = 0; v < VW; ++v) { s[v] += a.v[v]; q[v] = fmaf(a.v[v], a.v[v], q[v]); }
        }
    }
    __shared__ float red[256][2 * VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) { red[t][v] = active ? s[v] : 0.f; red[t][VW + v] = active ? q[v] : 0.f; }
    __syncthreads();
    if (t < CQ) {
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            double rs = 0.0, rq = 0.0;
            for (int l = 0; l < PPB; ++l) { rs += red[l * CQ + t][v]; rq += red[l * CQ + t][VW + v]; }
            atomicAdd(sums + t * VW + v, rs);
            atomicAdd(sums + C + t * VW + v, rq);
        }
    }
}

// out[c] += sum_p x[p][c]   (bias gradient of Conv2DTranspose)
template <int VW>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int C) {
    pdl_sync();
    const int CQ = C / VW, PPB = 256 / CQ;
    const int t = threadIdx.x;
    const bool active = t < PPB * CQ;
    const int lp = active ? t / CQ : 0, cq = active ? t % CQ : 0;
    const int c0 = cq * VW;
    float s[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) s[v] = 0.f;
    if (active) {
        const long long st = (long long)gridDim.x * PPB;
        long long p = (long long)blockIdx.x * PPB + lp;
        for (; p + 3 * st < M; p += 4 * st) {             // four loads in flight per thread (see bn_stats_kernel)
            Vec<VW> a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = Vec<VW>::load(x + (p + u * st) * C + c0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < VW; ++v) s[v] += a[u].v[v];
        }
        for (; p < M; p += st) {
            const Vec<VW> a = Vec<VW>::load(x + p * C + c0);
#pragma unroll
            for (int v = 0; v < VW; ++v) s[v] += a.v[v];
        }
    }
    __shared__ float red[256][VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) red[t][v] = active ? s[v] : 0.f;
    __syncthreads();
    if (t < CQ) {
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            float r = 0.f;
            for (int l = 0; l < PPB; ++l) r += red[l * CQ + t][v];
            atomicAdd(out + t * VW + v, r);
        }
    }
}

// y[b,p,c] = x[b,p,c] * gate[b,c]   (keras Multiply of squeeze_excite_block, layer_blocks.py:458-460, standalone use)
__global__ void __launch_bounds__(256) channel_scale_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                            float* __restrict__ y, long long total, int HWC, int C) {
    pdl_sync();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / HWC;
        y[i] = __ldg(x + i) * __ldg(gate + b * C + (int)(i % C));
    }
}

constexpr int kMaxCo = 4;

// y[p][co] = sum_f BN(x)[p][f] w[f][co] + b[co], BN folded into effective weights held in shared memory.
// C % 4 == 0 and C/4 a power of two <= 32: the C/4 lanes of a pixel reduce with shuffles.
__global__ void __launch_bounds__(256) bn_convout_fwd_kernel(const float* __restrict__ x, const double* __restrict__ sums,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ y, float* __restrict__ stats, long long M,
                                                             int C, int Co, float eps, float momentum, int training) {
    pdl_sync();
    extern __shared__ float sm[];
    float* weff = sm;                 // [C][kMaxCo]
    float* beff = sm + C * kMaxCo;    // [kMaxCo]
    float* mu = beff + kMaxCo;        // [C]
    float* rs = mu + C;               // [C]
    const int t = threadIdx.x;
    for (int f = t; f < C; f += blockDim.x) {
        float m, var;
        if (training) {
            const double dm = sums[f] / (double)M;
            double dvv = sums[C + f] / (double)M - dm * dm;
            if (dvv < 0.0) dvv = 0.0;
            m = (float)dm; var = (float)dvv;
            if (blockIdx.x == 0) {
                const float unb = (M > 1) ? (float)(dvv * (double)M / (double)(M - 1)) : var;
                moving_mean[f] = moving_mean[f] * momentum + m * (1.f - momentum);
                moving_var[f] = moving_var[f] * momentum + unb * (1.f - momentum);
            }
        } else {
            m = moving_mean[f]; var = moving_var[f];
        }
        mu[f] = m; rs[f] = rsqrtf(var + eps);
        if (blockIdx.x == 0) { stats[f] = m; stats[C + f] = rs[f]; }
    }
    __syncthreads();
    for (int i = t; i < C * kMaxCo; i += blockDim.x) {
        const int f = i / kMaxCo, co = i - f * kMaxCo;
        weff[i] = co < Co ? __ldg(gamma + f) * rs[f] * __ldg(w + f * Co + co) : 0.f;
    }
    __syncthreads();
    if (t < kMaxCo) {
        float acc = (t < Co && bias) ? __ldg(bias + t) : 0.f;
        if (t < Co)
            for (int f = 0; f < C; ++f) acc += (__ldg(beta + f) - __ldg(gamma + f) * mu[f] * rs[f]) * __ldg(w + f * Co + t);
        beff[t] = acc;
    }
    __syncthreads();
    const int CQ = C >> 2, PPB = 256 / CQ;
    const int lp = t / CQ, cq = t % CQ;
    // four pixel groups per trip: the four 16-byte loads of a thread are in flight together
    const long long st = (long long)gridDim.x * PPB;
    for (long long p0 = (long long)blockIdx.x * PPB; p0 < M; p0 += 4 * st) {
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long p = p0 + u * st + lp;
            xv[u] = p < M ? __ldg(reinterpret_cast<const float4*>(x + p * C + 4 * cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long p = p0 + u * st + lp;
            if (p0 + u * st >= M) break;                   // uniform over the block
            float acc[kMaxCo];
            const float* we = weff + 4 * cq * kMaxCo;
#pragma unroll
            for (int co = 0; co < kMaxCo; ++co)
                acc[co] = xv[u].x * we[co] + xv[u].y * we[kMaxCo + co] + xv[u].z * we[2 * kMaxCo + co] + xv[u].w * we[3 * kMaxCo + co];
            for (int o = CQ >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int co = 0; co < kMaxCo; ++co) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
            }
            if (cq == 0 && p < M) {
#pragma unroll
                for (int co = 0; co < kMaxCo; ++co) if (co < Co) y[p * Co + co] = acc[co] + beff[co];
            }
        }
    }
}

// pass 1: red[f][co] += sum_p xh[p][f] dy[p][co];  red[C*Co + co] += sum_p dy[p][co]
__global__ void __launch_bounds__(256) bn_convout_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    const float* __restrict__ stats, float* __restrict__ red,
                                                                    long long M, int C, int Co) {
    pdl_sync();
    const int CQ = C >> 2, PPB = 256 / CQ;
    const int t = threadIdx.x, lp = t / CQ, cq = t % CQ;
    float mu[4], rs[4], acc[4][kMaxCo], sdy[kMaxCo];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        mu[v] = stats[4 * cq + v]; rs[v] = stats[C + 4 * cq + v];
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) acc[v][co] = 0.f;
    }
#pragma unroll
    for (int co = 0; co < kMaxCo; ++co) sdy[co] = 0.f;
    const long long st = (long long)gridDim.x * PPB;
    for (long long pb = (long long)blockIdx.x * PPB + lp; pb < M; pb += 4 * st) {
        // four pixels per trip, every load of the four in flight before the first use (same summation order as one at a time)
        float4 xq[4];
        float dq[4][kMaxCo];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long p = pb + u * st;
            const bool in = p < M;
            xq[u] = in ? __ldg(reinterpret_cast<const float4*>(x + p * C + 4 * cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int co = 0; co < kMaxCo; ++co) dq[u][co] = (in && co < Co) ? __ldg(dy + p * Co + co) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (pb + u * st >= M) break;
            const float4 xv = xq[u];
            const float xh[4] = {(xv.x - mu[0]) * rs[0], (xv.y - mu[1]) * rs[1], (xv.z - mu[2]) * rs[2], (xv.w - mu[3]) * rs[3]};
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int co = 0; co < kMaxCo; ++co) acc[v][co] = fmaf(xh[v], dq[u][co], acc[v][co]);
            if (cq == 0) {
#pragma unroll
                for (int co = 0; co < kMaxCo; ++co) sdy[co] += dq[u][co];
            }
        }
    }
    __shared__ float sred[256][4 * kMaxCo + 1];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) sred[t][v * kMaxCo + co] = acc[v][co];
    __syncthreads();
    for (int o = t; o < C * Co; o += 256) {
        const int f = o / Co, co = o - f * Co;
        const int q = f >> 2, v = f & 3;
        float s = 0.f;
        for (int l = 0; l < PPB; ++l) s += sred[l * CQ + q][v * kMaxCo + co];
        atomicAdd(red + o, s);
    }
    __syncthreads();
    if (cq == 0) {
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) sred[lp][co] = sdy[co];
    }
    __syncthreads();
    if (t < Co) {
        float s = 0.f;
        for (int l = 0; l < PPB; ++l) s += sred[l][t];
        atomicAdd(red + C * Co + t, s);
    }
}

// pass 2: dx = gamma*rstd*(dbn - dbeta/M - xh*dgamma/M), dbn[p][f] = sum_co dy[p][co] w[f][co]; block 0 adds param grads
__global__ void __launch_bounds__(256) bn_convout_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ w,
                                                                   const float* __restrict__ red, float* __restrict__ dx,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   float* __restrict__ dw, float* __restrict__ dbias,
                                                                   long long M, int C, int Co) {
    pdl_sync();
    extern __shared__ float sm[];
    float* dgm = sm;          // [C] dgamma / M
    float* dbt = sm + C;      // [C] dbeta / M
    const int t = threadIdx.x;
    const float inv_m = 1.f / (float)M;
    for (int f = t; f < C; f += blockDim.x) {
        float g = 0.f, bsum = 0.f;
        for (int co = 0; co < Co; ++co) {
            const float wv = __ldg(w + f * Co + co);
            g = fmaf(wv, red[f * Co + co], g);
            bsum = fmaf(wv, red[C * Co + co], bsum);
        }
        dgm[f] = g * inv_m; dbt[f] = bsum * inv_m;
        if (blockIdx.x == 0) {
            atomicAdd(dgamma + f, g);
            atomicAdd(dbeta + f, bsum);
            for (int co = 0; co < Co; ++co)
                atomicAdd(dw + f * Co + co, __ldg(gamma + f) * red[f * Co + co] + __ldg(beta + f) * red[C * Co + co]);
        }
    }
    if (blockIdx.x == 0 && t < Co && dbias) atomicAdd(dbias + t, red[C * Co + t]);
    __syncthreads();
    const int CQ = C >> 2;
    const long long total = M * CQ;
    for (long long i = (long long)blockIdx.x * blockDim.x + t; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / CQ;
        const int cq = (int)(i - p * CQ);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + p * C + 4 * cq));
        float d[kMaxCo];
#pragma unroll
        for (int co = 0; co < kMaxCo; ++co) d[co] = co < Co ? __ldg(dy + p * Co + co) : 0.f;
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int f = 4 * cq + v;
            const float rs = stats[C + f];
            const float xh = (xs[v] - stats[f]) * rs;
            float dbn = 0.f;
#pragma unroll
            for (int co = 0; co < kMaxCo; ++co) if (co < Co) dbn = fmaf(d[co], __ldg(w + f * Co + co), dbn);
            o[v] = __ldg(gamma + f) * rs * (dbn - dbt[f] - xh * dgm[f]);
        }
        *reinterpret_cast<float4*>(dx + p * C + 4 * cq) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

static inline int img_grid_x(int B, int npix, int ppb) {
    int gx = ceil_div((long long)kNumSMs * 8, B);
    const int maxx = ceil_div(npix, ppb);
    if (gx > maxx) gx = maxx;
    return gx < 1 ? 1 : gx;
}

static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace mvae

using namespace mvae;

#define MVAE_VW_DISPATCH(C, ptr_ok, CALL4, CALL1) \
    do {                                          \
        if (((C) % 4) == 0 && (ptr_ok)) { CALL4; } else { CALL1; } \
    } while (0)

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

namespace mvae {
// dwconv_tiled.cu: shared-memory tiled variants; MVAE_ERR_UNSUPPORTED when the shape is not covered
int dw_fwd_tiled(const float* a, const float* w, const float* bias, float* u, float* gap_sum, int B, int H, int W, int C,
                 cudaStream_t s);
int dw_bwd_tiled(const float* a, const float* u, const float* dv, const float* gate, const float* dgap, const float* w,
                 float* da, float* dw, float* dbias, int B, int H, int W, int C, cudaStream_t s);
int dw_fwd_tiled_batched(int n, const float* const* a, const float* const* w, const float* const* bias, float* const* u,
                         float* const* gap_sum, int B, const int* H, const int* W, int C, cudaStream_t s);
int dw_bwd_tiled_batched(int n, const float* const* a, const float* const* u, const float* const* dv,
                         const float* const* gate, const float* const* dgap, const float* const* w, float* const* da,
                         float* const* dw, float* const* dbias, int B, const int* H, const int* W, int C, cudaStream_t s);
}  // namespace mvae

extern "C" int mvae_dwconv3x3_fwd(const float* a, const float* w, const float* bias, float* u, float* gap_sum, int B, int H,
                                  int W, int C, mvae_stream_t stream);
extern "C" int mvae_dwconv3x3_bwd(const float* a, const float* u, const float* dv, const float* gate, const float* dgap,
                                  const float* w, float* da, float* dw, float* dbias, int B, int H, int W, int C,
                                  mvae_stream_t stream);

extern "C" int mvae_dwconv3x3_fwd_batched(int n, const float* const* a, const float* const* w, const float* const* bias,
                                          float* const* u, float* const* gap_sum, int B, const int* H, const int* W, int C,
                                          mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && a && w && u && H && W && B > 0 && C > 0, "dwconv3x3_fwd_batched: bad arguments");
    if (n <= 8) {
        const int r = dw_fwd_tiled_batched(n, a, w, bias, u, gap_sum, B, H, W, C, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    for (int l = 0; l < n; ++l)
        if (int e = mvae_dwconv3x3_fwd(a[l], w[l], bias ? bias[l] : nullptr, u[l], gap_sum ? gap_sum[l] : nullptr, B, H[l], W[l],
                                       C, stream))
            return e;
    return MVAE_OK;
}

extern "C" int mvae_dwconv3x3_bwd_batched(int n, const float* const* a, const float* const* u, const float* const* dv,
                                          const float* const* gate, const float* const* dgap, const float* const* w,
                                          float* const* da, float* const* dw, float* const* dbias, int B, const int* H,
                                          const int* W, int C, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && a && u && dv && gate && dgap && w && da && dw && H && W && B > 0 && C > 0,
                 "dwconv3x3_bwd_batched: bad arguments");
    if (n <= 8) {
        const int r = dw_bwd_tiled_batched(n, a, u, dv, gate, dgap, w, da, dw, dbias, B, H, W, C, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    for (int l = 0; l < n; ++l)
        if (int e = mvae_dwconv3x3_bwd(a[l], u[l], dv[l], gate[l], dgap[l], w[l], da[l], dw[l], dbias ? dbias[l] : nullptr, B,
                                       H[l], W[l], C, stream))
            return e;
    return MVAE_OK;
}

extern "C" int mvae_dwconv3x3_fwd(const float* a, const float* w, const float* bias, float* u, float* gap_sum, int B,
                                  int H, int W, int C, mvae_stream_t stream) {
    MVAE_REQUIRE(a && w && u && B > 0 && H > 0 && W > 0 && C > 0 && B <= 65535, "dwconv3x3_fwd: bad arguments");
    MVAE_REQUIRE(C <= 256, "dwconv3x3_fwd: C=%d unsupported (max 256 channel groups)", C);
    cudaStream_t s = as_stream(stream);
    {
        const int r = dw_fwd_tiled(a, w, bias, u, gap_sum, B, H, W, C, s);
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    const bool v4 = (C % 4) == 0 && al16(a) && al16(u);
    const int cq = v4 ? C / 4 : C;
    dim3 grid(img_grid_x(B, H * W, 256 / cq), B);
    if (v4) MVAE_CUDA(launch_pdl(dw_fwd_kernel<4>, dim3(grid), dim3(256), 0, s, a, w, bias, u, gap_sum, H, W, C));
    else    MVAE_CUDA(launch_pdl(dw_fwd_kernel<1>, dim3(grid), dim3(256), 0, s, a, w, bias, u, gap_sum, H, W, C));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_dwconv3x3_bwd(const float* a, const float* u, const float* dv, const float* gate, const float* dgap,
                                  const float* w, float* da, float* dw, float* dbias, int B, int H, int W, int C,
                                  mvae_stream_t stream) {
    MVAE_REQUIRE(a && u && dv && gate && dgap && w && da && dw && B > 0 && H > 0 && W > 0 && C > 0 && B <= 65535,
                 "dwconv3x3_bwd: bad arguments");
    MVAE_REQUIRE(C <= 256, "dwconv3x3_bwd: C=%d unsupported", C);
    cudaStream_t s = as_stream(stream);
    {
        const int r = dw_bwd_tiled(a, u, dv, gate, dgap, w, da, dw, dbias, B, H, W, C, s);
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    const bool v4 = (C % 4) == 0 && al16(a) && al16(u) && al16(dv) && al16(da);
    const int cq = v4 ? C / 4 : C;
    dim3 grid(img_grid_x(B, H * W, 256 / cq), B);
    if (v4) MVAE_CUDA(launch_pdl(dw_bwd_kernel<4>, dim3(grid), dim3(256), 0, s, a, u, dv, gate, dgap, w, da, dw, dbias, H, W, C));
    else    MVAE_CUDA(launch_pdl(dw_bwd_kernel<1>, dim3(grid), dim3(256), 0, s, a, u, dv, gate, dgap, w, da, dw, dbias, H, W, C));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_se_dgate_reduce_batched(int n, const float* const* dv, const float* const* u, float* const* dg, int B,
                                            const int* HW, int C, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && dv && dg && HW && B > 0 && C > 0 && C <= 256 && B <= 65535, "se_dgate_reduce: bad arguments");
    cudaStream_t s = as_stream(stream);
    for (int l0 = 0; l0 < n; l0 += kMaxBatch) {
        DgBatch bt;
        bt.n = n - l0 < kMaxBatch ? n - l0 : kMaxBatch;
        bt.C = C;
        bool v4 = (C % 4) == 0;
        for (int l = 0; l < bt.n; ++l) v4 = v4 && al16(dv[l0 + l]) && (!u || u[l0 + l] == nullptr || al16(u[l0 + l]));
        const int cq = v4 ? C / 4 : C;
        int gx = 0;
        for (int l = 0; l < bt.n; ++l) {
            const int k = l0 + l;
            MVAE_REQUIRE(dv[k] && dg[k] && HW[k] > 0, "se_dgate_reduce: bad member %d", k);
            const int cnt = img_grid_x(B, HW[k], 256 / cq);
            bt.p[l] = DgP{dv[k], u ? u[k] : nullptr, dg[k], HW[k], gx, cnt};
            gx += cnt;
        }
        if (v4) MVAE_CUDA(launch_pdl(dgate_reduce_kernel<4>, dim3(gx, B), dim3(256), 0, s, bt));
        else    MVAE_CUDA(launch_pdl(dgate_reduce_kernel<1>, dim3(gx, B), dim3(256), 0, s, bt));
        MVAE_LAUNCH_CHECK();
    }
    return MVAE_OK;
}

extern "C" int mvae_se_dgate_reduce(const float* dv, const float* u, float* dg, int B, int HW, int C,
                                    mvae_stream_t stream) {
    return mvae_se_dgate_reduce_batched(1, &dv, u ? &u : nullptr, &dg, B, &HW, C, stream);
}

extern "C" int mvae_colsum(const float* x, float* out, long long M, int C, mvae_stream_t stream) {
    MVAE_REQUIRE(x && out && M > 0 && C > 0 && C <= 256, "colsum: bad arguments");
    cudaStream_t s = as_stream(stream);
    const bool v4 = (C % 4) == 0 && al16(x);
    const int cq = v4 ? C / 4 : C;
    const int ppb = 256 / cq;
    int grid = ceil_div(M, (long long)ppb * 8);
    if (grid > kNumSMs * 2) grid = kNumSMs * 2;      // every CTA ends in same-address atomics: few CTAs, many loads each
    if (grid < 1) grid = 1;
    if (v4) MVAE_CUDA(launch_pdl(colsum_kernel<4>, dim3(grid), dim3(256), 0, s, x, out, M, C));
    else    MVAE_CUDA(launch_pdl(colsum_kernel<1>, dim3(grid), dim3(256), 0, s, x, out, M, C));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_channel_scale(const float* x, const float* gate, float* y, int B, int HW, int C, mvae_stream_t stream) {
    MVAE_REQUIRE(x && gate && y && B > 0 && HW > 0 && C > 0, "channel_scale: bad arguments");
    const long long total = (long long)B * HW * C;
    long long g = (total + 255) / 256;
    if (g > kNumSMs * 16) g = kNumSMs * 16;
    MVAE_CUDA(launch_pdl(channel_scale_kernel, dim3((int)g), dim3(256), 0, as_stream(stream), x, gate, y, total, HW * C, C));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_bn_stats(const float* x, double* stat_sums, long long M, int Cf, mvae_stream_t stream) {
    MVAE_REQUIRE(x && stat_sums && M > 0 && Cf > 0 && Cf <= 256, "bn_stats: bad arguments");
    cudaStream_t s = as_stream(stream);
    const bool v4 = (Cf % 4) == 0 && al16(x);
    const int cq = v4 ? Cf / 4 : Cf;
    const int ppb = 256 / cq;
    int grid = ceil_div(M, (long long)ppb * 8);
    if (grid > kNumSMs * 2) grid = kNumSMs * 2;      // every CTA ends in same-address atomics: few CTAs, many loads each
    if (grid < 1) grid = 1;
    if (v4) MVAE_CUDA(launch_pdl(bn_stats_kernel<4>, dim3(grid), dim3(256), 0, s, x, stat_sums, M, Cf));
    else    MVAE_CUDA(launch_pdl(bn_stats_kernel<1>, dim3(grid), dim3(256), 0, s, x, stat_sums, M, Cf));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

static int check_tail(long long M, int Cf, int Co) {
    MVAE_REQUIRE(M > 0 && Co >= 1 && Co <= kMaxCo, "bn_convout: Co=%d unsupported (max %d)", Co, kMaxCo);
    MVAE_REQUIRE((Cf % 4) == 0 && pow2(Cf / 4) && Cf / 4 <= 32, "bn_convout: filters=%d must be 4*2^k <= 128", Cf);
    return MVAE_OK;
}

extern "C" int mvae_bn_convout_fwd(const float* x, const double* stat_sums, const float* gamma, const float* beta,
                                   float* moving_mean, float* moving_var, const float* w, const float* bias, float* y,
                                   float* stats, long long M, int Cf, int Co, float eps, float momentum, int training,
                                   mvae_stream_t stream) {
    MVAE_REQUIRE(x && gamma && beta && moving_mean && moving_var && w && y && stats, "bn_convout_fwd: null pointer");
    MVAE_REQUIRE(!training || stat_sums, "bn_convout_fwd: stat_sums required when training");
    if (int e = check_tail(M, Cf, Co)) return e;
    MVAE_REQUIRE(al16(x), "bn_convout_fwd: x must be 16-byte aligned");
    const int ppb = 256 / (Cf / 4);
    int grid = ceil_div(M, (long long)ppb * 4);
    // every CTA rebuilds the folded weights from the batch sums (three barriers, ~100 dependent global loads): few CTAs with
    // long pixel loops (MVAE_TAIL_CTAS per SM, default 2) instead of 8 per SM with seven trips each
    static int per_sm = 0;
    if (!per_sm) per_sm = env_int("MVAE_TAIL_CTAS", 2);
    if (grid > kNumSMs * per_sm) grid = kNumSMs * per_sm;
    const size_t smem = (size_t)(Cf * kMaxCo + kMaxCo + 2 * Cf) * sizeof(float);
    MVAE_CUDA(launch_pdl(bn_convout_fwd_kernel, dim3(grid), dim3(256), smem, as_stream(stream), x, stat_sums, gamma, beta, moving_mean, moving_var, w,
                                                                bias, y, stats, M, Cf, Co, eps, momentum, training));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_bn_convout_bwd(const float* x, const float* dy, const float* stats, const float* gamma,
                                   const float* beta, const float* w, float* red, float* dx, float* dgamma,
                                   float* dbeta, float* dw, float* dbias, long long M, int Cf, int Co,
                                   mvae_stream_t stream) {
    MVAE_REQUIRE(x && dy && stats && gamma && beta && w && red && dx && dgamma && dbeta && dw, "bn_convout_bwd: null pointer");
    if (int e = check_tail(M, Cf, Co)) return e;
    MVAE_REQUIRE(al16(x) && al16(dx), "bn_convout_bwd: x/dx must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int ppb = 256 / (Cf / 4);
    int grid = ceil_div(M, (long long)ppb * 8);
    if (grid > kNumSMs * 2) grid = kNumSMs * 2;      // every CTA ends in same-address atomics: few CTAs, many loads each
    MVAE_CUDA(launch_pdl(bn_convout_bwd_reduce_kernel, dim3(grid), dim3(256), 0, s, x, dy, stats, red, M, Cf, Co));
    MVAE_LAUNCH_CHECK();
    int grid2 = ceil_div(M * (Cf / 4), 256 * 4);
    static int per_sm2 = 0;
    if (!per_sm2) per_sm2 = env_int("MVAE_TAIL_APPLY_CTAS", 4);
    if (grid2 > kNumSMs * per_sm2) grid2 = kNumSMs * per_sm2;
    MVAE_CUDA(launch_pdl(bn_convout_bwd_apply_kernel, dim3(grid2), dim3(256), 2 * Cf * sizeof(float), s, x, dy, stats, gamma, beta, w, red, dx, dgamma,
                                                                         dbeta, dw, dbias, M, Cf, Co));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}
