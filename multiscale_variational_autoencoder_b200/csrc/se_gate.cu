// Squeeze-excite gate (layer_blocks.py:418-462, use_batchnorm=True):
//     gate = hard_sigmoid( BN_batch( relu(gap W0 + b0) ) W1 + b1 ),   gap = mean_hw(u)
// The problem is tiny (B*C values, two CxC matrices) but sits on the critical path of every mobilenetV3 block, forward
// and backward, and the batch-statistics BatchNorm couples all samples.  One thread-block CLUSTER of 8 CTAs splits the
// batch; the per-channel batch sums cross CTAs through a small global scratch + barrier.cluster (release/acquire).
// ws layout (floats, n = B*C): gap[n] h1[n] hn[n] s[n] (unused 2n) mean[C] rstd[C] partF[8*2*C] partB[8*2*C]
#include "common.cuh"

namespace mvae {

extern long long* g_trace;        // debug timeline buffer (conv_tc.cu, mvae_debug_trace)
__device__ __forceinline__ void se_trace(long long* tr, int ev) {
    if (tr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[1 + 3 * ev] = ev; tr[2 + 3 * ev] = 0; tr[3 + 3 * ev] = (long long)t;
        if (tr[0] < ev + 1) tr[0] = ev + 1;
    }
}

constexpr int kCS = 8;            // CTAs per cluster
constexpr int kSeThreads = 512;

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ int cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return (int)r;
}
// read a float of CTA `rank`'s shared memory (distributed shared memory): same offset as `local` in the peer
__device__ __forceinline__ float dsmem_ld(const float* local, int rank) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(local), raddr;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(addr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
    return v;
}

// One cluster per problem: a launch serves the same block of every pyramid level (grid = 8 x n CTAs).
constexpr int kMaxBatch = 8;
struct FwdP { const float* gap_sum; const float* w0; const float* b0; const float* gamma; const float* beta; const float* w1;
              const float* b1; float* moving_mean; float* moving_var; float* gate; float* ws; float inv_hw; };
struct BwdP { const float* dg; const float* w0; const float* gamma; const float* beta; const float* w1; float* ws; float* dgap;
              float* dw0; float* db0; float* dgamma; float* dbeta; float* dw1; float* db1; float inv_hw; };
struct FwdBatch { FwdP p[kMaxBatch]; int n, B, C, training; float eps, momentum; long long* trace; };
struct BwdBatch { BwdP p[kMaxBatch]; int n, B, C; long long* trace; };

// ws layout (floats, n = B*C): gap[n] h1[n] (unused n) s[n] (unused 2n) mean[C] rstd[C]
//
// Forward.  Everything the CTA needs (both weight matrices, the vectors, its slice of the GAP sums) is fetched up front with
// 16-byte loads -- one exposed global latency -- and lives in shared memory afterwards.  The batch statistics cross the
// cluster ONCE, through distributed shared memory: each CTA publishes (mean_i, M2_i) of its samples, every CTA combines the
// eight pairs with Chan's parallel-variance formula (exactly the biased batch variance, no E[x^2] - mean^2 cancellation).
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kSeThreads)
se_gate_fwd_kernel(const __grid_constant__ FwdBatch bt) {
    pdl_sync();
    se_trace(bt.trace, 0);
    const FwdP& pr = bt.p[blockIdx.x / kCS];
    const int B = bt.B, C = bt.C, training = bt.training;
    const float inv_hw = pr.inv_hw, eps = bt.eps, momentum = bt.momentum;
    extern __shared__ __align__(16) float sm[];
    const int rank = cluster_rank();
    const int per = (B + kCS - 1) / kCS;
    const int bs = min(B, rank * per), be = min(B, bs + per);
    const int nb = be - bs, nloc = nb * C, NL = per * C;
    const long long n = (long long)B * C;
    float* W0 = sm; float* W1 = W0 + C * C; float* vb0 = W1 + C * C; float* vb1 = vb0 + C; float* vg = vb1 + C; float* vbe = vg + C;
    float* gap = vbe + C; float* h1 = gap + NL; float* hn = h1 + NL;
    float* part = hn + NL;                         // [2][C]: mean_i, M2_i of this CTA's samples (read by the peers)
    float* mean = part + 2 * C; float* rstd = mean + C;
    float* g_gap = pr.ws + (long long)bs * C; float* g_h1 = pr.ws + n + (long long)bs * C;
    float* g_sp = pr.ws + 3 * n + (long long)bs * C; float* g_stat = pr.ws + 6 * n;
    const int t = threadIdx.x, nt = blockDim.x;
    const int warp = t >> 5, lane = t & 31, nw = nt >> 5;

    const bool v4 = (C & 3) == 0;
    if (v4) {
        for (int i = t; i < C * C / 4; i += nt) {
            reinterpret_cast<float4*>(W0)[i] = __ldg(reinterpret_cast<const float4*>(pr.w0) + i);
            reinterpret_cast<float4*>(W1)[i] = __ldg(reinterpret_cast<const float4*>(pr.w1) + i);
        }
    } else {
        for (int i = t; i < C * C; i += nt) { W0[i] = __ldg(pr.w0 + i); W1[i] = __ldg(pr.w1 + i); }
    }
    for (int i = t; i < C; i += nt) { vb0[i] = __ldg(pr.b0 + i); vb1[i] = __ldg(pr.b1 + i); vg[i] = __ldg(pr.gamma + i); vbe[i] = __ldg(pr.beta + i); }
    for (int i = t; i < nloc; i += nt) { const float v = __ldg(pr.gap_sum + (long long)bs * C + i) * inv_hw; gap[i] = v; g_gap[i] = v; }
    __syncthreads();
    se_trace(bt.trace, 1);
    // dense0 + relu: consecutive lanes = consecutive output channels (W0 row reads conflict-free, gap reads broadcast)
    for (int i = t; i < nloc; i += nt) {
        const int b = i / C, j = i - b * C;
        float a0 = vb0[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* gr = gap + b * C;
        int c = 0;
        for (; c + 3 < C; c += 4) {
            a0 = fmaf(gr[c], W0[c * C + j], a0);
            a1 = fmaf(gr[c + 1], W0[(c + 1) * C + j], a1);
            a2 = fmaf(gr[c + 2], W0[(c + 2) * C + j], a2);
            a3 = fmaf(gr[c + 3], W0[(c + 3) * C + j], a3);
        }
        for (; c < C; ++c) a0 = fmaf(gr[c], W0[c * C + j], a0);
        const float v = fmaxf((a0 + a1) + (a2 + a3), 0.f);
        h1[i] = v; g_h1[i] = v;
    }
    __syncthreads();
    se_trace(bt.trace, 2);
    if (training) {
        for (int j = warp; j < C; j += nw) {
            float sacc = 0.f;
            for (int b = lane; b < nb; b += 32) sacc += h1[b * C + j];
            const float mi = nb > 0 ? warp_sum(sacc) / (float)nb : 0.f;
            float q = 0.f;
            for (int b = lane; b < nb; b += 32) { const float d = h1[b * C + j] - mi; q = fmaf(d, d, q); }
            q = warp_sum(q);
            if (lane == 0) { part[j] = mi; part[C + j] = q; }
        }
        se_trace(bt.trace, 3);
        cluster_arrive();
        cluster_wait();                            // every CTA's (mean_i, M2_i) is published
        se_trace(bt.trace, 4);
        for (int j = t; j < C; j += nt) {
            float mi[kCS], qi[kCS], m = 0.f;
#pragma unroll
            for (int r = 0; r < kCS; ++r) {
                mi[r] = dsmem_ld(part + j, r); qi[r] = dsmem_ld(part + C + j, r);
                const int nr = min(B, (r + 1) * per) - min(B, r * per);
                m = fmaf((float)nr, mi[r], m);
            }
            m /= (float)B;
            float M2 = 0.f;
#pragma unroll
            for (int r = 0; r < kCS; ++r) {
                const int nr = min(B, (r + 1) * per) - min(B, r * per);
                const float d = mi[r] - m;
                M2 += qi[r] + (float)nr * d * d;
            }
            const float var = M2 / (float)B;
            mean[j] = m;
            rstd[j] = rsqrtf(var + eps);
            if (rank == 0) {
                pr.moving_mean[j] = pr.moving_mean[j] * momentum + m * (1.f - momentum);
                pr.moving_var[j] = pr.moving_var[j] * momentum + var * (1.f - momentum);
            }
        }
        se_trace(bt.trace, 5);
        cluster_arrive();                          // this CTA is done reading its peers (waited on before exit)
    } else {
        for (int j = t; j < C; j += nt) { mean[j] = pr.moving_mean[j]; rstd[j] = rsqrtf(pr.moving_var[j] + eps); }
    }
    __syncthreads();
    if (rank == 0) for (int j = t; j < C; j += nt) { g_stat[j] = mean[j]; g_stat[C + j] = rstd[j]; }
    for (int i = t; i < nloc; i += nt) {
        const int j = i % C;
        hn[i] = fmaf(vg[j] * rstd[j], h1[i] - mean[j], vbe[j]);
    }
    __syncthreads();
    for (int i = t; i < nloc; i += nt) {
        const int b = i / C, c = i - b * C;
        float a0 = vb1[c], a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* hr = hn + b * C;
        int j = 0;
        for (; j + 3 < C; j += 4) {
            a0 = fmaf(hr[j], W1[j * C + c], a0);
            a1 = fmaf(hr[j + 1], W1[(j + 1) * C + c], a1);
            a2 = fmaf(hr[j + 2], W1[(j + 2) * C + c], a2);
            a3 = fmaf(hr[j + 3], W1[(j + 3) * C + c], a3);
        }
        for (; j < C; ++j) a0 = fmaf(hr[j], W1[j * C + c], a0);
        const float acc = (a0 + a1) + (a2 + a3);
        g_sp[i] = acc;
        pr.gate[(long long)bs * C + i] = fminf(fmaxf(fmaf(0.2f, acc, 0.5f), 0.f), 1.f);
    }
    se_trace(bt.trace, 6);
    if (training) cluster_wait();                  // peers may still be reading this CTA's `part`
    se_trace(bt.trace, 7);
}

__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kSeThreads)
se_gate_bwd_kernel(const __grid_constant__ BwdBatch bt) {
    pdl_sync();
    se_trace(bt.trace, 0);
    const BwdP& pr = bt.p[blockIdx.x / kCS];
    const float* __restrict__ dg = pr.dg; const float* __restrict__ w0 = pr.w0; const float* __restrict__ gamma = pr.gamma;
    const float* __restrict__ beta = pr.beta; const float* __restrict__ w1 = pr.w1; float* ws = pr.ws;
    float* __restrict__ dgap = pr.dgap; float* __restrict__ dw0 = pr.dw0; float* __restrict__ db0 = pr.db0;
    float* __restrict__ dgamma = pr.dgamma; float* __restrict__ dbeta = pr.dbeta; float* __restrict__ dw1 = pr.dw1;
    float* __restrict__ db1 = pr.db1;
    const int B = bt.B, C = bt.C;
    const float inv_hw = pr.inv_hw;
    extern __shared__ __align__(16) float sm[];
    const int rank = cluster_rank();
    const int per = (B + kCS - 1) / kCS;
    const int bs = min(B, rank * per), be = min(B, bs + per);
    const int nb = be - bs, nloc = nb * C, NL = per * C, CP = C + 1;
    const long long n = (long long)B * C;
    float* gap = sm; float* h1 = sm + NL; float* ds = sm + 2 * NL; float* dh = sm + 3 * NL;
    float* wT1 = sm + 4 * NL;                      // C*(C+1): W1^T
    float* wT0 = wT1 + C * CP;                     // C*(C+1): W0^T
    float* mean = wT0 + C * CP; float* rstd = mean + C; float* sg = rstd + C; float* sb = sg + C;
    float* vg = sb + C; float* vbe = vg + C;
    float* part = vbe + C;                         // [2][C] BatchNorm-backward partial sums, read by the peers
    const float* g_gap = ws + (long long)bs * C; const float* g_h1 = ws + n + (long long)bs * C;
    const float* g_sp = ws + 3 * n + (long long)bs * C;
    const float* g_stat = ws + 6 * n;
    const int t = threadIdx.x, nt = blockDim.x;
    const int warp = t >> 5, lane = t & 31, nw = nt >> 5;

    // everything up front: one exposed global latency
    for (int j = t; j < C; j += nt) { mean[j] = g_stat[j]; rstd[j] = g_stat[C + j]; vg[j] = __ldg(gamma + j); vbe[j] = __ldg(beta + j); }
    for (int i = t; i < nloc; i += nt) {
        gap[i] = g_gap[i]; h1[i] = g_h1[i];
        const float h = fmaf(0.2f, g_sp[i], 0.5f);                 // hard_sigmoid passes where 0 <= h <= 1
        ds[i] = (h >= 0.f && h <= 1.f) ? 0.2f * dg[(long long)bs * C + i] : 0.f;
    }
    for (int o = t; o < C * C; o += nt) {
        const int r = o / C, c = o - r * C;
        wT1[c * CP + r] = __ldg(w1 + o);           // w1[j][c] -> wT1[c][j]
        wT0[c * CP + r] = __ldg(w0 + o);           // w0[c][j] -> wT0[j][c]
    }
    __syncthreads();
    se_trace(bt.trace, 1);
    // dense1: dW1[j][c] += sum_b hn[b][j] ds[b][c]; db1[c] += sum_b ds[b][c]; dhn[b][j] = sum_c ds[b][c] W1[j][c]
    for (int o = t; o < C * C; o += nt) {
        const int j = o / C, c = o - j * C;
        const float gr = vg[j] * rstd[j], mj = mean[j], bj = vbe[j];
        float a0 = 0.f, a1 = 0.f;
        int b = 0;
        for (; b + 1 < nb; b += 2) {
            a0 = fmaf(fmaf(gr, h1[b * C + j] - mj, bj), ds[b * C + c], a0);
            a1 = fmaf(fmaf(gr, h1[(b + 1) * C + j] - mj, bj), ds[(b + 1) * C + c], a1);
        }
        for (; b < nb; ++b) a0 = fmaf(fmaf(gr, h1[b * C + j] - mj, bj), ds[b * C + c], a0);
        if (nb > 0) atomicAdd(dw1 + o, a0 + a1);
    }
    for (int c = warp; c < C; c += nw) {
        float acc = 0.f;
        for (int b = lane; b < nb; b += 32) acc += ds[b * C + c];
        acc = warp_sum(acc);
        if (lane == 0 && nb > 0) atomicAdd(db1 + c, acc);
    }
    for (int i = t; i < nloc; i += nt) {
        const int b = i / C, j = i - b * C;
        float a0 = 0.f, a1 = 0.f;
        int c = 0;
        for (; c + 1 < C; c += 2) {
            a0 = fmaf(ds[b * C + c], wT1[c * CP + j], a0);
            a1 = fmaf(ds[b * C + c + 1], wT1[(c + 1) * CP + j], a1);
        }
        for (; c < C; ++c) a0 = fmaf(ds[b * C + c], wT1[c * CP + j], a0);
        dh[i] = a0 + a1;   // dhn
    }
    __syncthreads();
    se_trace(bt.trace, 2);
    // BatchNorm backward: batch-wide sums of dhn*xh and dhn, exchanged through distributed shared memory
    for (int j = warp; j < C; j += nw) {
        float s0 = 0.f, s1 = 0.f;
        for (int b = lane; b < nb; b += 32) {
            const float d = dh[b * C + j];
            s0 = fmaf(d, (h1[b * C + j] - mean[j]) * rstd[j], s0);
            s1 += d;
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { part[j] = s0; part[C + j] = s1; }
    }
    se_trace(bt.trace, 3);
    cluster_arrive();
    cluster_wait();
    se_trace(bt.trace, 4);
    for (int j = t; j < C; j += nt) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int r = 0; r < kCS; ++r) { s0 += dsmem_ld(part + j, r); s1 += dsmem_ld(part + C + j, r); }
        sg[j] = s0; sb[j] = s1;
        if (rank == 0) { atomicAdd(dgamma + j, s0); atomicAdd(dbeta + j, s1); }
    }
    cluster_arrive();                              // done reading the peers
    __syncthreads();
    const float inv_b = 1.f / (float)B;
    for (int i = t; i < nloc; i += nt) {
        const int j = i % C;
        const float xh = (h1[i] - mean[j]) * rstd[j];
        const float d = vg[j] * rstd[j] * (dh[i] - sb[j] * inv_b - xh * sg[j] * inv_b);
        dh[i] = h1[i] > 0.f ? d : 0.f;      // relu
    }
    __syncthreads();
    se_trace(bt.trace, 5);
    // dense0
    for (int o = t; o < C * C; o += nt) {
        const int c = o / C, j = o - c * C;
        float a0 = 0.f, a1 = 0.f;
        int b = 0;
        for (; b + 1 < nb; b += 2) {
            a0 = fmaf(gap[b * C + c], dh[b * C + j], a0);
            a1 = fmaf(gap[(b + 1) * C + c], dh[(b + 1) * C + j], a1);
        }
        for (; b < nb; ++b) a0 = fmaf(gap[b * C + c], dh[b * C + j], a0);
        if (nb > 0) atomicAdd(dw0 + o, a0 + a1);
    }
    for (int j = warp; j < C; j += nw) {
        float acc = 0.f;
        for (int b = lane; b < nb; b += 32) acc += dh[b * C + j];
        acc = warp_sum(acc);
        if (lane == 0 && nb > 0) atomicAdd(db0 + j, acc);
    }
    se_trace(bt.trace, 6);
    for (int i = t; i < nloc; i += nt) {
        const int b = i / C, c = i - b * C;
        float a0 = 0.f, a1 = 0.f;
        int j = 0;
        for (; j + 1 < C; j += 2) {
            a0 = fmaf(dh[b * C + j], wT0[j * CP + c], a0);
            a1 = fmaf(dh[b * C + j + 1], wT0[(j + 1) * CP + c], a1);
        }
        for (; j < C; ++j) a0 = fmaf(dh[b * C + j], wT0[j * CP + c], a0);
        dgap[(long long)bs * C + i] = (a0 + a1) * inv_hw;
    }
    se_trace(bt.trace, 7);
    cluster_wait();                                // peers may still be reading this CTA's `part`
    se_trace(bt.trace, 8);
}


// ---------------------------------------------------------------------------------------------------------------------
// C = 32 (every mobilenetV3 block of the 32-filter configurations): lane = channel.
// The generic kernels above walk runtime-length loops over shared memory and spend ~10 / 15 us on a few hundred kFLOP of
// latency.  Here a warp owns SPW samples, each lane keeps the weight column (or row) it needs in 32 registers, the
// per-sample vectors sit in registers with a warp-private shared-memory copy for broadcast reads (one LDS.128 feeds four
// FMAs), and the only block-wide steps are the batch-statistics reductions.  Same cluster of 8 CTAs, same ws layout.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kC32Warps = 8;
constexpr int kC32Threads = kC32Warps * 32;

// acc += sum_c row[c] * w[c]   (row: 32 floats in shared memory, read as broadcast float4s)
__device__ __forceinline__ float dot32(const float* row, const float (&w)[32], float init) {
    float a0 = init, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
        a0 = fmaf(v.x, w[4 * q], a0);
        a1 = fmaf(v.y, w[4 * q + 1], a1);
        a2 = fmaf(v.z, w[4 * q + 2], a2);
        a3 = fmaf(v.w, w[4 * q + 3], a3);
    }
    return (a0 + a1) + (a2 + a3);
}

template <int SPW>
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kC32Threads)
se_gate_fwd_c32_kernel(const __grid_constant__ FwdBatch bt) {
    constexpr int C = 32, NW = kC32Warps;
    pdl_launch();        // a programmatic dependent of this launch may begin its own prologue
    const FwdP& pr = bt.p[blockIdx.x / kCS];
    const int B = bt.B, training = bt.training;
    const float inv_hw = pr.inv_hw, eps = bt.eps, momentum = bt.momentum;
    const int rank = cluster_rank();
    const int per = (B + kCS - 1) / kCS;
    const int bs = min(B, rank * per), nb = min(B, bs + per) - bs;
    const long long n = (long long)B * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(16) float rows[NW][SPW][C];
    __shared__ float red[NW][C];
    __shared__ float part[2 * C];                   // (mean_i, M2_i) of this CTA's samples, read by the peers

    float w0c[C], w1c[C];                           // columns `lane` of W0 and W1
#pragma unroll
    for (int c = 0; c < C; ++c) { w0c[c] = __ldg(pr.w0 + c * C + lane); w1c[c] = __ldg(pr.w1 + c * C + lane); }
    const float b0 = __ldg(pr.b0 + lane), b1 = __ldg(pr.b1 + lane), gam = __ldg(pr.gamma + lane), bet = __ldg(pr.beta + lane);
    float mmean = 0.f, mvar = 0.f;
    if (!training || (rank == 0 && warp == 0)) { mmean = pr.moving_mean[lane]; mvar = pr.moving_var[lane]; }
    pdl_wait();          // weights and vectors are in registers: only the GAP sums depend on the predecessor
    float h[SPW];
    bool ok[SPW];
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const int sidx = warp + k * NW;
        ok[k] = sidx < nb;
        const long long off = (long long)(bs + sidx) * C + lane;
        const float v = ok[k] ? __ldg(pr.gap_sum + off) * inv_hw : 0.f;
        rows[warp][k][lane] = v;
        if (ok[k]) pr.ws[off] = v;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        h[k] = fmaxf(dot32(rows[warp][k], w0c, b0), 0.f);
        if (ok[k]) pr.ws[n + (long long)(bs + warp + k * NW) * C + lane] = h[k];
    }
    float mean, rstd;
    if (training) {
        float sl = 0.f;
#pragma unroll
        for (int k = 0; k < SPW; ++k) sl += ok[k] ? h[k] : 0.f;
        red[warp][lane] = sl;
        __syncthreads();
        float mi = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) mi += red[w][lane];
        mi = nb > 0 ? mi / (float)nb : 0.f;
        __syncthreads();
        float ql = 0.f;
#pragma unroll
        for (int k = 0; k < SPW; ++k) { const float d = h[k] - mi; ql = ok[k] ? fmaf(d, d, ql) : ql; }
        red[warp][lane] = ql;
        __syncthreads();
        if (warp == 0) {
            float q = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) q += red[w][lane];
            part[lane] = mi; part[C + lane] = q;
        }
        cluster_arrive();
        cluster_wait();                             // every CTA's (mean_i, M2_i) is published
        float mr[kCS], qr[kCS], m = 0.f;
#pragma unroll
        for (int r = 0; r < kCS; ++r) {
            mr[r] = dsmem_ld(part + lane, r); qr[r] = dsmem_ld(part + C + lane, r);
            const int nr = min(B, (r + 1) * per) - min(B, r * per);
            m = fmaf((float)nr, mr[r], m);
        }
        m /= (float)B;
        float M2 = 0.f;
#pragma unroll
        for (int r = 0; r < kCS; ++r) {
            const int nr = min(B, (r + 1) * per) - min(B, r * per);
            const float d = mr[r] - m;
            M2 += qr[r] + (float)nr * d * d;
        }
        cluster_arrive();                           // done reading the peers (waited on before exit)
        const float var = M2 / (float)B;
        mean = m;
        rstd = rsqrtf(var + eps);
        if (rank == 0 && warp == 0) {
            pr.moving_mean[lane] = mmean * momentum + m * (1.f - momentum);
            pr.moving_var[lane] = mvar * momentum + var * (1.f - momentum);
        }
    } else {
        mean = mmean;
        rstd = rsqrtf(mvar + eps);
    }
    if (rank == 0 && warp == 0) { pr.ws[6 * n + lane] = mean; pr.ws[6 * n + C + lane] = rstd; }
    const float gr = gam * rstd;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SPW; ++k) rows[warp][k][lane] = fmaf(gr, h[k] - mean, bet);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const float acc = dot32(rows[warp][k], w1c, b1);
        if (ok[k]) {
            const long long off = (long long)(bs + warp + k * NW) * C + lane;
            pr.ws[3 * n + off] = acc;
            pr.gate[off] = fminf(fmaxf(fmaf(0.2f, acc, 0.5f), 0.f), 1.f);
        }
    }
    if (training) cluster_wait();                   // peers may still be reading this CTA's `part`
}

template <int SPW>
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kC32Threads)
se_gate_bwd_c32_kernel(const __grid_constant__ BwdBatch bt) {
    constexpr int C = 32, NW = kC32Warps;
    pdl_launch();
    const BwdP& pr = bt.p[blockIdx.x / kCS];
    const int B = bt.B;
    const float inv_hw = pr.inv_hw;
    const int rank = cluster_rank();
    const int per = (B + kCS - 1) / kCS;
    const int bs = min(B, rank * per), nb = min(B, bs + per) - bs;
    const long long n = (long long)B * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    __shared__ __align__(16) float rowsA[NW][SPW][C];
    __shared__ __align__(16) float rowsB[NW][SPW][C];
    __shared__ float redW[NW][C][C];                // per-warp partial weight gradients
    __shared__ float red[3][NW][C];
    __shared__ float part[2 * C];                   // BatchNorm-backward partial sums, read by the peers

    float w1r[C], w0r[C];                           // ROWS `lane` of W1 and W0
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(pr.w1 + lane * C) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(pr.w0 + lane * C) + q);
        w1r[4 * q] = a.x; w1r[4 * q + 1] = a.y; w1r[4 * q + 2] = a.z; w1r[4 * q + 3] = a.w;
        w0r[4 * q] = b.x; w0r[4 * q + 1] = b.y; w0r[4 * q + 2] = b.z; w0r[4 * q + 3] = b.w;
    }
    const float gam = __ldg(pr.gamma + lane), bet = __ldg(pr.beta + lane);
    pdl_wait();          // the weight rows are in registers: the gate-gradient sums (and ws) depend on the predecessor
    const float mean = pr.ws[6 * n + lane], rstd = pr.ws[6 * n + C + lane];
    const float gr = gam * rstd;
    float gp[SPW], h[SPW], ds[SPW];
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const int sidx = warp + k * NW;
        const bool ok = sidx < nb;
        const long long off = (long long)(bs + sidx) * C + lane;
        gp[k] = ok ? pr.ws[off] : 0.f;
        h[k] = ok ? pr.ws[n + off] : 0.f;
        const float hv = ok ? fmaf(0.2f, pr.ws[3 * n + off], 0.5f) : -1.f;       // hard_sigmoid passes where 0 <= hv <= 1
        ds[k] = (hv >= 0.f && hv <= 1.f) ? 0.2f * __ldg(pr.dg + off) : 0.f;
        rowsA[warp][k][lane] = ds[k];
        rowsB[warp][k][lane] = ok ? fmaf(gr, h[k] - mean, bet) : 0.f;             // hn
    }
    __syncwarp();
    // dense1: dhn[b][j] = sum_c ds[b][c] W1[j][c];  dW1[j][c] += sum_b hn[b][j] ds[b][c];  db1[c] += sum_b ds[b][c]
    float dhn[SPW], acc[C];
#pragma unroll
    for (int j = 0; j < C; ++j) acc[j] = 0.f;
    float s0 = 0.f, s1 = 0.f, sd = 0.f;
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        dhn[k] = dot32(rowsA[warp][k], w1r, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(&rowsB[warp][k][4 * q]);
            acc[4 * q] = fmaf(v.x, ds[k], acc[4 * q]);
            acc[4 * q + 1] = fmaf(v.y, ds[k], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v.z, ds[k], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v.w, ds[k], acc[4 * q + 3]);
        }
        s0 = fmaf(dhn[k], (h[k] - mean) * rstd, s0);
        s1 += dhn[k];
        sd += ds[k];
    }
#pragma unroll
    for (int j = 0; j < C; ++j) redW[warp][j][lane] = acc[j];
    red[0][warp][lane] = s0; red[1][warp][lane] = s1; red[2][warp][lane] = sd;
    __syncthreads();
    if (warp == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { a += red[0][w][lane]; b += red[1][w][lane]; c += red[2][w][lane]; }
        part[lane] = a; part[C + lane] = b;
        if (nb > 0) atomicAdd(pr.db1 + lane, c);
    }
    if (nb > 0) {
#pragma unroll
        for (int o = t; o < C * C; o += kC32Threads) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += (&redW[w][0][0])[o];
            atomicAdd(pr.dw1 + o, v);
        }
    }
    cluster_arrive();
    cluster_wait();                                 // BatchNorm-backward partial sums of every CTA are published
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int r = 0; r < kCS; ++r) { sg += dsmem_ld(part + lane, r); sb += dsmem_ld(part + C + lane, r); }
    cluster_arrive();                               // done reading the peers
    if (rank == 0 && warp == 0) { atomicAdd(pr.dgamma + lane, sg); atomicAdd(pr.dbeta + lane, sb); }
    const float inv_b = 1.f / (float)B;
    float dh[SPW];
    float sd0 = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const float xh = (h[k] - mean) * rstd;
        const float d = gr * (dhn[k] - sb * inv_b - xh * sg * inv_b);
        dh[k] = h[k] > 0.f ? d : 0.f;               // relu (h == 0 for the padding samples)
        sd0 += dh[k];
        rowsA[warp][k][lane] = dh[k];
        rowsB[warp][k][lane] = gp[k];
    }
    __syncwarp();
    // dense0: dgap[b][c] = sum_j dh[b][j] W0[c][j];  dW0[c][j] += sum_b gap[b][c] dh[b][j];  db0[j] += sum_b dh[b][j]
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const int sidx = warp + k * NW;
        const float g = dot32(rowsA[warp][k], w0r, 0.f) * inv_hw;
        if (sidx < nb) pr.dgap[(long long)(bs + sidx) * C + lane] = g;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(&rowsB[warp][k][4 * q]);
            acc[4 * q] = fmaf(v.x, dh[k], acc[4 * q]);
            acc[4 * q + 1] = fmaf(v.y, dh[k], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v.z, dh[k], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v.w, dh[k], acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) redW[warp][c][lane] = acc[c];
    red[0][warp][lane] = sd0;
    __syncthreads();
    if (nb > 0) {
        if (warp == 0) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) a += red[0][w][lane];
            atomicAdd(pr.db0 + lane, a);
        }
#pragma unroll
        for (int o = t; o < C * C; o += kC32Threads) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += (&redW[w][0][0])[o];
            atomicAdd(pr.dw0 + o, v);
        }
    }
    cluster_wait();                                 // peers may still be reading this CTA's `part`
}

constexpr size_t kSeSmemMax = 227 * 1024;

}  // namespace mvae

using namespace mvae;

extern "C" long long mvae_se_gate_ws_floats(int B, int C) { return 6LL * B * C + 2LL * C + 2LL * kCS * 2 * C; }

// samples per warp of the C = 32 kernels (0: use the generic kernel)
static int c32_spw(int B, int C) {
    if (C != 32 || env_int("MVAE_SE_GENERIC", 0)) return 0;
    const int per = (B + kCS - 1) / kCS;
    const int spw = (per + kC32Warps - 1) / kC32Warps;
    return spw <= 1 ? 1 : spw <= 2 ? 2 : spw <= 4 ? 4 : 0;
}

static int se_fwd_launch(FwdBatch& bt, cudaStream_t s) {
    if (const int spw = c32_spw(bt.B, bt.C)) {
        const dim3 grid(kCS * bt.n), block(kC32Threads);
        if (spw == 1) MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_fwd_c32_kernel<1>, grid, block, 0, s, bt));
        else if (spw == 2) MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_fwd_c32_kernel<2>, grid, block, 0, s, bt));
        else MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_fwd_c32_kernel<4>, grid, block, 0, s, bt));
        MVAE_LAUNCH_CHECK();
        return MVAE_OK;
    }
    const int per = (bt.B + kCS - 1) / kCS;
    const size_t smem = ((size_t)3 * per * bt.C + 2 * (size_t)bt.C * bt.C + 8 * bt.C) * sizeof(float);
    MVAE_REQUIRE(smem <= kSeSmemMax, "se_gate_fwd: B*C = %d too large for the shared-memory gate kernel", bt.B * bt.C);
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(se_gate_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeSmemMax));
    }
    MVAE_CUDA(launch_pdl(se_gate_fwd_kernel, dim3(kCS * bt.n), dim3(kSeThreads), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

static int se_bwd_launch(BwdBatch& bt, cudaStream_t s) {
    if (const int spw = c32_spw(bt.B, bt.C)) {
        const dim3 grid(kCS * bt.n), block(kC32Threads);
        if (spw == 1) MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_bwd_c32_kernel<1>, grid, block, 0, s, bt));
        else if (spw == 2) MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_bwd_c32_kernel<2>, grid, block, 0, s, bt));
        else MVAE_CUDA(launch_pdl_ex(pdl_enabled() || pdl_chain_enabled(), se_gate_bwd_c32_kernel<4>, grid, block, 0, s, bt));
        MVAE_LAUNCH_CHECK();
        return MVAE_OK;
    }
    const int per = (bt.B + kCS - 1) / kCS;
    const size_t smem = ((size_t)4 * per * bt.C + 2 * (size_t)bt.C * (bt.C + 1) + 8 * bt.C) * sizeof(float);
    MVAE_REQUIRE(smem <= kSeSmemMax, "se_gate_bwd: B*C = %d too large for the shared-memory gate kernel", bt.B * bt.C);
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        MVAE_CUDA(cudaFuncSetAttribute(se_gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeSmemMax));
    }
    MVAE_CUDA(launch_pdl(se_gate_bwd_kernel, dim3(kCS * bt.n), dim3(kSeThreads), smem, s, bt));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_se_gate_fwd_batched(int n, const float* const* gap_sum, const float* const* w0, const float* const* b0,
                                        const float* const* gamma, const float* const* beta, const float* const* w1,
                                        const float* const* b1, float* const* moving_mean, float* const* moving_var,
                                        float* const* gate, float* const* ws, int B, int C, const int* HW, float eps,
                                        float momentum, int training, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && gap_sum && w0 && b0 && gamma && beta && w1 && b1 && moving_mean && moving_var && gate && ws && HW,
                 "se_gate_fwd_batched: null pointer");
    MVAE_REQUIRE(B > 0 && C > 0, "se_gate_fwd_batched: bad sizes");
    for (int l0 = 0; l0 < n; l0 += kMaxBatch) {
        FwdBatch bt;
        bt.n = n - l0 < kMaxBatch ? n - l0 : kMaxBatch;
        bt.B = B; bt.C = C; bt.training = training; bt.eps = eps; bt.momentum = momentum; bt.trace = g_trace;
        for (int l = 0; l < bt.n; ++l) {
            const int k = l0 + l;
            MVAE_REQUIRE(gap_sum[k] && w0[k] && b0[k] && gamma[k] && beta[k] && w1[k] && b1[k] && moving_mean[k] &&
                         moving_var[k] && gate[k] && ws[k] && HW[k] > 0, "se_gate_fwd_batched: bad member %d", k);
            bt.p[l] = FwdP{gap_sum[k], w0[k], b0[k], gamma[k], beta[k], w1[k], b1[k], moving_mean[k], moving_var[k], gate[k],
                           ws[k], 1.f / (float)HW[k]};
        }
        if (int e = se_fwd_launch(bt, as_stream(stream))) return e;
    }
    return MVAE_OK;
}

extern "C" int mvae_se_gate_bwd_batched(int n, const float* const* dg, const float* const* w0, const float* const* gamma,
                                        const float* const* beta, const float* const* w1, float* const* ws,
                                        float* const* dgap, float* const* dw0, float* const* db0, float* const* dgamma,
                                        float* const* dbeta, float* const* dw1, float* const* db1, int B, int C,
                                        const int* HW, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && dg && w0 && gamma && beta && w1 && ws && dgap && dw0 && db0 && dgamma && dbeta && dw1 && db1 && HW,
                 "se_gate_bwd_batched: null pointer");
    MVAE_REQUIRE(B > 0 && C > 0, "se_gate_bwd_batched: bad sizes");
    for (int l0 = 0; l0 < n; l0 += kMaxBatch) {
        BwdBatch bt;
        bt.n = n - l0 < kMaxBatch ? n - l0 : kMaxBatch;
        bt.B = B; bt.C = C; bt.trace = g_trace;
        for (int l = 0; l < bt.n; ++l) {
            const int k = l0 + l;
            MVAE_REQUIRE(dg[k] && w0[k] && gamma[k] && beta[k] && w1[k] && ws[k] && dgap[k] && dw0[k] && db0[k] && dgamma[k] &&
                         dbeta[k] && dw1[k] && db1[k] && HW[k] > 0, "se_gate_bwd_batched: bad member %d", k);
            bt.p[l] = BwdP{dg[k], w0[k], gamma[k], beta[k], w1[k], ws[k], dgap[k], dw0[k], db0[k], dgamma[k], dbeta[k], dw1[k],
                           db1[k], 1.f / (float)HW[k]};
        }
        if (int e = se_bwd_launch(bt, as_stream(stream))) return e;
    }
    return MVAE_OK;
}

extern "C" int mvae_se_gate_fwd(const float* gap_sum, const float* w0, const float* b0, const float* gamma,
                                const float* beta, const float* w1, const float* b1, float* moving_mean,
                                float* moving_var, float* gate, float* ws, int B, int C, int HW, float eps,
                                float momentum, int training, mvae_stream_t stream) {
    MVAE_REQUIRE(HW > 0, "se_gate_fwd: bad sizes");
    return mvae_se_gate_fwd_batched(1, &gap_sum, &w0, &b0, &gamma, &beta, &w1, &b1, &moving_mean, &moving_var, &gate, &ws, B, C,
                                    &HW, eps, momentum, training, stream);
}

extern "C" int mvae_se_gate_bwd(const float* dg, const float* w0, const float* gamma, const float* beta, const float* w1,
                                float* ws, float* dgap, float* dw0, float* db0, float* dgamma, float* dbeta, float* dw1,
                                float* db1, int B, int C, int HW, mvae_stream_t stream) {
    MVAE_REQUIRE(HW > 0, "se_gate_bwd: bad sizes");
    return mvae_se_gate_bwd_batched(1, &dg, &w0, &gamma, &beta, &w1, &ws, &dgap, &dw0, &db0, &dgamma, &dbeta, &dw1, &db1, B, C, &HW,
                                    stream);
}
