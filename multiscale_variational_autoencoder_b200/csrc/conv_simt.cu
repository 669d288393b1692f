// fp32 (CUDA-core) implicit-GEMM convolution: forward, dgrad (== Conv2DTranspose forward) and wgrad.
// This is the MVAE_PREC_FP32 path: bit-faithful fp32 accumulation used for tight parity runs and for the shapes
// the tensor-core path does not take (Cin = 3 + CoordConv channels).  Reference call sites: see mvae_b200.h.
#include "common.cuh"

namespace mvae {

struct ConvGeom {
    int B, H, W, Cin;       // forward input
    int Ho, Wo, Cout;       // forward output
    int kh, kw, sh, sw, pt, pl;
    int coord;              // number of generated CoordConv channels (0, 2, 3)
    int CinT;               // Cin + coord
};

constexpr int BM = 128, BN = 32, BK = 32, LDS_PAD = 36;

__device__ __forceinline__ float coord_value(int which, int iy, int ix, int H, int W) {
    const float xx = (float)iy / (float)(H - 1) * 2.f - 1.f;
    const float yy = (float)ix / (float)(W - 1) * 2.f - 1.f;
    if (which == 0) return xx;
    if (which == 1) return yy;
    return sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f));
}

// ---------------------------------------------------------------------------------------------------------
// MODE 0 (fwd)  : out[m=(b,oy,ox)][n=co] = sum_k A[m][k=(tap,ci)] * w[k][n],   A gathers x (x gate, + coord)
// MODE 1 (dgrad): out[m=(b,iy,ix)][n=ci] = sum_k A[m][k=(tap,co)] * w[tap][n][co], A gathers dy
// epilogue: v = act(acc + bias) + residual; v *= gact'(act_out)   (act: forward activation; gact: whose derivative)
// VEC: the reduction channel count (CinT for fwd, Cout for dgrad) and N are multiples of 4 and coord == 0.
// ---------------------------------------------------------------------------------------------------------
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) igemm_kernel(ConvGeom g, const float* __restrict__ src,
                                                    const float* __restrict__ wt, const float* __restrict__ bias,
                                                    const float* __restrict__ gate, const float* __restrict__ residual,
                                                    const float* __restrict__ act_out, int act, int gact,
                                                    float* __restrict__ out, int M, int N, int K, int klen) {
    pdl_sync();
    __shared__ __align__(16) float As[BM][LDS_PAD];
    __shared__ __align__(16) float Bs[BK][LDS_PAD];

    const int tid = threadIdx.x;
    // (m tile, n tile) flattened into grid.x: a Dense layer with millions of outputs has more n tiles than grid.y allows
    const int mtl = (M + BM - 1) / BM;
    const int m0 = (int)(blockIdx.x % mtl) * BM, n0 = (int)(blockIdx.x / mtl) * BN;
    const int kbeg = blockIdx.z * klen;
    const int kend = min(K, kbeg + klen);
    const int CR = (MODE == 0) ? g.CinT : g.Cout;   // channels per tap along the reduction

    // ---- per-thread A rows: r = tid/8 + 32*j, quad q = tid%8 -------------------------------------------
    const int aq = tid & 7;
    int rb[4], ry[4], rx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = m0 + (tid >> 3) + 32 * j;
        if (m < M) {
            if (MODE == 0) {
                const int ox = m % g.Wo, t = m / g.Wo;
                const int oy = t % g.Ho;
                rb[j] = t / g.Ho; ry[j] = oy * g.sh - g.pt; rx[j] = ox * g.sw - g.pl;
            } else {
                const int ix = m % g.W, t = m / g.W;
                const int iy = t % g.H;
                rb[j] = t / g.H; ry[j] = iy + g.pt; rx[j] = ix + g.pl;
            }
        } else {
            rb[j] = -1; ry[j] = 0; rx[j] = 0;
        }
    }

    auto load_a = [&](int j, int k) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rb[j] < 0 || k >= kend) return v;
        if (VEC) {
            const int tap = k / CR, ch = k - tap * CR;
            const int ky = tap / g.kw, kx = tap - ky * g.kw;
            if (MODE == 0) {
                const int iy = ry[j] + ky, ix = rx[j] + kx;
                if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                    v = __ldg(reinterpret_cast<const float4*>(src + (((long long)rb[j] * g.H + iy) * g.W + ix) * g.Cin + ch));
                    if (gate) {
                        const float4 gt = __ldg(reinterpret_cast<const float4*>(gate + (long long)rb[j] * g.Cin + ch));
                        v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
                    }
                }
            } else {
                const int ty = ry[j] - ky, tx = rx[j] - kx;
                if (ty >= 0 && tx >= 0 && (ty % g.sh) == 0 && (tx % g.sw) == 0) {
                    const int oy = ty / g.sh, ox = tx / g.sw;
                    if (oy < g.Ho && ox < g.Wo)
                        v = __ldg(reinterpret_cast<const float4*>(src + (((long long)rb[j] * g.Ho + oy) * g.Wo + ox) * g.Cout + ch));
                }
            }
        } else {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = k + i;
                if (kk >= kend) break;
                const int tap = kk / CR, ch = kk - tap * CR;
                const int ky = tap / g.kw, kx = tap - ky * g.kw;
                if (MODE == 0) {
                    const int iy = ry[j] + ky, ix = rx[j] + kx;
                    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                        if (ch < g.Cin) {
                            float t = __ldg(src + (((long long)rb[j] * g.H + iy) * g.W + ix) * g.Cin + ch);
                            if (gate) t *= __ldg(gate + (long long)rb[j] * g.Cin + ch);
                            e[i] = t;
                        } else {
                            e[i] = coord_value(ch - g.Cin, iy, ix, g.H, g.W);
                        }
                    }
                } else {
                    const int ty = ry[j] - ky, tx = rx[j] - kx;
                    if (ty >= 0 && tx >= 0 && (ty % g.sh) == 0 && (tx % g.sw) == 0) {
                        const int oy = ty / g.sh, ox = tx / g.sw;
                        if (oy < g.Ho && ox < g.Wo)
                            e[i] = __ldg(src + (((long long)rb[j] * g.Ho + oy) * g.Wo + ox) * g.Cout + ch);
                    }
                }
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        return v;
    };

    // ---- per-thread B quad: kk = tid/8, nq = tid%8 -----------------------------------------------------
    const int bk = tid >> 3, bn = n0 + 4 * (tid & 7);
    auto load_b = [&](int k0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + bk;
        if (!(MODE == 1 && VEC) && k >= kend) return v;
        if (MODE == 0) {
            if (VEC) {
                if (bn < N) v = __ldg(reinterpret_cast<const float4*>(wt + (long long)k * N + bn));
            } else {
                float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; ++i) if (bn + i < N) e[i] = __ldg(wt + (long long)k * N + bn + i);
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
        } else if (VEC) {
            // dgrad, Cout % 4 == 0: the Keras kernel (tap, n, co) is contiguous along co = the reduction index, so a thread
            // fetches 4 consecutive k of ONE output column n (16-byte load, a row of 32 k per 8 threads) and the store below
            // transposes into Bs[k][n]; reading 4 columns at one k touched 16 bytes of 32 different 128-byte lines
            const int nn = n0 + (tid >> 3), kq = k0 + 4 * (tid & 7);
            if (nn < N && kq < kend) {
                const int tap = kq / g.Cout, co = kq - tap * g.Cout;
                v = __ldg(reinterpret_cast<const float4*>(wt + ((long long)tap * g.CinT + nn) * g.Cout + co));
            }
        } else {
            const int tap = k / g.Cout, co = k - tap * g.Cout;
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (bn + i < N) e[i] = __ldg(wt + ((long long)tap * g.CinT + bn + i) * g.Cout + co);
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        return v;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int tm = tid >> 3, tn = tid & 7;
    float4 pa[4], pb;
#pragma unroll
    for (int j = 0; j < 4; ++j) pa[j] = load_a(j, kbeg + 4 * aq);
    pb = load_b(kbeg);

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(&As[(tid >> 3) + 32 * j][4 * aq]) = pa[j];
        if (MODE == 1 && VEC) {
            const int nn = tid >> 3, kq = 4 * (tid & 7);
            Bs[kq][nn] = pb.x; Bs[kq + 1][nn] = pb.y; Bs[kq + 2][nn] = pb.z; Bs[kq + 3][nn] = pb.w;
        } else {
            *reinterpret_cast<float4*>(&Bs[bk][4 * (tid & 7)]) = pb;
        }
        __syncthreads();
        if (k0 + BK < kend) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[j] = load_a(j, k0 + BK + 4 * aq);
            pb = load_b(k0 + BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&As[tm + 32 * i][kk]);
#pragma unroll
            for (int e = 0; e < 4; ++e) b[e] = *reinterpret_cast<const float4*>(&Bs[kk + e][4 * tn]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(a[i].x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a[i].x, b[0].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a[i].x, b[0].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a[i].y, b[1].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a[i].y, b[1].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a[i].z, b[2].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a[i].z, b[2].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a[i].w, b[3].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a[i].w, b[3].w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ---------------------------------------------------------------------------------------
    const bool split = gridDim.z > 1;
    const bool lead = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm + 32 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tn + j;
            if (n >= N) continue;
            const long long o = (long long)m * N + n;
            float v = acc[i][j];
            if (!split) {
                if (bias) v += __ldg(bias + n);
                v = act_apply(v, act);
                if (residual) v += __ldg(residual + o);
                if (act_out) v *= act_grad_from_out(__ldg(act_out + o), gact);
                out[o] = v;
            } else {
                if (lead) {
                    if (bias) v += __ldg(bias + n);
                    if (residual) v += __ldg(residual + o);
                }
                atomicAdd(out + o, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dw[k'=(tap,ci)][n=co] += sum_p A'[p][k'] * dy[p][n];  dbias[n] += sum_p dy[p][n]
// grid = (K' tiles of 128, N tiles of 32, pixel splits)
// ---------------------------------------------------------------------------------------------------------
constexpr int WP = 32;   // pixels per chunk

template <bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(ConvGeom g, const float* __restrict__ x, const float* __restrict__ gate,
                                                    const float* __restrict__ dy, float* __restrict__ dw,
                                                    float* __restrict__ dbias, int P, int N, int KP, int plen) {
    pdl_sync();
    __shared__ __align__(16) float As[WP][BM];
    __shared__ __align__(16) float Bs[WP][LDS_PAD];

    const int tid = threadIdx.x;
    const int ktl = (KP + BM - 1) / BM;                         // (k' tile, n tile) flattened into grid.x
    const int kt0 = (int)(blockIdx.x % ktl) * BM, n0 = (int)(blockIdx.x / ktl) * BN;
    const int pbeg = blockIdx.z * plen;
    const int pend = min(P, pbeg + plen);
    if (pbeg >= pend) return;

    // A' loader: quad q = tid%32 (k' = kt0 + 4q, fixed), pixel pp = tid/32 + 8j
    const int akq = kt0 + 4 * (tid & 31);
    int a_ky[4], a_kx[4], a_ch[4];
    bool a_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = akq + i;
        a_ok[i] = k < KP;
        const int tap = a_ok[i] ? k / g.CinT : 0;
        a_ch[i] = a_ok[i] ? k - tap * g.CinT : 0;
        a_ky[i] = tap / g.kw; a_kx[i] = tap - a_ky[i] * g.kw;
    }

    auto load_a = [&](int j, int p0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p = p0 + (tid >> 5) + 8 * j;
        if (p >= pend || !a_ok[0]) return v;
        const int ox = p % g.Wo, t = p / g.Wo;
        const int oy = t % g.Ho, b = t / g.Ho;
        if (VEC) {
            const int iy = oy * g.sh - g.pt + a_ky[0], ix = ox * g.sw - g.pl + a_kx[0];
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * g.H + iy) * g.W + ix) * g.Cin + a_ch[0]));
                if (gate) {
                    const float4 gt = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * g.Cin + a_ch[0]));
                    v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
                }
            }
        } else {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!a_ok[i]) continue;
                const int iy = oy * g.sh - g.pt + a_ky[i], ix = ox * g.sw - g.pl + a_kx[i];
                if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                    if (a_ch[i] < g.Cin) {
                        float tv = __ldg(x + (((long long)b * g.H + iy) * g.W + ix) * g.Cin + a_ch[i]);
                        if (gate) tv *= __ldg(gate + (long long)b * g.Cin + a_ch[i]);
                        e[i] = tv;
                    } else {
                        e[i] = coord_value(a_ch[i] - g.Cin, iy, ix, g.H, g.W);
                    }
                }
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        return v;
    };

    const int bp = tid >> 3, bn = n0 + 4 * (tid & 7);
    auto load_b = [&](int p0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p = p0 + bp;
        if (p >= pend) return v;
        if (VEC) {
            if (bn < N) v = __ldg(reinterpret_cast<const float4*>(dy + (long long)p * N + bn));
        } else {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i) if (bn + i < N) e[i] = __ldg(dy + (long long)p * N + bn + i);
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        return v;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool do_bias = dbias != nullptr && (blockIdx.x % ktl) == 0;

    const int tm = tid >> 3, tn = tid & 7;
    float4 pa[4], pb;
#pragma unroll
    for (int j = 0; j < 4; ++j) pa[j] = load_a(j, pbeg);
    pb = load_b(pbeg);

    for (int p0 = pbeg; p0 < pend; p0 += WP) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(&As[(tid >> 5) + 8 * j][4 * (tid & 31)]) = pa[j];
        *reinterpret_cast<float4*>(&Bs[bp][4 * (tid & 7)]) = pb;
        __syncthreads();
        if (p0 + WP < pend) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[j] = load_a(j, p0 + WP);
            pb = load_b(p0 + WP);
        }
#pragma unroll 8
        for (int pp = 0; pp < WP; ++pp) {
            const float4 a = *reinterpret_cast<const float4*>(&As[pp][4 * tm]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[pp][4 * tn]);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
            acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
            acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
            acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
            acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
        if (do_bias && tid < BN) {
#pragma unroll 8
            for (int pp = 0; pp < WP; ++pp) bsum += Bs[pp][tid];
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = kt0 + 4 * tm + i;
        if (k >= KP) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tn + j;
            if (n < N) atomicAdd(dw + (long long)k * N + n, acc[i][j]);
        }
    }
    if (do_bias && tid < BN && n0 + tid < N) atomicAdd(dbias + n0 + tid, bsum);
}

static int make_geom(const mvae_conv_desc* d, ConvGeom& g) {
    MVAE_REQUIRE(d, "conv: null descriptor");
    MVAE_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: bad sizes");
    MVAE_REQUIRE(d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0, "conv: bad kernel/stride");
    MVAE_REQUIRE(d->coord_mode == 0 || d->coord_mode == 2 || d->coord_mode == 3, "conv: coord_mode must be 0, 2 or 3");
    MVAE_REQUIRE(d->coord_mode == 0 || (d->H > 1 && d->W > 1), "conv: CoordConv needs H,W > 1 (coord.py:118,122)");
    g.B = d->B; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Cout = d->Cout;
    g.kh = d->kh; g.kw = d->kw; g.sh = d->sh; g.sw = d->sw;
    same_pad(d->H, d->kh, d->sh, &g.Ho, &g.pt);
    same_pad(d->W, d->kw, d->sw, &g.Wo, &g.pl);
    g.coord = d->coord_mode;
    g.CinT = d->Cin + d->coord_mode;
    MVAE_REQUIRE((long long)g.B * g.H * g.W < (1LL << 31) && (long long)g.B * g.Ho * g.Wo < (1LL << 31), "conv: too many pixels");
    return MVAE_OK;
}

static inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pick_ksplit(int mtiles, int ntiles, int K, int* klen) {
    int splits = 1;
    const int ctas = mtiles * ntiles;
    if (ctas < kNumSMs && K >= 4 * BK) {
        splits = (2 * kNumSMs + ctas - 1) / ctas;
        const int maxs = K / (2 * BK);
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    int len = (K + splits - 1) / splits;
    len = ((len + BK - 1) / BK) * BK;
    *klen = len;
    return (K + len - 1) / len;
}

int conv_fwd_fp32(const ConvGeom& g, const float* x, const float* w, const float* bias, const float* gate,
                  const float* residual, int act, float* y, cudaStream_t s) {
    const int M = g.B * g.Ho * g.Wo, N = g.Cout, K = g.kh * g.kw * g.CinT;
    const int mt = ceil_div(M, BM), nt = ceil_div(N, BN);
    int klen;
    int splits = pick_ksplit(mt, nt, K, &klen);
    if (act != MVAE_ACT_NONE) { splits = 1; klen = ((K + BK - 1) / BK) * BK; }
    if (splits > 1) MVAE_CUDA(cudaMemsetAsync(y, 0, (size_t)M * N * sizeof(float), s));
    const bool vec = g.coord == 0 && (g.Cin % 4) == 0 && (N % 4) == 0 && aligned16(x) && aligned16(w) && aligned16(gate);
    dim3 grid((unsigned)mt * (unsigned)nt, 1, splits);
    if (vec) MVAE_CUDA(launch_pdl(igemm_kernel<0, true>, dim3(grid), dim3(256), 0, s, g, x, w, bias, gate, residual, nullptr, act, 0, y, M, N, K, klen));
    else     MVAE_CUDA(launch_pdl(igemm_kernel<0, false>, dim3(grid), dim3(256), 0, s, g, x, w, bias, gate, residual, nullptr, act, 0, y, M, N, K, klen));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

int conv_dgrad_fp32(const ConvGeom& g, const float* dy, const float* w, const float* bias, const float* residual,
                    const float* act_out, int act, float* dx, cudaStream_t s) {
    const int M = g.B * g.H * g.W, N = g.Cin, K = g.kh * g.kw * g.Cout;
    const int mt = ceil_div(M, BM), nt = ceil_div(N, BN);
    int klen;
    int splits = pick_ksplit(mt, nt, K, &klen);
    if (act_out != nullptr) { splits = 1; klen = ((K + BK - 1) / BK) * BK; }
    if (splits > 1) MVAE_CUDA(cudaMemsetAsync(dx, 0, (size_t)M * N * sizeof(float), s));
    const bool vec = (g.Cout % 4) == 0 && aligned16(dy);
    dim3 grid((unsigned)mt * (unsigned)nt, 1, splits);
    if (vec) MVAE_CUDA(launch_pdl(igemm_kernel<1, true>, dim3(grid), dim3(256), 0, s, g, dy, w, bias, nullptr, residual, act_out, 0, act, dx, M, N, K, klen));
    else     MVAE_CUDA(launch_pdl(igemm_kernel<1, false>, dim3(grid), dim3(256), 0, s, g, dy, w, bias, nullptr, residual, act_out, 0, act, dx, M, N, K, klen));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

int conv_wgrad_fp32(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                    cudaStream_t s) {
    const int P = g.B * g.Ho * g.Wo, N = g.Cout, KP = g.kh * g.kw * g.CinT;
    const int kt = ceil_div(KP, BM), nt = ceil_div(N, BN);
    int splits = ceil_div(2 * kNumSMs, kt * nt);
    const int maxs = ceil_div(P, 2 * WP);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    int plen = ceil_div(P, splits);
    plen = ((plen + WP - 1) / WP) * WP;
    splits = ceil_div(P, plen);
    const bool vec = g.coord == 0 && (g.Cin % 4) == 0 && (N % 4) == 0 && aligned16(x) && aligned16(dy) && aligned16(gate);
    dim3 grid((unsigned)kt * (unsigned)nt, 1, splits);
    if (vec) MVAE_CUDA(launch_pdl(wgrad_kernel<true>, dim3(grid), dim3(256), 0, s, g, x, gate, dy, dw, dbias, P, N, KP, plen));
    else     MVAE_CUDA(launch_pdl(wgrad_kernel<false>, dim3(grid), dim3(256), 0, s, g, x, gate, dy, dw, dbias, P, N, KP, plen));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

// tensor-core path (conv_tc.cu); returns MVAE_ERR_UNSUPPORTED when the shape is not covered
int conv_fwd_tc(const ConvGeom& g, const float* x, const float* w, const float* bias, const float* gate,
                const float* residual, int act, float* y, cudaStream_t s);
int conv_dgrad_tc(const ConvGeom& g, const float* dy, const float* w, const float* bias, const float* residual,
                  const float* act_out, int act, float* dx, cudaStream_t s);
size_t dense_tc_workspace_bytes(int M, int K, int N);
int dense_fwd_tc(int M, int K, int N, const float* x, const float* w, const float* bias, int act, float* y, float* ws,
                 size_t ws_bytes, cudaStream_t s);
int dense_dgrad_tc(int M, int K, int N, const float* dy, const float* w, const float* act_out, int gact, float* dx, float* ws,
                   size_t ws_bytes, cudaStream_t s);
int conv_fwd_small_cin(const ConvGeom& g, const float* x, const float* w, const float* bias, const float* gate,
                       const float* residual, int act, float* y, cudaStream_t s);
int conv_wgrad_small_cin(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                         cudaStream_t s);
int conv_fwd_tc_batched(int n, const ConvGeom* g, const float* const* x, const float* const* w, const float* const* bias,
                        const float* const* gate, const float* const* residual, int act, float* const* y, cudaStream_t s);
int conv_dgrad_tc_batched(int n, const ConvGeom* g, const float* const* dy, const float* const* w, const float* const* bias,
                          const float* const* residual, const float* const* act_out, int act, float* const* dx,
                          cudaStream_t s);
int conv_wgrad_tc_batched(int n, const ConvGeom* g, const float* const* x, const float* const* gate, const float* const* dy,
                          float* const* dw, float* const* dbias, cudaStream_t s);
int conv_wgrad_tc(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                  cudaStream_t s);

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_conv2d_fwd(const mvae_conv_desc* d, const float* x, const float* w, const float* bias,
                               const float* gate, const float* residual, int act, float* y, mvae_stream_t stream) {
    ConvGeom g;
    if (int e = make_geom(d, g)) return e;
    MVAE_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
    MVAE_REQUIRE(act >= MVAE_ACT_NONE && act <= MVAE_ACT_ELU, "conv2d_fwd: bad activation");
    MVAE_REQUIRE(!(gate && g.coord), "conv2d_fwd: gate with CoordConv channels is not supported");
    if (d->precision == MVAE_PREC_TF32) {
        const int r = conv_fwd_tc(g, x, w, bias, gate, residual, act, y, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    {
        // few-channel first layer (conv_base): direct kernel, fp32 in both precision modes
        const int r = conv_fwd_small_cin(g, x, w, bias, gate, residual, act, y, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    return conv_fwd_fp32(g, x, w, bias, gate, residual, act, y, as_stream(stream));
}

extern "C" int mvae_conv2d_dgrad(const mvae_conv_desc* d, const float* dy, const float* w, const float* bias,
                                 const float* residual, const float* act_out, int act, float* dx,
                                 mvae_stream_t stream) {
    ConvGeom g;
    if (int e = make_geom(d, g)) return e;
    MVAE_REQUIRE(dy && w && dx, "conv2d_dgrad: null pointer");
    MVAE_REQUIRE(g.coord == 0, "conv2d_dgrad: not defined for CoordConv inputs (the input is data)");
    MVAE_REQUIRE(act >= MVAE_ACT_NONE && act <= MVAE_ACT_ELU, "conv2d_dgrad: bad activation");
    if (d->precision == MVAE_PREC_TF32) {
        const int r = conv_dgrad_tc(g, dy, w, bias, residual, act_out, act, dx, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    return conv_dgrad_fp32(g, dy, w, bias, residual, act_out, act, dx, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------------------------------
// Dense layers: the tensor-core skinny GEMM of dense_tc.cu when the shape fits, else the convolution path with H = W = 1
// ---------------------------------------------------------------------------------------------------------------------
extern "C" size_t mvae_dense_workspace_bytes(int M, int K, int N) { return dense_tc_workspace_bytes(M, K, N); }

extern "C" int mvae_dense_fwd(int M, int K, int N, const float* x, const float* w, const float* bias, int act, float* y,
                              float* ws, size_t ws_bytes, int precision, mvae_stream_t stream) {
    MVAE_REQUIRE(M > 0 && K > 0 && N > 0 && x && w && y, "dense_fwd: bad arguments");
    MVAE_REQUIRE(act >= MVAE_ACT_NONE && act <= MVAE_ACT_ELU, "dense_fwd: bad activation");
    if (precision == MVAE_PREC_TF32) {
        const int r = dense_fwd_tc(M, K, N, x, w, bias, act, y, ws, ws_bytes, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    const mvae_conv_desc d = {M, 1, 1, K, 1, 1, 1, 1, N, 0, precision};
    return mvae_conv2d_fwd(&d, x, w, bias, nullptr, nullptr, act, y, stream);
}

extern "C" int mvae_dense_dgrad(int M, int K, int N, const float* dy, const float* w, const float* act_out, int act,
                                float* dx, float* ws, size_t ws_bytes, int precision, mvae_stream_t stream) {
    MVAE_REQUIRE(M > 0 && K > 0 && N > 0 && dy && w && dx, "dense_dgrad: bad arguments");
    MVAE_REQUIRE(act >= MVAE_ACT_NONE && act <= MVAE_ACT_ELU, "dense_dgrad: bad activation");
    if (precision == MVAE_PREC_TF32) {
        const int r = dense_dgrad_tc(M, K, N, dy, w, act_out, act, dx, ws, ws_bytes, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    const mvae_conv_desc d = {M, 1, 1, K, 1, 1, 1, 1, N, 0, precision};
    return mvae_conv2d_dgrad(&d, dy, w, nullptr, nullptr, act_out, act, dx, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient of a Dense layer with few inputs and many outputs (the decoder heads: z -> h*w*c, multiscale_vae.py:402-406;
// 128 x 8192 at cfg2, 32 x 2097152 at cfg4): dW[k][n] = sum_b x[b][k] dy[b][n] is an outer-product sum over a small batch whose
// cost is writing dW once and reading dy once.  The implicit-GEMM kernel above spends it on index arithmetic (41 us at cfg2
// level 0, 1.4 ms at cfg4).  Here a thread owns 4 output columns and a chunk of 32 input rows: 128 accumulators in registers,
// dy streamed with 16-byte loads, the x chunk broadcast from shared memory; one writer per element, so the result does not
// depend on any arrival order.
// ---------------------------------------------------------------------------------------------------------------------
namespace dwo {
constexpr int kThreads = 128, kKC = 32, kBT = 64;

__global__ void __launch_bounds__(kThreads) dense_wgrad_outer_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     float* __restrict__ dw, float* __restrict__ dbias, int M,
                                                                     int K, int N) {
    pdl_sync();
    __shared__ __align__(16) float xs[kBT][kKC];
    const int n0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
    const int k0 = blockIdx.y * kKC;
    const bool live = n0 < N;
    float4 acc[kKC];
#pragma unroll
    for (int k = 0; k < kKC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b0 = 0; b0 < M; b0 += kBT) {
        __syncthreads();
        for (int i = threadIdx.x; i < kBT * kKC; i += kThreads) {
            const int bb = i / kKC, kk = i - bb * kKC;
            xs[bb][kk] = (b0 + bb < M && k0 + kk < K) ? __ldg(x + (long long)(b0 + bb) * K + k0 + kk) : 0.f;
        }
        __syncthreads();
        if (live) {
            const int nb = min(kBT, M - b0);
#pragma unroll 4
            for (int bb = 0; bb < nb; ++bb) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (long long)(b0 + bb) * N + n0));
                bs.x += d.x; bs.y += d.y; bs.z += d.z; bs.w += d.w;
#pragma unroll
                for (int k4 = 0; k4 < kKC / 4; ++k4) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[bb][4 * k4]);
                    const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float4& a = acc[4 * k4 + e];
                        a.x = fmaf(xk[e], d.x, a.x); a.y = fmaf(xk[e], d.y, a.y);
                        a.z = fmaf(xk[e], d.z, a.z); a.w = fmaf(xk[e], d.w, a.w);
                    }
                }
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int k = 0; k < kKC; ++k)
        if (k0 + k < K)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + (long long)(k0 + k) * N + n0), "f"(acc[k].x),
                         "f"(acc[k].y), "f"(acc[k].z), "f"(acc[k].w) : "memory");
    if (dbias && blockIdx.y == 0)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbias + n0), "f"(bs.x), "f"(bs.y), "f"(bs.z), "f"(bs.w)
                     : "memory");
}
}  // namespace dwo

static int dense_wgrad_outer(const ConvGeom& g, const float* x, const float* gate, const float* dy, float* dw, float* dbias,
                             cudaStream_t s) {
    const int N = g.Cout, K = g.CinT, M = g.B;
    if (g.H != 1 || g.W != 1 || g.kh != 1 || g.kw != 1 || g.coord != 0 || gate != nullptr) return MVAE_ERR_UNSUPPORTED;
    if (N < 1024 || (N & 3) || K > 1024 || N < 8 * K) return MVAE_ERR_UNSUPPORTED;      // wide outputs, few inputs
    if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(dw) & 15) ||
        (dbias && (reinterpret_cast<uintptr_t>(dbias) & 15)))
        return MVAE_ERR_UNSUPPORTED;
    dim3 grid(ceil_div(N / 4, dwo::kThreads), ceil_div(K, dwo::kKC));
    MVAE_CUDA(launch_pdl(dwo::dense_wgrad_outer_kernel, grid, dim3(dwo::kThreads), 0, s, x, dy, dw, dbias, M, K, N));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
}

extern "C" int mvae_conv2d_wgrad(const mvae_conv_desc* d, const float* x, const float* gate, const float* dy, float* dw,
                                 float* dbias, mvae_stream_t stream) {
    ConvGeom g;
    if (int e = make_geom(d, g)) return e;
    MVAE_REQUIRE(x && dy && dw, "conv2d_wgrad: null pointer");
    MVAE_REQUIRE(!(gate && g.coord), "conv2d_wgrad: gate with CoordConv channels is not supported");
    if (d->precision == MVAE_PREC_TF32) {
        const int r = conv_wgrad_tc(g, x, gate, dy, dw, dbias, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    {
        const int r = dense_wgrad_outer(g, x, gate, dy, dw, dbias, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    {
        const int r = conv_wgrad_small_cin(g, x, gate, dy, dw, dbias, as_stream(stream));
        if (r != MVAE_ERR_UNSUPPORTED) return r;
    }
    return conv_wgrad_fp32(g, x, gate, dy, dw, dbias, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------------------------------
// Batched entries: n independent problems of the same layer (the pyramid levels).  Semantics == n single calls; when all
// members fit the TMA tensor-core kernels they share ONE launch, otherwise they are issued one by one.
// ---------------------------------------------------------------------------------------------------------------------
static inline const float* opt(const float* const* a, int l) { return a ? a[l] : nullptr; }

extern "C" int mvae_conv2d_fwd_batched(int n, const mvae_conv_desc* d, const float* const* x, const float* const* w,
                                       const float* const* bias, const float* const* gate, const float* const* residual,
                                       int act, float* const* y, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && d && x && w && y, "conv2d_fwd_batched: bad arguments");
    if (n <= 8 && d[0].precision == MVAE_PREC_TF32) {
        ConvGeom g[8];
        bool ok = true;
        for (int l = 0; l < n && ok; ++l) ok = make_geom(d + l, g[l]) == MVAE_OK && d[l].precision == MVAE_PREC_TF32 && x[l] && w[l] && y[l];
        if (ok) {
            const int r = conv_fwd_tc_batched(n, g, x, w, bias, gate, residual, act, y, as_stream(stream));
            if (r != MVAE_ERR_UNSUPPORTED) return r;
        }
    }
    for (int l = 0; l < n; ++l)
        if (int e = mvae_conv2d_fwd(d + l, x[l], w[l], opt(bias, l), opt(gate, l), opt(residual, l), act, y[l], stream)) return e;
    return MVAE_OK;
}

extern "C" int mvae_conv2d_dgrad_batched(int n, const mvae_conv_desc* d, const float* const* dy, const float* const* w,
                                         const float* const* bias, const float* const* residual,
                                         const float* const* act_out, int act, float* const* dx, mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && d && dy && w && dx, "conv2d_dgrad_batched: bad arguments");
    if (n <= 8 && d[0].precision == MVAE_PREC_TF32) {
        ConvGeom g[8];
        bool ok = true;
        for (int l = 0; l < n && ok; ++l) ok = make_geom(d + l, g[l]) == MVAE_OK && d[l].precision == MVAE_PREC_TF32 && dy[l] && w[l] && dx[l];
        if (ok) {
            const int r = conv_dgrad_tc_batched(n, g, dy, w, bias, residual, act_out, act, dx, as_stream(stream));
            if (r != MVAE_ERR_UNSUPPORTED) return r;
        }
    }
    for (int l = 0; l < n; ++l)
        if (int e = mvae_conv2d_dgrad(d + l, dy[l], w[l], opt(bias, l), opt(residual, l), opt(act_out, l), act, dx[l], stream)) return e;
    return MVAE_OK;
}

extern "C" int mvae_conv2d_wgrad_batched(int n, const mvae_conv_desc* d, const float* const* x, const float* const* gate,
                                         const float* const* dy, float* const* dw, float* const* dbias,
                                         mvae_stream_t stream) {
    MVAE_REQUIRE(n > 0 && d && x && dy && dw, "conv2d_wgrad_batched: bad arguments");
    if (n <= 8 && d[0].precision == MVAE_PREC_TF32) {
        ConvGeom g[8];
        bool ok = true;
        for (int l = 0; l < n && ok; ++l) ok = make_geom(d + l, g[l]) == MVAE_OK && d[l].precision == MVAE_PREC_TF32 && x[l] && dy[l] && dw[l];
        if (ok) {
            const int r = conv_wgrad_tc_batched(n, g, x, gate, dy, dw, dbias, as_stream(stream));
            if (r != MVAE_ERR_UNSUPPORTED) return r;
        }
    }
    for (int l = 0; l < n; ++l)
        if (int e = mvae_conv2d_wgrad(d + l, x[l], opt(gate, l), dy[l], dw[l], dbias ? dbias[l] : nullptr, stream)) return e;
    return MVAE_OK;
}
