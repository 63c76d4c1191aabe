// K5: fused RAT-block backward kernels (recompute-in-shared-memory, flash-attention style).
//
//   k_attn_bwd : given x (input of the PreNorm+Attention sub-block) and dout (gradient of its output), recompute
//                LayerNorm / q|k|v / softmax statistics per head chunk in shared memory and produce
//                dx = base + alpha * dLN(...)  plus the CTA-private partial sums of dWq,dWk,dWv,dWo,dbo,dgamma,dbeta.
//   k_ff_bwd   : same for the FeedForward sub-block.
//   k_ln_bwd   : final-LayerNorm backward (RAT_m0/m1).
//
// These replace autograd's reverse of RAT_m2.py:155-236 (a11 in SURVEY.md 8a).  Weight gradients are accumulated
// in shared memory over all tiles a CTA owns (static tile->CTA map), written once to a per-CTA partial buffer
// and summed over CTAs in fixed order by k_reduce_partials  => bitwise run-to-run deterministic.
#include "tile.cuh"
#include "../../include/rat_b200.h"

namespace rat {

static inline int next_pow2_(int v) { int p = 1; while (p < v) p <<= 1; return p; }

constexpr int BWD_THREADS = 512;

struct AttnBwdPlan {
    int SPT, hc, Dp, Cq, Cqp, C3p, lg, lpt, nchunks;
    int psize;            // floats in one CTA's partial-gradient record
    size_t smem_bytes;
};

struct AttnBwdArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo;
    float* partials;                 // [gridDim.x][psize]
    long long nseq;
    SeqGeom g;
    int D, H, I;
    float scale, alpha;
    AttnBwdPlan p;
};

// ---- attention pieces ---------------------------------------------------------------------------------------
// forward recompute: o -> os, logsumexp -> lse   (q,k,v left intact)
template <int DH>
__device__ __forceinline__ void attn_fwd_recompute(const float* __restrict__ qkv, int ld, int Cq, float* __restrict__ os,
                                                   int ldo, float* __restrict__ lse, int nseq_tile, int S, int hc,
                                                   int lpt, float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tpw = 32 / lpt, ntasks = nseq_tile * hc, sub = lane / lpt, li = lane % lpt;
    for (int task0 = warp * tpw; task0 < ntasks; task0 += nwarps * tpw) {
        const int task = task0 + sub;
        if (task >= ntasks) continue;
        const int ls = task / hc, hl = task % hc;
        const float* base = qkv + (size_t)ls * S * ld + hl * DH;
        for (int i = li; i < S; i += lpt) {
            float q[DH], acc[DH];
            const float* qrow = base + (size_t)i * ld;
#pragma unroll
            for (int d = 0; d < DH; ++d) { q[d] = qrow[d] * scale; acc[d] = 0.f; }
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < S; ++j) {
                const float* krow = base + (size_t)j * ld + Cq;
                const float* vrow = krow + Cq;
                float sc = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) sc = fmaf(q[d], krow[d], sc);
                const float mn = fmaxf(m, sc), corr = expf(m - mn), pj = expf(sc - mn);
                l = fmaf(l, corr, pj);
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] = fmaf(acc[d], corr, pj * vrow[d]);
                m = mn;
            }
            const float inv = 1.0f / l;
            float* orow = os + (size_t)(ls * S + i) * ldo + hl * DH;
#pragma unroll
            for (int d = 0; d < DH; ++d) orow[d] = acc[d] * inv;
            lse[(ls * S + i) * hc + hl] = m + logf(l);
        }
    }
}

// softmax/attention backward.  Lane i as a query row -> dq_i ; lane j as a key row -> dk_j, dv_j.
//   P_ij = exp(scale q_i.k_j - L_i) ; dP_ij = do_i.v_j ; dS_ij = P_ij (dP_ij - delta_i)
template <int DH>
__device__ __forceinline__ void attn_bwd_core(const float* __restrict__ qkv, float* __restrict__ dqkv, int ld, int Cq,
                                              const float* __restrict__ dos, int ldo, const float* __restrict__ lse,
                                              const float* __restrict__ delta, int nseq_tile, int S, int hc, int lpt,
                                              float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tpw = 32 / lpt, ntasks = nseq_tile * hc, sub = lane / lpt, li = lane % lpt;
    for (int task0 = warp * tpw; task0 < ntasks; task0 += nwarps * tpw) {
        const int task = task0 + sub;
        if (task >= ntasks) continue;
        const int ls = task / hc, hl = task % hc;
        const size_t row0 = (size_t)ls * S;
        const float* base = qkv + row0 * ld + hl * DH;
        float* dbase = dqkv + row0 * ld + hl * DH;
        const float* dobase = dos + row0 * ldo + hl * DH;
        // ---- query role: dq_i
        for (int i = li; i < S; i += lpt) {
            float q[DH], dov[DH], dq[DH];
            const float* qrow = base + (size_t)i * ld;
            const float* dorow = dobase + (size_t)i * ldo;
#pragma unroll
            for (int d = 0; d < DH; ++d) { q[d] = qrow[d] * scale; dov[d] = dorow[d]; dq[d] = 0.f; }
            const float Li = lse[(row0 + i) * hc + hl], di = delta[(row0 + i) * hc + hl];
            for (int j = 0; j < S; ++j) {
                const float* krow = base + (size_t)j * ld + Cq;
                const float* vrow = krow + Cq;
                float sc = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) { sc = fmaf(q[d], krow[d], sc); dp = fmaf(dov[d], vrow[d], dp); }
                const float ds = expf(sc - Li) * (dp - di);
#pragma unroll
                for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, krow[d], dq[d]);
            }
            float* dqrow = dbase + (size_t)i * ld;
#pragma unroll
            for (int d = 0; d < DH; ++d) dqrow[d] = dq[d] * scale;
        }
        // ---- key role: dk_j, dv_j
        for (int j = li; j < S; j += lpt) {
            float k[DH], v[DH], dk[DH], dv[DH];
            const float* krow = base + (size_t)j * ld + Cq;
#pragma unroll
            for (int d = 0; d < DH; ++d) { k[d] = krow[d]; v[d] = krow[Cq + d]; dk[d] = 0.f; dv[d] = 0.f; }
            for (int i = 0; i < S; ++i) {
                const float* qrow = base + (size_t)i * ld;
                const float* dorow = dobase + (size_t)i * ldo;
                float sc = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) { sc = fmaf(qrow[d], k[d], sc); dp = fmaf(dorow[d], v[d], dp); }
                const float p = expf(sc * scale - lse[(row0 + i) * hc + hl]);
                const float ds = p * (dp - delta[(row0 + i) * hc + hl]);
#pragma unroll
                for (int d = 0; d < DH; ++d) { dk[d] = fmaf(ds, qrow[d], dk[d]); dv[d] = fmaf(p, dorow[d], dv[d]); }
            }
            float* dkrow = dbase + (size_t)j * ld + Cq;
#pragma unroll
            for (int d = 0; d < DH; ++d) { dkrow[d] = dk[d] * scale; dkrow[Cq + d] = dv[d]; }
        }
    }
}

// stage W chunk in natural layout: Wn[c][d] (c in q|k|v chunk columns, zero padded to C3p x Dp)
__device__ __forceinline__ void stage_qkv_natural(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                  const float* __restrict__ Wv, int D, int Dp, int row0, int Cq,
                                                  int C3p, float* __restrict__ Wn) {
    const int total = C3p * Dp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int d = i % Dp, c = i / Dp;
        float v = 0.f;
        if (d < D) {
            if (c < Cq) v = __ldg(Wq + (size_t)(row0 + c) * D + d);
            else if (c < 2 * Cq) v = __ldg(Wk + (size_t)(row0 + c - Cq) * D + d);
            else if (c < 3 * Cq) v = __ldg(Wv + (size_t)(row0 + c - 2 * Cq) * D + d);
        }
        Wn[i] = v;
    }
}
// WoN[d][c] = Wo[d][col0 + c]  (d < D rows, c < Cqp cols zero padded)
__device__ __forceinline__ void stage_out_natural(const float* __restrict__ Wo, int D, int I, int col0, int Cq, int Cqp,
                                                  float* __restrict__ WoN) {
    const int total = D * Cqp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % Cqp, d = i / Cqp;
        WoN[i] = (c < Cq) ? __ldg(Wo + (size_t)d * I + col0 + c) : 0.f;
    }
}
__device__ __forceinline__ void stage_qkv_T(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                            const float* __restrict__ Wv, int D, int row0, int Cq, int C3p,
                                            float* __restrict__ Wt) {
    const int total = C3p * D;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % D, c = i / D;
        float v = 0.f;
        if (c < Cq) v = __ldg(Wq + (size_t)(row0 + c) * D + k);
        else if (c < 2 * Cq) v = __ldg(Wk + (size_t)(row0 + c - Cq) * D + k);
        else if (c < 3 * Cq) v = __ldg(Wv + (size_t)(row0 + c - 2 * Cq) * D + k);
        Wt[(size_t)k * C3p + c] = v;
    }
}

// LayerNorm backward over the R rows of a tile + deterministic accumulation of per-column sums.
//   g = grad wrt LN output (smem, ld) ; dx[gr] = base[gr] + rstd*(g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
//   acc3[0][d] += sum_r g*xhat (dgamma), acc3[1][d] += sum_r g (dbeta), acc3[2][d] += sum_r extra[r][d] (optional)
__device__ __forceinline__ void ln_bwd_rows(const float* __restrict__ x, const float* __restrict__ base,
                                            float* __restrict__ dx, const SeqGeom& g, long long s0, int R, int D, int Dp,
                                            const float* __restrict__ gamma, const float* __restrict__ gsm, int ldg,
                                            const float* __restrict__ stats, const float* __restrict__ extra,
                                            int ldx, int lg, float* __restrict__ scratch, float* __restrict__ acc3) {
    const int groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
    const float invD = 1.0f / (float)D;
    constexpr int MAXPER = 4;                 // D <= 128
    float pg[MAXPER], pb[MAXPER], pe[MAXPER];
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) { pg[k] = 0.f; pb[k] = 0.f; pe[k] = 0.f; }
    for (int r0 = 0; r0 < R; r0 += groups) {
        const int r = r0 + gi;
        const bool ok = r < R;
        long long gr = 0;
        float mean = 0.f, rstd = 0.f;
        if (ok) { gr = g.grow(s0 + r / g.S, r % g.S); mean = stats[2 * r]; rstd = stats[2 * r + 1]; }
        float s1 = 0.f, s2 = 0.f;
        float xh[MAXPER], gg[MAXPER];
#pragma unroll
        for (int k = 0; k < MAXPER; ++k) {
            const int d = li + k * lg;
            xh[k] = 0.f; gg[k] = 0.f;
            if (ok && d < D) {
                xh[k] = (x[gr * D + d] - mean) * rstd;
                const float gv = gsm[(size_t)r * ldg + d];
                gg[k] = gv * gamma[d];
                s1 += gg[k];
                s2 = fmaf(gg[k], xh[k], s2);
                pg[k] = fmaf(gv, xh[k], pg[k]);
                pb[k] += gv;
                if (extra) pe[k] += extra[(size_t)r * ldx + d];
            }
        }
        s1 = group_sum(s1, lg) * invD;
        s2 = group_sum(s2, lg) * invD;
        if (ok) {
#pragma unroll
            for (int k = 0; k < MAXPER; ++k) {
                const int d = li + k * lg;
                if (d < D) {
                    float v = rstd * (gg[k] - s1 - xh[k] * s2);
                    if (base) v += base[gr * D + d];
                    dx[gr * D + d] = v;
                }
            }
        }
    }
    // deterministic cross-group reduction through shared scratch [groups][3][Dp]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXPER; ++k) {
        const int d = li + k * lg;
        if (d < D) {
            scratch[(gi * 3 + 0) * Dp + d] = pg[k];
            scratch[(gi * 3 + 1) * Dp + d] = pb[k];
            scratch[(gi * 3 + 2) * Dp + d] = pe[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
        const int which = i / D, d = i % D;
        float s = 0.f;
        for (int q = 0; q < groups; ++q) s += scratch[(q * 3 + which) * Dp + d];
        acc3[which * Dp + d] += s;
    }
}

template <int DH>
__global__ void __launch_bounds__(BWD_THREADS, 1) k_attn_bwd(AttnBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const AttnBwdPlan& p = a.p;
    const int S = a.g.S, D = a.D, Dp = p.Dp, C3p = p.C3p, Cq = p.Cq, Cqp = p.Cqp;
    const int Rmax = p.SPT * S;
    float* as = smem;                                  // [Rmax][Dp] LN(x)
    float* da = as + (size_t)Rmax * Dp;                // [Rmax][Dp] grad wrt LN output
    float* dys = da + (size_t)Rmax * Dp;               // [Rmax][Dp] alpha*dout
    float* qkv = dys + (size_t)Rmax * Dp;              // [Rmax][C3p]
    float* dqkv = qkv + (size_t)Rmax * C3p;            // [Rmax][C3p]
    float* os = dqkv + (size_t)Rmax * C3p;             // [Rmax][Cqp]
    float* dos = os + (size_t)Rmax * Cqp;              // [Rmax][Cqp]
    float* stats = dos + (size_t)Rmax * Cqp;           // [Rmax][2]  mean, rstd
    float* lse = stats + (size_t)round_up(2 * Rmax, 4);        // [Rmax][hc]
    float* delta = lse + (size_t)round_up(Rmax * p.hc, 4);     // [Rmax][hc]
    float* Wt = delta + (size_t)round_up(Rmax * p.hc, 4);      // [D][C3p]
    float* Wn = Wt + (size_t)D * C3p;                  // [C3p][Dp]
    float* WoN = Wn + (size_t)C3p * Dp;                // [D][Cqp]
    float* gW = WoN + (size_t)D * Cqp;                 // [nchunks][C3p][Dp]
    float* gWo = gW + (size_t)p.nchunks * C3p * Dp;    // [nchunks][Dp][Cqp]
    float* g3 = gWo + (size_t)p.nchunks * Dp * Cqp;    // [3][Dp] dgamma, dbeta, dbo
    float* scratch = g3 + 3 * Dp;                      // [groups][3][Dp]
    {
        const int nacc = p.nchunks * C3p * Dp + p.nchunks * Dp * Cqp + 3 * Dp;
        for (int i = threadIdx.x; i < nacc; i += blockDim.x) gW[i] = 0.f;
    }
    const long long ntiles = (a.nseq + p.SPT - 1) / p.SPT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * p.SPT;
        const int nseq_t = (int)min((long long)p.SPT, a.nseq - s0);
        const int R = nseq_t * S;
        __syncthreads();
        {
            // LayerNorm forward recompute (as + stats)
            const int lg = p.lg, groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
            const float invD = 1.0f / (float)D;
            for (int r0 = 0; r0 < R; r0 += groups) {
                const int r = r0 + gi;
                const bool ok = r < R;
                const float* src = a.x;
                if (ok) src = a.x + a.g.grow(s0 + r / S, r % S) * D;
                float sum = 0.f;
                if (ok) for (int d = li; d < D; d += lg) sum += src[d];
                const float mean = group_sum(sum, lg) * invD;
                float sq = 0.f;
                if (ok) for (int d = li; d < D; d += lg) { float t = src[d] - mean; sq = fmaf(t, t, sq); }
                const float rstd = 1.0f / sqrtf(group_sum(sq, lg) * invD + 1e-5f);
                if (ok) {
                    for (int d = li; d < D; d += lg)
                        as[(size_t)r * Dp + d] = (src[d] - mean) * rstd * a.ln_w[d] + a.ln_b[d];
                    for (int d = D + li; d < Dp; d += lg) as[(size_t)r * Dp + d] = 0.f;
                    if (li == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
                }
            }
        }
        for (int i = threadIdx.x; i < R * Dp; i += blockDim.x) {
            const int r = i / Dp, d = i % Dp;
            da[i] = 0.f;
            float v = 0.f;
            if (d < D) v = a.alpha * a.dout[a.g.grow(s0 + r / S, r % S) * D + d];
            dys[i] = v;
        }
        for (int ch = 0; ch < p.nchunks; ++ch) {
            const int row0 = ch * Cq;
            __syncthreads();
            stage_qkv_T(a.Wq, a.Wk, a.Wv, D, row0, Cq, C3p, Wt);
            stage_qkv_natural(a.Wq, a.Wk, a.Wv, D, Dp, row0, Cq, C3p, Wn);
            stage_out_natural(a.Wo, D, a.I, row0, Cq, Cqp, WoN);
            __syncthreads();
            // (1) qkv = as . Wt ; (3) dos = dys . WoN
            tile_gemm<4>(as, Dp, Wt, C3p, qkv, C3p, R, C3p, D, false, EpiNone());
            tile_gemm<4>(dys, Dp, WoN, Cqp, dos, Cqp, R, Cqp, D, false, EpiNone());
            __syncthreads();
            // (2) attention forward recompute
            attn_fwd_recompute<DH>(qkv, C3p, Cq, os, Cqp, lse, nseq_t, S, p.hc, p.lpt, a.scale);
            if (Cqp > Cq)
                for (int i = threadIdx.x; i < R * (Cqp - Cq); i += blockDim.x)
                    os[(size_t)(i / (Cqp - Cq)) * Cqp + Cq + i % (Cqp - Cq)] = 0.f;
            __syncthreads();
            // delta[r][hl] = do . o
            for (int i = threadIdx.x; i < R * p.hc; i += blockDim.x) {
                const int r = i / p.hc, hl = i % p.hc;
                const float* o = os + (size_t)r * Cqp + hl * DH;
                const float* dd = dos + (size_t)r * Cqp + hl * DH;
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) s = fmaf(o[d], dd[d], s);
                delta[i] = s;
            }
            // (4) gWo[ch] += dys^T . os      [Dp][Cqp]
            tile_gemm_tn_acc(dys, Dp, os, Cqp, gWo + (size_t)ch * Dp * Cqp, Cqp, R, Dp, Cqp);
            __syncthreads();
            // (5) attention backward -> dqkv (zero the pad columns first)
            if (C3p > 3 * Cq)
                for (int i = threadIdx.x; i < R * (C3p - 3 * Cq); i += blockDim.x)
                    dqkv[(size_t)(i / (C3p - 3 * Cq)) * C3p + 3 * Cq + i % (C3p - 3 * Cq)] = 0.f;
            attn_bwd_core<DH>(qkv, dqkv, C3p, Cq, dos, Cqp, lse, delta, nseq_t, S, p.hc, p.lpt, a.scale);
            __syncthreads();
            // (6) gW[ch] += dqkv^T . as      [C3p][Dp]   ; (7) da += dqkv . Wn
            tile_gemm_tn_acc(dqkv, C3p, as, Dp, gW + (size_t)ch * C3p * Dp, Dp, R, C3p, Dp);
            tile_gemm<4>(dqkv, C3p, Wn, Dp, da, Dp, R, Dp, C3p, true, EpiNone());
        }
        __syncthreads();
        ln_bwd_rows(a.x, a.base, a.dx, a.g, s0, R, D, Dp, a.ln_w, da, Dp, stats, dys, Dp, p.lg, scratch, g3);
    }
    __syncthreads();
    // ---- flush the CTA-private partial sums in natural parameter layout:
    //      [dWq I*D | dWk I*D | dWv I*D | dWo D*I | dbo D | dgamma D | dbeta D]
    float* out = a.partials + (size_t)blockIdx.x * p.psize;
    const int I = a.I;
    for (int i = threadIdx.x; i < 3 * I * D; i += blockDim.x) {
        const int which = i / (I * D), rem = i % (I * D);
        const int row = rem / D, d = rem % D;
        const int ch = row / Cq, c = row % Cq;
        out[i] = gW[((size_t)ch * C3p + which * Cq + c) * Dp + d];
    }
    float* o2 = out + 3 * I * D;
    for (int i = threadIdx.x; i < D * I; i += blockDim.x) {
        const int d = i / I, col = i % I;
        const int ch = col / Cq, c = col % Cq;
        o2[i] = gWo[((size_t)ch * Dp + d) * Cqp + c];
    }
    float* o3 = o2 + D * I;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        o3[i] = g3[2 * Dp + i];            // dbo  (sum of alpha*dout)
        o3[D + i] = g3[0 * Dp + i];        // dgamma
        o3[2 * D + i] = g3[1 * Dp + i];    // dbeta
    }
}

// ------------------------------------------------------------------------------------------------------------
struct FFBwdPlan { int RPT, Dp, Mp, lg; int psize; size_t smem_bytes; };
struct FFBwdArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* W1; const float* b1; const float* W2;
    float* partials;
    long long rows;
    int D, M;
    FFBwdPlan p;
};

struct EpiBias {
    const float* b;
    __device__ __forceinline__ void operator()(int, int c0, float4& v) const {
        v.x += b[c0]; v.y += b[c0 + 1]; v.z += b[c0 + 2]; v.w += b[c0 + 3];
    }
};
// v = dh ; hs holds pre-activation: hs <- gelu(pre), v <- dh * gelu'(pre)
struct EpiGeluBwd {
    float* hs; int ld;
    __device__ __forceinline__ void operator()(int r, int c0, float4& v) const {
        float4* hp = reinterpret_cast<float4*>(hs + (size_t)r * ld + c0);
        const float4 pre = *hp;
        v.x *= gelu_erf_grad(pre.x); v.y *= gelu_erf_grad(pre.y); v.z *= gelu_erf_grad(pre.z); v.w *= gelu_erf_grad(pre.w);
        *hp = make_float4(gelu_erf(pre.x), gelu_erf(pre.y), gelu_erf(pre.z), gelu_erf(pre.w));
    }
};

__device__ __forceinline__ void stage_natural(const float* __restrict__ W, int Rw, int Cw, int Cp, float* __restrict__ dst) {
    const int total = Rw * Cp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % Cp, r = i / Cp;
        dst[i] = (c < Cw) ? __ldg(W + (size_t)r * Cw + c) : 0.f;
    }
}
__device__ __forceinline__ void stage_T(const float* __restrict__ W, int C, int K, int Cp, float* __restrict__ Wt) {
    const int total = Cp * K;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % K, c = i / K;
        Wt[(size_t)k * Cp + c] = (c < C) ? __ldg(W + (size_t)c * K + k) : 0.f;
    }
}

__global__ void __launch_bounds__(BWD_THREADS, 1) k_ff_bwd(FFBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const FFBwdPlan& p = a.p;
    const int D = a.D, M = a.M, Dp = p.Dp, Mp = p.Mp;
    float* xs = smem;                                   // [RPT][Dp]  FF input (LN(x) or x)
    float* dys = xs + (size_t)p.RPT * Dp;               // [RPT][Dp]  dout
    float* ys = dys + (size_t)p.RPT * Dp;               // [RPT][Dp]  grad wrt FF input
    float* hs = ys + (size_t)p.RPT * Dp;                // [RPT][Mp]  pre -> h
    float* dhs = hs + (size_t)p.RPT * Mp;               // [RPT][Mp]  dh -> dpre
    float* stats = dhs + (size_t)p.RPT * Mp;            // [RPT][2]
    float* W1t = stats + (size_t)round_up(2 * p.RPT, 4);   // [D][Mp]   (k-major of W1 [M,D])
    float* W2n = W1t + (size_t)D * Mp;                  // [D][Mp]   natural W2 [D,M]
    float* W1n = W2n + (size_t)D * Mp;                  // [M][Dp]   natural W1 [M,D]
    float* b1s = W1n + (size_t)M * Dp;                  // [Mp]
    float* gW1 = b1s + Mp;                              // [Mp][Dp]
    float* gW2 = gW1 + (size_t)Mp * Dp;                 // [Dp][Mp]
    float* gb1 = gW2 + (size_t)Dp * Mp;                 // [Mp]
    float* g3 = gb1 + Mp;                               // [3][Dp]  dgamma, dbeta, db2
    float* scratch = g3 + 3 * Dp;                       // [groups][3][Dp]
    stage_T(a.W1, M, D, Mp, W1t);
    stage_natural(a.W2, D, M, Mp, W2n);
    stage_natural(a.W1, M, D, Dp, W1n);
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) b1s[i] = i < M ? a.b1[i] : 0.f;
    {
        const int nacc = 2 * Mp * Dp + Mp + 3 * Dp;
        for (int i = threadIdx.x; i < nacc; i += blockDim.x) gW1[i] = 0.f;
    }
    const long long ntiles = (a.rows + p.RPT - 1) / p.RPT;
    SeqGeom flat{1, 0, 1, 1};
    const int lg = p.lg, groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * p.RPT;
        const int R = (int)min((long long)p.RPT, a.rows - r0);
        __syncthreads();
        if (a.ln_w) {
            const float invD = 1.0f / (float)D;
            for (int rr = 0; rr < R; rr += groups) {
                const int r = rr + gi;
                const bool ok = r < R;
                const float* src = a.x + (ok ? (r0 + r) : 0) * D;
                float sum = 0.f;
                if (ok) for (int d = li; d < D; d += lg) sum += src[d];
                const float mean = group_sum(sum, lg) * invD;
                float sq = 0.f;
                if (ok) for (int d = li; d < D; d += lg) { float t = src[d] - mean; sq = fmaf(t, t, sq); }
                const float rstd = 1.0f / sqrtf(group_sum(sq, lg) * invD + 1e-5f);
                if (ok) {
                    for (int d = li; d < D; d += lg) xs[(size_t)r * Dp + d] = (src[d] - mean) * rstd * a.ln_w[d] + a.ln_b[d];
                    for (int d = D + li; d < Dp; d += lg) xs[(size_t)r * Dp + d] = 0.f;
                    if (li == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
                }
            }
        } else {
            for (int i = threadIdx.x; i < R * Dp; i += blockDim.x) {
                const int r = i / Dp, d = i % Dp;
                xs[i] = d < D ? a.x[(r0 + r) * D + d] : 0.f;
            }
        }
        for (int i = threadIdx.x; i < R * Dp; i += blockDim.x) {
            const int r = i / Dp, d = i % Dp;
            dys[i] = d < D ? a.dout[(r0 + r) * D + d] : 0.f;
        }
        __syncthreads();
        tile_gemm<4>(xs, Dp, W1t, Mp, hs, Mp, R, Mp, D, false, EpiBias{b1s});            // pre
        __syncthreads();
        tile_gemm<4>(dys, Dp, W2n, Mp, dhs, Mp, R, Mp, D, false, EpiGeluBwd{hs, Mp});    // dpre ; hs <- h
        __syncthreads();
        tile_gemm_tn_acc(dys, Dp, hs, Mp, gW2, Mp, R, Dp, Mp);                           // dW2 [D][M]
        tile_gemm_tn_acc(dhs, Mp, xs, Dp, gW1, Dp, R, Mp, Dp);                           // dW1 [M][D]
        tile_colsum_acc(dhs, Mp, gb1, R, Mp);                                            // db1
        tile_gemm<4>(dhs, Mp, W1n, Dp, ys, Dp, R, Dp, M, false, EpiNone());              // grad wrt FF input
        __syncthreads();
        if (a.ln_w) {
            ln_bwd_rows(a.x, a.base, a.dx, flat, r0, R, D, Dp, a.ln_w, ys, Dp, stats, dys, Dp, lg, scratch, g3);
        } else {
            // dx = base + ys ; db2 += colsum(dys)  (deterministic per-group partials as in ln_bwd_rows)
            float pe[4] = {0.f, 0.f, 0.f, 0.f};
            for (int rr = 0; rr < R; rr += groups) {
                const int r = rr + gi;
                if (r < R) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int d = li + k * lg;
                        if (d < D) {
                            float v = ys[(size_t)r * Dp + d];
                            if (a.base) v += a.base[(r0 + r) * D + d];
                            a.dx[(r0 + r) * D + d] = v;
                            pe[k] += dys[(size_t)r * Dp + d];
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (d < D) scratch[gi * Dp + d] = pe[k]; }
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += blockDim.x) {
                float s = 0.f;
                for (int q = 0; q < groups; ++q) s += scratch[q * Dp + d];
                g3[2 * Dp + d] += s;
            }
        }
    }
    __syncthreads();
    // flush: [dW1 M*D | db1 M | dW2 D*M | db2 D | dgamma D | dbeta D]
    float* out = a.partials + (size_t)blockIdx.x * p.psize;
    for (int i = threadIdx.x; i < M * D; i += blockDim.x) out[i] = gW1[(size_t)(i / D) * Dp + i % D];
    for (int i = threadIdx.x; i < M; i += blockDim.x) out[M * D + i] = gb1[i];
    float* o2 = out + M * D + M;
    for (int i = threadIdx.x; i < D * M; i += blockDim.x) o2[i] = gW2[(size_t)(i / M) * Mp + i % M];
    float* o3 = o2 + D * M;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        o3[i] = g3[2 * Dp + i];
        o3[D + i] = g3[0 * Dp + i];
        o3[2 * D + i] = g3[1 * Dp + i];
    }
}

// out[i] (+)= sum_c partials[c][off + i]   for up to 8 destination segments; fixed CTA order => deterministic
struct ReduceSeg { float* dst; int off; int len; int accumulate; };
struct ReduceArgs { const float* partials; int nparts; int psize; int nseg; ReduceSeg seg[8]; };
__global__ void k_reduce_partials(ReduceArgs a) {
    int total = 0;
    for (int s = 0; s < a.nseg; ++s) total += a.seg[s].len;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int s = 0, j = i;
        while (j >= a.seg[s].len) { j -= a.seg[s].len; ++s; }
        const float* p = a.partials + a.seg[s].off + j;
        float sum = 0.f;
        for (int c = 0; c < a.nparts; ++c) sum += p[(size_t)c * a.psize];
        if (a.seg[s].dst) {
            if (a.seg[s].accumulate) a.seg[s].dst[j] += sum;
            else a.seg[s].dst[j] = sum;
        }
    }
}

// final LayerNorm backward (RAT_m0/m1): dx = LNbwd(dout) ; per-CTA partial dgamma/dbeta
__global__ void __launch_bounds__(256) k_ln_bwd(const float* __restrict__ x, const float* __restrict__ dout,
                                                float* __restrict__ dx, const float* __restrict__ w, long long rows,
                                                int D, int lg, float* __restrict__ partials) {
    extern __shared__ __align__(16) float sm[];
    const int groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
    const float invD = 1.0f / (float)D;
    float pg[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0};
    for (long long r0 = (long long)blockIdx.x * groups; r0 < rows; r0 += (long long)gridDim.x * groups) {
        const long long r = r0 + gi;
        const bool ok = r < rows;
        const float* src = x + (ok ? r : 0) * D;
        float sum = 0.f;
        if (ok) for (int d = li; d < D; d += lg) sum += src[d];
        const float mean = group_sum(sum, lg) * invD;
        float sq = 0.f;
        if (ok) for (int d = li; d < D; d += lg) { float t = src[d] - mean; sq = fmaf(t, t, sq); }
        const float rstd = 1.0f / sqrtf(group_sum(sq, lg) * invD + 1e-5f);
        float s1 = 0.f, s2 = 0.f, xh[4], gg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int d = li + k * lg;
            xh[k] = gg[k] = 0.f;
            if (ok && d < D) {
                xh[k] = (src[d] - mean) * rstd;
                const float gv = dout[r * D + d];
                gg[k] = gv * w[d];
                s1 += gg[k]; s2 = fmaf(gg[k], xh[k], s2);
                pg[k] = fmaf(gv, xh[k], pg[k]); pb[k] += gv;
            }
        }
        s1 = group_sum(s1, lg) * invD; s2 = group_sum(s2, lg) * invD;
        if (ok)
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (d < D) dx[r * D + d] = rstd * (gg[k] - s1 - xh[k] * s2); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (d < D) { sm[(gi * 2) * D + d] = pg[k]; sm[(gi * 2 + 1) * D + d] = pb[k]; } }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
        const int which = i / D, d = i % D;
        float s = 0.f;
        for (int q = 0; q < groups; ++q) s += sm[(q * 2 + which) * D + d];
        partials[(size_t)blockIdx.x * 2 * D + i] = s;
    }
}

// ---- planning ------------------------------------------------------------------------------------------------
static size_t attn_bwd_floats(int R, int S, int D, int Dp, int hc, int dh, int H, int* psize_out, AttnBwdPlan* pl) {
    const int Cq = hc * dh, Cqp = round_up(Cq, 4), C3p = round_up(3 * Cq, 4), nch = H / hc;
    const int lg = min(32, next_pow2_(D));
    const int groups = BWD_THREADS / lg;
    size_t fl = (size_t)R * (3 * Dp + 2 * C3p + 2 * Cqp) + round_up(2 * R, 4) + 2 * (size_t)round_up(R * hc, 4);
    fl += (size_t)D * C3p + (size_t)C3p * Dp + (size_t)D * Cqp;
    fl += (size_t)nch * C3p * Dp + (size_t)nch * Dp * Cqp + 3 * Dp + (size_t)groups * 3 * Dp;
    if (pl) { pl->hc = hc; pl->Dp = Dp; pl->Cq = Cq; pl->Cqp = Cqp; pl->C3p = C3p; pl->nchunks = nch; pl->lg = lg; }
    (void)S; (void)psize_out;
    return fl;
}

int plan_attn_bwd(int S, int D, int H, int dh, AttnBwdPlan* out) {
    const int Dp = round_up(D, 4);
    const size_t bud = (size_t)(max_smem_optin() - 1024) / 4;
    const int cap_spt = max(1, 160 / S);
    int bestR = 0;
    AttnBwdPlan best{};
    for (int hc = H; hc >= 1; --hc) {
        if (H % hc) continue;
        AttnBwdPlan cand{};
        int spt = cap_spt;
        for (; spt >= 1; --spt)
            if (attn_bwd_floats(spt * S, S, D, Dp, hc, dh, H, nullptr, &cand) <= bud) break;
        if (spt < 1) continue;
        const int R = spt * S;
        if (R > bestR) {
            bestR = R;
            best = cand;
            best.SPT = spt;
            best.smem_bytes = attn_bwd_floats(R, S, D, Dp, hc, dh, H, nullptr, nullptr) * 4;
        }
        if (R >= min(96, cap_spt * S)) break;
    }
    if (!bestR) return RAT_ESMEM;
    best.lpt = min(32, next_pow2_(S));
    const int I = H * dh;
    best.psize = round_up(3 * I * D + D * I + 3 * D, 4);
    *out = best;
    return RAT_OK;
}

int plan_ff_bwd(int D, int M, FFBwdPlan* out) {
    FFBwdPlan p{};
    p.Dp = round_up(D, 4); p.Mp = round_up(M, 4);
    p.lg = min(32, next_pow2_(D));
    const int groups = BWD_THREADS / p.lg;
    const size_t bud = (size_t)(max_smem_optin() - 1024) / 4;
    const size_t fixed = 2 * (size_t)D * p.Mp + (size_t)M * p.Dp + p.Mp + 2 * (size_t)p.Mp * p.Dp + p.Mp + 3 * p.Dp +
                         (size_t)groups * 3 * p.Dp + 8;
    int rpt = 160;
    for (; rpt >= 8; rpt -= 8) {
        size_t fl = (size_t)rpt * (3 * p.Dp + 2 * p.Mp) + round_up(2 * rpt, 4) + fixed;
        if (fl <= bud) break;
    }
    if (rpt < 8) return RAT_ESMEM;
    p.RPT = rpt;
    p.smem_bytes = ((size_t)rpt * (3 * p.Dp + 2 * p.Mp) + round_up(2 * rpt, 4) + fixed) * 4;
    p.psize = round_up(2 * M * D + M + 3 * D, 4);
    *out = p;
    return RAT_OK;
}

static int bwd_grid(long long ntiles) { return (int)min(ntiles, (long long)num_sms()); }

}  // namespace rat

using namespace rat;

static int run_reduce(const float* partials, int nparts, int psize, int nseg, const ReduceSeg* segs, cudaStream_t st) {
    ReduceArgs r{};
    r.partials = partials; r.nparts = nparts; r.psize = psize; r.nseg = nseg;
    int total = 0;
    for (int i = 0; i < nseg; ++i) { r.seg[i] = segs[i]; total += segs[i].len; }
    k_reduce_partials<<<max(1, min(ceil_div(total, 256), 1024)), 256, 0, st>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_partials");
    return RAT_OK;
}

extern "C" size_t rat_attn_bwd_workspace_bytes(int B, int T, int N, int D, int heads, int dim_head, int mode) {
    AttnBwdPlan p{};
    const int S = mode == 0 ? N : T;
    if (plan_attn_bwd(S, D, heads, dim_head, &p) != RAT_OK) return 0;
    const long long nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    const long long ntiles = (nseq + p.SPT - 1) / p.SPT;
    return (size_t)bwd_grid(ntiles) * p.psize * sizeof(float);
}

template <int DH>
static int launch_attn_bwd(const AttnBwdArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_bwd<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_bwd)");
        attr_set = true;
    }
    k_attn_bwd<DH><<<grid, BWD_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_bwd");
    return RAT_OK;
}

extern "C" int rat_attn_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                            const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                            float* dWq, float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b,
                            int accumulate_wq, int B, int T, int N, int D, int heads, int dim_head, float scale,
                            float alpha, int mode, float* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && N > 0 && D > 0 && heads > 0, "rat_attn_bwd: bad shape");
    RAT_REQUIRE(D <= 128, "rat_attn_bwd: D=%d > 128 not supported", D);
    AttnBwdArgs a{};
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.ln_w = ln_w; a.ln_b = ln_b;
    a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo;
    a.g.S = mode == 0 ? N : T; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.D = D; a.H = heads; a.I = heads * dim_head; a.scale = scale; a.alpha = alpha;
    int rc = plan_attn_bwd(a.g.S, D, heads, dim_head, &a.p);
    if (rc != RAT_OK) { set_error("rat_attn_bwd: sequence length %d x dim %d does not fit in shared memory", a.g.S, D); return rc; }
    const long long ntiles = (a.nseq + a.p.SPT - 1) / a.p.SPT;
    const int grid = bwd_grid(ntiles);
    RAT_REQUIRE(workspace && workspace_bytes >= (size_t)grid * a.p.psize * sizeof(float),
                "rat_attn_bwd: workspace too small (%zu < %zu)", workspace_bytes, (size_t)grid * a.p.psize * sizeof(float));
    a.partials = workspace;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim_head) {
        case 4: rc = launch_attn_bwd<4>(a, grid, st); break;
        case 8: rc = launch_attn_bwd<8>(a, grid, st); break;
        case 10: rc = launch_attn_bwd<10>(a, grid, st); break;
        case 16: rc = launch_attn_bwd<16>(a, grid, st); break;
        case 20: rc = launch_attn_bwd<20>(a, grid, st); break;
        case 32: rc = launch_attn_bwd<32>(a, grid, st); break;
        default: set_error("rat_attn_bwd: dim_head=%d not instantiated", dim_head); return RAT_EINVAL;
    }
    if (rc != RAT_OK) return rc;
    const int I = a.I;
    ReduceSeg segs[7] = {
        {dWq, 0, I * D, accumulate_wq}, {dWk, I * D, I * D, 0}, {dWv, 2 * I * D, I * D, 0},
        {dWo, 3 * I * D, D * I, 0}, {dbo, 4 * I * D, D, 0}, {dln_w, 4 * I * D + D, D, 0}, {dln_b, 4 * I * D + 2 * D, D, 0}};
    return run_reduce(workspace, grid, a.p.psize, 7, segs, st);
}

extern "C" size_t rat_ff_bwd_workspace_bytes(long long rows, int D, int M) {
    FFBwdPlan p{};
    if (plan_ff_bwd(D, M, &p) != RAT_OK) return 0;
    const long long ntiles = (rows + p.RPT - 1) / p.RPT;
    return (size_t)bwd_grid(ntiles) * p.psize * sizeof(float);
}

extern "C" int rat_ff_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                          const float* ln_b, const float* W1, const float* b1, const float* W2, float* dW1, float* db1,
                          float* dW2, float* db2, float* dln_w, float* dln_b, long long rows, int D, int M,
                          float* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && M > 0, "rat_ff_bwd: bad shape");
    RAT_REQUIRE(D <= 128, "rat_ff_bwd: D=%d > 128 not supported", D);
    FFBwdArgs a{};
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.W2 = W2;
    a.rows = rows; a.D = D; a.M = M;
    int rc = plan_ff_bwd(D, M, &a.p);
    if (rc != RAT_OK) { set_error("rat_ff_bwd: D=%d M=%d does not fit in shared memory", D, M); return rc; }
    const long long ntiles = (rows + a.p.RPT - 1) / a.p.RPT;
    const int grid = bwd_grid(ntiles);
    RAT_REQUIRE(workspace && workspace_bytes >= (size_t)grid * a.p.psize * sizeof(float), "rat_ff_bwd: workspace too small");
    a.partials = workspace;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_bwd)");
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_ff_bwd<<<grid, BWD_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_ff_bwd");
    ReduceSeg segs[6] = {{dW1, 0, M * D, 0}, {db1, M * D, M, 0}, {dW2, M * D + M, D * M, 0},
                         {db2, 2 * M * D + M, D, 0}, {dln_w, 2 * M * D + M + D, D, 0}, {dln_b, 2 * M * D + M + 2 * D, D, 0}};
    return run_reduce(workspace, grid, a.p.psize, 6, segs, st);
}

extern "C" size_t rat_layernorm_bwd_workspace_bytes(long long rows, int D) {
    (void)rows;
    return (size_t)num_sms() * 4 * 2 * D * sizeof(float);
}

extern "C" int rat_layernorm_bwd(const float* x, const float* dout, float* dx, const float* w, float* dw, float* db,
                                 long long rows, int D, float* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && D <= 128, "rat_layernorm_bwd: bad shape");
    const int lg = min(32, next_pow2_(D));
    const int groups = 256 / lg;
    const int grid = (int)min((rows + groups - 1) / groups, (long long)num_sms() * 4);
    RAT_REQUIRE(workspace && workspace_bytes >= (size_t)grid * 2 * D * sizeof(float), "rat_layernorm_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    k_ln_bwd<<<grid, 256, (size_t)groups * 2 * D * sizeof(float), st>>>(x, dout, dx, w, rows, D, lg, workspace);
    RAT_CHECK_LAUNCH("k_ln_bwd");
    ReduceSeg segs[2] = {{dw, 0, D, 0}, {db, D, D, 0}};
    return run_reduce(workspace, grid, 2 * D, 2, segs, st);
}
