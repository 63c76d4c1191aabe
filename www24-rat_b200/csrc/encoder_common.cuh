// Pieces shared by the fused RAT-block forward and backward kernels: tile plans, weight staging, LayerNorm rows,
// and the SIMT softmax-attention cores (forward, forward-recompute and backward).
#pragma once
#include "tile.cuh"

namespace rat {

constexpr int ENC_THREADS = 512;

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct AttnPlan {
    int SPT;        // sequences per tile
    int hc;         // heads per chunk
    int nchunks;
    int Cq, Cq8;    // hc*dh and padded to 8
    int C3p8;       // pad8(3*Cq): q|k|v columns of one chunk
    int Dp8;        // pad8(D)
    int Dl, C3l, Cql, Il;   // leading dimensions (pad8 + 4: conflict-free mma fragment loads)
    int Rmax16;     // allocated rows
    int lg;         // lanes per LayerNorm row group
    int lpt;        // lanes per attention task
    int psize;      // (backward) floats in one CTA's partial-gradient record
    size_t smem_bytes;
};

static inline void fill_attn_plan(int S, int D, int H, int dh, int hc, AttnPlan* c) {
    c->hc = hc; c->nchunks = H / hc;
    c->Cq = hc * dh; c->Cq8 = pad8(c->Cq); c->C3p8 = pad8(3 * c->Cq);
    c->Dp8 = pad8(D);
    c->Dl = pad_ld(D); c->C3l = pad_ld(3 * c->Cq); c->Cql = pad_ld(c->Cq);
    c->Il = pad8(H * dh + 8) + 4;
    c->lg = min(32, next_pow2(D));
    c->lpt = min(32, next_pow2(S));
}

struct FFPlan { int RPT, Dp8, Mp8, Dl, Ml, lg; int psize; size_t smem_bytes; };

static inline void fill_ff_plan(int D, int M, FFPlan* p) {
    p->Dp8 = pad8(D); p->Mp8 = pad8(M); p->Dl = pad_ld(D); p->Ml = pad_ld(M);
    p->lg = min(32, next_pow2(D));
}

#ifdef __CUDACC__
__device__ __forceinline__ void zero_floats(float* p, size_t n) {
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0.f;
}
// zero rows [r0, r1) of a [.][ld] tile
__device__ __forceinline__ void zero_rows(float* p, int ld, int r0, int r1) {
    const int n = (r1 - r0) * ld;
    float* q = p + (size_t)r0 * ld;
    for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] = 0.f;
}
// dst[r][c] = W[r][c] for r < rows, c < cols ; zero elsewhere in the [rows_p][ld] tile
__device__ __forceinline__ void stage_padded(const float* __restrict__ W, int rows, int cols, int rows_p, int ld,
                                             float* __restrict__ dst) {
    const int total = rows_p * ld;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / ld, c = i - r * ld;
        dst[i] = (r < rows && c < cols) ? __ldg(W + (size_t)r * cols + c) : 0.f;
    }
}
// Wc[ch][c][d]: rows c = [q rows of chunk ch | k rows | v rows | zero pad], natural [.,D] layout, zero padded to Dl
__device__ __forceinline__ void stage_qkv_chunks(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                 const float* __restrict__ Wv, int D, const AttnPlan& p,
                                                 float* __restrict__ Wc) {
    const int per = p.C3p8 * p.Dl;
    const int total = p.nchunks * per;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int ch = i / per, rem = i - ch * per;
        const int c = rem / p.Dl, d = rem - c * p.Dl;
        float v = 0.f;
        if (d < D) {
            const int row0 = ch * p.Cq;
            if (c < p.Cq) v = __ldg(Wq + (size_t)(row0 + c) * D + d);
            else if (c < 2 * p.Cq) v = __ldg(Wk + (size_t)(row0 + c - p.Cq) * D + d);
            else if (c < 3 * p.Cq) v = __ldg(Wv + (size_t)(row0 + c - 2 * p.Cq) * D + d);
        }
        Wc[i] = v;
    }
}

// LayerNorm of R rows (global -> smem), eps=1e-5, biased variance (torch.nn.LayerNorm semantics); pad columns
// [D, Dp8) are zeroed; stats (optional) receives per-row (mean, rstd) for the backward pass.
// rowidx[r] = global token row of local row r (computed once per tile: no div/mod in the row loops)
__device__ __forceinline__ void fill_rowidx(long long* __restrict__ rowidx, const SeqGeom& g, long long s0, int R) {
    for (int r = threadIdx.x; r < R; r += blockDim.x) rowidx[r] = g.grow(s0 + r / g.S, r % g.S);
}

__device__ __forceinline__ void ln_rows_to_smem(const float* __restrict__ x, const long long* __restrict__ rowidx, int R,
                                                int D, int Dp8, const float* __restrict__ w,
                                                const float* __restrict__ b, float* __restrict__ dst, int ld, int lg,
                                                float* __restrict__ stats) {
    const int groups = blockDim.x / lg;
    const int gi = threadIdx.x / lg, li = threadIdx.x % lg;
    const float invD = 1.0f / (float)D;
    for (int r0 = 0; r0 < R; r0 += groups) {
        const int r = r0 + gi;
        const bool ok = r < R;
        const float* src = x;
        if (ok) src = x + rowidx[r] * D;
        float xv[4] = {0.f, 0.f, 0.f, 0.f};          // D <= 4*lg (D <= 128)
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (ok && d < D) { xv[k] = src[d]; sum += xv[k]; } }
        const float mean = group_sum(sum, lg) * invD;
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (ok && d < D) { float t = xv[k] - mean; sq = fmaf(t, t, sq); } }
        const float var = group_sum(sq, lg) * invD;
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        if (ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (d < D) dst[(size_t)r * ld + d] = (xv[k] - mean) * rstd * w[d] + b[d]; }
            for (int d = D + li; d < Dp8; d += lg) dst[(size_t)r * ld + d] = 0.f;
            if (stats && li == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
        }
    }
}

// ---- softmax(q k^T * scale) v ---------------------------------------------------------------------------------
// One lane per query row.  S <= 16: all scores live in registers (3 fully unrolled passes, independent dot
// products => ILP); larger S: online softmax.  o -> os ; natural-log logsumexp of the SCALED scores -> lse.
constexpr int ATT_SMAX = 16;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int DH>
__device__ __forceinline__ void attn_core(const float* __restrict__ qkv, int ld, int Cq, float* __restrict__ os,
                                          int ldo, float* __restrict__ lse, int nseq_tile, int S, int hc, int lpt,
                                          float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tpw = 32 / lpt;
    const int ntasks = nseq_tile * hc;
    const int sub = lane / lpt, li = lane % lpt;
    const float sl2 = scale * LOG2E;
    for (int task0 = warp * tpw; task0 < ntasks; task0 += nwarps * tpw) {
        const int task = task0 + sub;
        if (task >= ntasks) continue;
        const int ls = task / hc, hl = task - ls * hc;
        const float* base = qkv + (size_t)ls * S * ld + hl * DH;
        if (S <= ATT_SMAX) {
            if (li < S) {
                const float* qrow = base + (size_t)li * ld;
                float q[DH], sc[ATT_SMAX];
#pragma unroll
                for (int d = 0; d < DH; ++d) q[d] = qrow[d] * sl2;
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < ATT_SMAX; ++j) {
                    sc[j] = -INFINITY;
                    if (j < S) {
                        const float* krow = base + (size_t)j * ld + Cq;
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int d = 0; d + 1 < DH; d += 2) { s0 = fmaf(q[d], krow[d], s0); s1 = fmaf(q[d + 1], krow[d + 1], s1); }
                        if (DH & 1) s0 = fmaf(q[DH - 1], krow[DH - 1], s0);
                        sc[j] = s0 + s1;
                        m = fmaxf(m, sc[j]);
                    }
                }
                float l = 0.f;
#pragma unroll
                for (int j = 0; j < ATT_SMAX; ++j) { sc[j] = exp2f(sc[j] - m); l += sc[j]; }   // exp2(-inf) = 0
                float acc[DH];
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll
                for (int j = 0; j < ATT_SMAX; ++j) {
                    if (j < S) {
                        const float* vrow = base + (size_t)j * ld + 2 * Cq;
#pragma unroll
                        for (int d = 0; d < DH; ++d) acc[d] = fmaf(sc[j], vrow[d], acc[d]);
                    }
                }
                const float inv = 1.0f / l;
                float* orow = os + (size_t)(ls * S + li) * ldo + hl * DH;
#pragma unroll
                for (int d = 0; d < DH; ++d) orow[d] = acc[d] * inv;
                if (lse) lse[(ls * S + li) * hc + hl] = (m + log2f(l)) * LN2;
            }
        } else {
            for (int i = li; i < S; i += lpt) {
                float q[DH], acc[DH];
                const float* qrow = base + (size_t)i * ld;
#pragma unroll
                for (int d = 0; d < DH; ++d) { q[d] = qrow[d] * sl2; acc[d] = 0.f; }
                float m = -INFINITY, l = 0.f;
                for (int j = 0; j < S; ++j) {
                    const float* krow = base + (size_t)j * ld + Cq;
                    const float* vrow = krow + Cq;
                    float s = 0.f;
#pragma unroll
                    for (int d = 0; d < DH; ++d) s = fmaf(q[d], krow[d], s);
                    const float mn = fmaxf(m, s);
                    const float corr = exp2f(m - mn);
                    const float pj = exp2f(s - mn);
                    l = fmaf(l, corr, pj);
#pragma unroll
                    for (int d = 0; d < DH; ++d) acc[d] = fmaf(acc[d], corr, pj * vrow[d]);
                    m = mn;
                }
                const float inv = 1.0f / l;
                float* orow = os + (size_t)(ls * S + i) * ldo + hl * DH;
#pragma unroll
                for (int d = 0; d < DH; ++d) orow[d] = acc[d] * inv;
                if (lse) lse[(ls * S + i) * hc + hl] = (m + log2f(l)) * LN2;
            }
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
    const float sl2 = scale * LOG2E;
    for (int task0 = warp * tpw; task0 < ntasks; task0 += nwarps * tpw) {
        const int task = task0 + sub;
        if (task >= ntasks) continue;
        const int ls = task / hc, hl = task - ls * hc;
        const size_t row0 = (size_t)ls * S;
        const float* base = qkv + row0 * ld + hl * DH;
        float* dbase = dqkv + row0 * ld + hl * DH;
        const float* dobase = dos + row0 * ldo + hl * DH;
        const float* lse_t = lse + row0 * hc + hl;
        const float* del_t = delta + row0 * hc + hl;
        // ---- query role: dq_i = scale * sum_j dS_ij k_j
        for (int i = li; i < S; i += lpt) {
            float q[DH], dov[DH], dq[DH];
            const float* qrow = base + (size_t)i * ld;
            const float* dorow = dobase + (size_t)i * ldo;
#pragma unroll
            for (int d = 0; d < DH; ++d) { q[d] = qrow[d] * sl2; dov[d] = dorow[d]; dq[d] = 0.f; }
            const float Li = lse_t[(size_t)i * hc] * LOG2E, di = del_t[(size_t)i * hc];
#pragma unroll 2
            for (int j = 0; j < S; ++j) {
                const float* krow = base + (size_t)j * ld + Cq;
                const float* vrow = krow + Cq;
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) { s = fmaf(q[d], krow[d], s); dp = fmaf(dov[d], vrow[d], dp); }
                const float ds = exp2f(s - Li) * (dp - di);
#pragma unroll
                for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, krow[d], dq[d]);
            }
            float* dqrow = dbase + (size_t)i * ld;
#pragma unroll
            for (int d = 0; d < DH; ++d) dqrow[d] = dq[d] * scale;
        }
        // ---- key role: dk_j = scale * sum_i dS_ij q_i ; dv_j = sum_i P_ij do_i
        for (int j = li; j < S; j += lpt) {
            float k[DH], v[DH], dk[DH], dv[DH];
            const float* krow = base + (size_t)j * ld + Cq;
#pragma unroll
            for (int d = 0; d < DH; ++d) { k[d] = krow[d] * sl2; v[d] = krow[Cq + d]; dk[d] = 0.f; dv[d] = 0.f; }
#pragma unroll 2
            for (int i = 0; i < S; ++i) {
                const float* qrow = base + (size_t)i * ld;
                const float* dorow = dobase + (size_t)i * ldo;
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) { s = fmaf(qrow[d], k[d], s); dp = fmaf(dorow[d], v[d], dp); }
                const float p = exp2f(s - lse_t[(size_t)i * hc] * LOG2E);
                const float ds = p * (dp - del_t[(size_t)i * hc]);
#pragma unroll
                for (int d = 0; d < DH; ++d) { dk[d] = fmaf(ds, qrow[d], dk[d]); dv[d] = fmaf(p, dorow[d], dv[d]); }
            }
            float* dkrow = dbase + (size_t)j * ld + Cq;
#pragma unroll
            for (int d = 0; d < DH; ++d) { dkrow[d] = dk[d] * scale; dkrow[Cq + d] = dv[d]; }
        }
    }
}
#endif  // __CUDACC__

}  // namespace rat
