// Tensor-core softmax-attention cores for short sequences (S <= 16), one WARP per task.
//
// A task is one (sequence, head) pair when 8 < S <= 16, or TWO sequences of the same head packed into the two
// 8-row halves of the m16 tile when S <= 8 (cross-sample attention over 1+K <= 8 retrieved rows, movielens
// intra attention over 4 fields): the off-diagonal score blocks are masked.  All products are mma.sync m16n8k8
// TF32 with fp32 accumulate; the head dimension is zero-padded to a multiple of 8 by predicated fragment loads.
//
// Register-level trick (as in flash-attention): the C-fragment of the score tile (row g / g+8, cols 2t, 2t+1 of
// each 8-column block) is reused directly as the A-fragment of the following P.V (or dS.K) product by permuting
// the reduction index: k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1, and loading V (or K) rows in that order.
// The transposed products of the backward pass (dK = dS^T Q, dV = P^T dO) recompute S^T = K Q^T and dP^T = V dO^T
// with swapped operands instead of transposing fragments; per-row statistics travel through 32 floats of
// per-warp shared scratch.
#pragma once
#include "tile.cuh"

namespace rat {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnTaskMap {
    int S, nseq, seq0;
    bool packed;     // S <= 8: rows 0-7 = sequence seq0, rows 8-15 = sequence seq0+1
    // shared-memory row of fragment row r (0..15), or -1 if it does not exist
    __device__ __forceinline__ int row(int r) const {
        int seq, pos;
        if (packed) { seq = seq0 + (r >> 3); pos = r & 7; }
        else { seq = seq0; pos = r; }
        return (pos < S && seq < nseq) ? seq * S + pos : -1;
    }
    // may query row i attend to key row j ?
    __device__ __forceinline__ bool pair_ok(int i, int j) const { return !packed || ((i >> 3) == (j >> 3)); }
};

template <int DH>
__device__ __forceinline__ float ld_masked(const float* base, int row, int ld, int d) {
    return (row >= 0 && d < DH) ? base[(size_t)row * ld + d] : 0.f;
}

// C[16x16] (+)= X[16 x DHK] . Y[16 x DHK]^T  with X rows xr(g), xr(g+8) and Y rows given per 8-column block
template <int DH>
__device__ __forceinline__ void mma_xyT(float (&c)[2][4], const float* X, int ldx, int xlo, int xhi, const float* Y,
                                        int ldy, const int (&yr)[2], int t) {
    constexpr int KS = (DH + 7) / 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int d0 = ks * 8 + t, d1 = d0 + 4;
        unsigned a[4];
        a[0] = f2tf32(ld_masked<DH>(X, xlo, ldx, d0));
        a[1] = f2tf32(ld_masked<DH>(X, xhi, ldx, d0));
        a[2] = f2tf32(ld_masked<DH>(X, xlo, ldx, d1));
        a[3] = f2tf32(ld_masked<DH>(X, xhi, ldx, d1));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const unsigned b0 = f2tf32(ld_masked<DH>(Y, yr[nt], ldy, d0));
            const unsigned b1 = f2tf32(ld_masked<DH>(Y, yr[nt], ldy, d1));
            mma_tf32_16x8x8(c[nt], a, b0, b1);
        }
    }
}

// O[16 x DHK] = P[16x16] . Z[16 x DHK]  with P given as a score-tile C-fragment (permuted reduction index)
template <int DH>
__device__ __forceinline__ void mma_pz(float (&o)[(DH + 7) / 8][4], const float (&p)[2][4], const float* Z, int ldz,
                                       const AttnTaskMap& tm, int g, int t) {
    constexpr int NT = (DH + 7) / 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        unsigned a[4];
        a[0] = f2tf32(p[ks][0]); a[1] = f2tf32(p[ks][2]); a[2] = f2tf32(p[ks][1]); a[3] = f2tf32(p[ks][3]);
        const int z0 = tm.row(8 * ks + 2 * t), z1 = tm.row(8 * ks + 2 * t + 1);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int dc = 8 * nt + g;
            const unsigned b0 = f2tf32(ld_masked<DH>(Z, z0, ldz, dc));
            const unsigned b1 = f2tf32(ld_masked<DH>(Z, z1, ldz, dc));
            mma_tf32_16x8x8(o[nt], a, b0, b1);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// number of warp tasks of a tile
__device__ __forceinline__ int attn_mma_ntasks(int nseq_tile, int S, int hc) {
    return (S <= 8 ? (nseq_tile + 1) / 2 : nseq_tile) * hc;
}

// ---- forward: o = softmax(q k^T scale) v -> os ; natural-log logsumexp -> lse (optional) -------------------
template <int DH>
__device__ __forceinline__ void attn_fwd_mma(const float* __restrict__ qkv, int ld, int Cq, float* __restrict__ os,
                                             int ldo, float* __restrict__ lse, int nseq_tile, int S, int hc,
                                             float scale) {
    constexpr int NT = (DH + 7) / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntasks = attn_mma_ntasks(nseq_tile, S, hc);
    const float sl2 = scale * 1.4426950408889634f;
    for (int task = warp; task < ntasks; task += nwarps) {
        const int sp = task / hc, hl = task - sp * hc;
        AttnTaskMap tm{S, nseq_tile, S <= 8 ? 2 * sp : sp, S <= 8};
        const float* Q = qkv + hl * DH;
        const float* K = Q + Cq;
        const float* V = K + Cq;
        const int rlo = tm.row(g), rhi = tm.row(g + 8);
        const int kr[2] = {tm.row(g), tm.row(8 + g)};
        float sc[2][4] = {};
        mma_xyT<DH>(sc, Q, ld, rlo, rhi, K, ld, kr, t);
        // mask, scale, row softmax (rows g and g+8 of this lane; 4 columns each)
        float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = (e < 2) ? g : g + 8, j = 8 * nt + 2 * t + (e & 1);
                const bool ok = tm.row(j) >= 0 && tm.pair_ok(i, j);
                sc[nt][e] = ok ? sc[nt][e] * sl2 : -INFINITY;
                if (e < 2) mlo = fmaxf(mlo, sc[nt][e]); else mhi = fmaxf(mhi, sc[nt][e]);
            }
        mlo = quad_max(mlo); mhi = quad_max(mhi);
        if (rlo < 0) mlo = 0.f;
        if (rhi < 0) mhi = 0.f;
        float llo = 0.f, lhi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] = ex2_approx(sc[nt][e] - ((e < 2) ? mlo : mhi));
                if (e < 2) llo += sc[nt][e]; else lhi += sc[nt][e];
            }
        llo = quad_sum(llo); lhi = quad_sum(lhi);
        float o[NT][4] = {};
        mma_pz<DH>(o, sc, V, ld, tm, g, t);
        const float ilo = rlo >= 0 ? 1.0f / llo : 0.f, ihi = rhi >= 0 ? 1.0f / lhi : 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int dc = 8 * nt + 2 * t;
            if (rlo >= 0) {
                if (dc < DH) os[(size_t)rlo * ldo + hl * DH + dc] = o[nt][0] * ilo;
                if (dc + 1 < DH) os[(size_t)rlo * ldo + hl * DH + dc + 1] = o[nt][1] * ilo;
            }
            if (rhi >= 0) {
                if (dc < DH) os[(size_t)rhi * ldo + hl * DH + dc] = o[nt][2] * ihi;
                if (dc + 1 < DH) os[(size_t)rhi * ldo + hl * DH + dc + 1] = o[nt][3] * ihi;
            }
        }
        if (lse && t == 0) {
            if (rlo >= 0) lse[rlo * hc + hl] = (mlo + log2f(llo)) * 0.6931471805599453f;
            if (rhi >= 0) lse[rhi * hc + hl] = (mhi + log2f(lhi)) * 0.6931471805599453f;
        }
    }
}

// ---- backward: (q,k,v,do) -> dq,dk,dv written into dqkv (same column layout as qkv) ---------------------------
// wscr: per-warp scratch of 32 floats (log2-domain logsumexp and delta of the 16 rows)
template <int DH>
__device__ __forceinline__ void attn_bwd_mma(const float* __restrict__ qkv, float* __restrict__ dqkv, int ld, int Cq,
                                             const float* __restrict__ dos, int ldo, int nseq_tile, int S, int hc,
                                             float scale, float* __restrict__ wscr_all) {
    constexpr int NT = (DH + 7) / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntasks = attn_mma_ntasks(nseq_tile, S, hc);
    const float sl2 = scale * 1.4426950408889634f;
    float* wscr = wscr_all + warp * 32;
    for (int task = warp; task < ntasks; task += nwarps) {
        const int sp = task / hc, hl = task - sp * hc;
        AttnTaskMap tm{S, nseq_tile, S <= 8 ? 2 * sp : sp, S <= 8};
        const float* Q = qkv + hl * DH;
        const float* K = Q + Cq;
        const float* V = K + Cq;
        const float* DO = dos + hl * DH;
        float* DQ = dqkv + hl * DH;
        float* DK = DQ + Cq;
        float* DV = DK + Cq;
        const int rlo = tm.row(g), rhi = tm.row(g + 8);
        const int cr[2] = {tm.row(g), tm.row(8 + g)};       // rows used as the "column" operand (n = g)
        // ---- pass 1 (rows = queries i, cols = keys j): P, dP, delta, dS, dQ
        float sc[2][4] = {}, dp[2][4] = {};
        mma_xyT<DH>(sc, Q, ld, rlo, rhi, K, ld, cr, t);
        mma_xyT<DH>(dp, DO, ldo, rlo, rhi, V, ld, cr, t);
        float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = (e < 2) ? g : g + 8, j = 8 * nt + 2 * t + (e & 1);
                const bool ok = tm.row(j) >= 0 && tm.pair_ok(i, j);
                sc[nt][e] = ok ? sc[nt][e] * sl2 : -INFINITY;
                if (e < 2) mlo = fmaxf(mlo, sc[nt][e]); else mhi = fmaxf(mhi, sc[nt][e]);
            }
        mlo = quad_max(mlo); mhi = quad_max(mhi);
        if (rlo < 0) mlo = 0.f;
        if (rhi < 0) mhi = 0.f;
        float llo = 0.f, lhi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] = ex2_approx(sc[nt][e] - ((e < 2) ? mlo : mhi));
                if (e < 2) llo += sc[nt][e]; else lhi += sc[nt][e];
            }
        llo = quad_sum(llo); lhi = quad_sum(lhi);
        const float ilo = rlo >= 0 ? 1.0f / llo : 0.f, ihi = rhi >= 0 ? 1.0f / lhi : 0.f;
        float dlo = 0.f, dhi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] *= (e < 2) ? ilo : ihi;                       // P
                if (e < 2) dlo = fmaf(sc[nt][e], dp[nt][e], dlo); else dhi = fmaf(sc[nt][e], dp[nt][e], dhi);
            }
        dlo = quad_sum(dlo); dhi = quad_sum(dhi);                       // delta_i = sum_j P_ij dP_ij
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[nt][e] *= dp[nt][e] - ((e < 2) ? dlo : dhi);     // dS
        __syncwarp();
        if (t == 0) {
            wscr[g] = rlo >= 0 ? mlo + log2f(llo) : 0.f;  wscr[g + 8] = rhi >= 0 ? mhi + log2f(lhi) : 0.f;
            wscr[16 + g] = dlo;                            wscr[16 + g + 8] = dhi;
        }
        {
            float dq[NT][4] = {};
            mma_pz<DH>(dq, sc, K, ld, tm, g, t);                        // dQ = dS . K
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int dc = 8 * nt + 2 * t;
                if (rlo >= 0) {
                    if (dc < DH) DQ[(size_t)rlo * ld + dc] = dq[nt][0] * scale;
                    if (dc + 1 < DH) DQ[(size_t)rlo * ld + dc + 1] = dq[nt][1] * scale;
                }
                if (rhi >= 0) {
                    if (dc < DH) DQ[(size_t)rhi * ld + dc] = dq[nt][2] * scale;
                    if (dc + 1 < DH) DQ[(size_t)rhi * ld + dc + 1] = dq[nt][3] * scale;
                }
            }
        }
        __syncwarp();
        // ---- pass 2 (rows = keys j, cols = queries i): P^T, dP^T, dS^T, dK, dV
        float st[2][4] = {}, dpt[2][4] = {};
        mma_xyT<DH>(st, K, ld, rlo, rhi, Q, ld, cr, t);
        mma_xyT<DH>(dpt, V, ld, rlo, rhi, DO, ldo, cr, t);
        float pt[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = (e < 2) ? g : g + 8, i = 8 * nt + 2 * t + (e & 1);
                const bool ok = tm.row(i) >= 0 && tm.row(j) >= 0 && tm.pair_ok(i, j);
                const float p = ok ? ex2_approx(st[nt][e] * sl2 - wscr[i]) : 0.f;
                pt[nt][e] = p;                                           // P^T[j][i]
                st[nt][e] = p * (dpt[nt][e] - wscr[16 + i]);             // dS^T[j][i]
            }
        {
            float dk[NT][4] = {}, dv[NT][4] = {};
            mma_pz<DH>(dk, st, Q, ld, tm, g, t);                        // dK = dS^T . Q
            mma_pz<DH>(dv, pt, DO, ldo, tm, g, t);                      // dV = P^T . dO
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int dc = 8 * nt + 2 * t;
                if (rlo >= 0) {
                    if (dc < DH) { DK[(size_t)rlo * ld + dc] = dk[nt][0] * scale; DV[(size_t)rlo * ld + dc] = dv[nt][0]; }
                    if (dc + 1 < DH) { DK[(size_t)rlo * ld + dc + 1] = dk[nt][1] * scale; DV[(size_t)rlo * ld + dc + 1] = dv[nt][1]; }
                }
                if (rhi >= 0) {
                    if (dc < DH) { DK[(size_t)rhi * ld + dc] = dk[nt][2] * scale; DV[(size_t)rhi * ld + dc] = dv[nt][2]; }
                    if (dc + 1 < DH) { DK[(size_t)rhi * ld + dc + 1] = dk[nt][3] * scale; DV[(size_t)rhi * ld + dc + 1] = dv[nt][3]; }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace rat
