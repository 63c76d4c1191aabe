// K5: fused RAT-block backward kernels (recompute-in-shared-memory, flash-attention style).
//
//   k_attn_bwd : given x (input of the PreNorm+Attention sub-block) and dout (gradient of its output), recompute
//                LayerNorm / q|k|v / softmax statistics per head chunk in shared memory and produce
//                dx = base + alpha * dLN(...)  plus the CTA-private partial sums of dWq,dWk,dWv,dWo,dbo,dgamma,dbeta.
//   k_ff_bwd   : same for the FeedForward sub-block.
//   k_ln_bwd   : final-LayerNorm backward (RAT_m0/m1).
//
// These replace autograd's reverse of RAT_m2.py:155-236 (a11 in SURVEY.md 8a).  One persistent CTA per SM keeps
// the sub-block's weights resident in shared memory (natural layout: the same copy is the n-major operand of the
// forward recompute and the k-major operand of the data-gradient product).  All seven products per head chunk run
// on the tensor cores (mma.sync TF32) or on the exact SIMT twin (precision=fp32).  Weight-gradient products
// (reduction over token rows) accumulate into a CTA-private record in global memory (L2 resident, every element
// owned by one thread, static tile->CTA map), and k_reduce_* sums the records over CTAs in fixed order
// => bitwise run-to-run deterministic.
#include <algorithm>
#include "tile.cuh"
#include "encoder_common.cuh"
#include "attn_mma.cuh"
#include "../../include/rat_b200.h"

namespace rat {

int precision_mode();

struct AttnBwdArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo;
    float* partials;                 // [gridDim.x][psize]
    long long nseq;
    SeqGeom g;
    int D, H, I;
    float scale, alpha;
    AttnPlan p;
};

// LayerNorm backward over the R rows of a tile + deterministic accumulation of per-column sums.
//   g = grad wrt LN output (smem, ld) ; dx[gr] = base[gr] + rstd*(g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
//   acc3[0][d] += sum_r g*xhat (dgamma), acc3[1][d] += sum_r g (dbeta), acc3[2][d] += sum_r extra[r][d] (optional)
__device__ __forceinline__ void ln_bwd_rows(const float* __restrict__ x, const float* __restrict__ base,
                                            float* __restrict__ dx, const long long* __restrict__ rowidx, int R, int D, int Dp,
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
        if (ok) { gr = rowidx[r]; mean = stats[2 * r]; rstd = stats[2 * r + 1]; }
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

template <int DH, bool MMA>
__global__ void __launch_bounds__(ENC_THREADS, 1) k_attn_bwd(AttnBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const AttnPlan& p = a.p;
    const int S = a.g.S, D = a.D, Dl = p.Dl, C3l = p.C3l, Cql = p.Cql, Il = p.Il, Dp8 = p.Dp8;
    const int Rmax = p.Rmax16;
    float* Wc = smem;                                   // [nchunks][C3p8][Dl]
    float* WoN = Wc + (size_t)p.nchunks * p.C3p8 * Dl;  // [Dp8][Il]
    float* g3 = WoN + (size_t)Dp8 * Il;                 // [3][Dp8] dgamma, dbeta, dbo
    float* scratch = g3 + 3 * Dp8;                      // [groups][3][Dp8]
    float* stats = scratch + (size_t)(ENC_THREADS / p.lg) * 3 * Dp8;   // [Rmax][2]
    float* lse = stats + 2 * Rmax;                      // [Rmax][hc]
    float* delta = lse + (size_t)Rmax * p.hc;           // [Rmax][hc]
    float* as = delta + (size_t)Rmax * p.hc;            // [Rmax][Dl] LN(x)
    float* da = as + (size_t)Rmax * Dl;                 // [Rmax][Dl] grad wrt LN output
    float* dys = da + (size_t)Rmax * Dl;                // [Rmax][Dl] alpha*dout
    float* qkv = dys + (size_t)Rmax * Dl;               // [Rmax][C3l]
    float* dqkv = qkv + (size_t)Rmax * C3l;             // [Rmax][C3l]
    float* os = dqkv + (size_t)Rmax * C3l;              // [Rmax][Cql]
    float* dos = os + (size_t)Rmax * Cql;               // [Rmax][Cql]
    float* wscr = dos + (size_t)Rmax * Cql;             // [warps][32] per-warp softmax statistics (mma attention core)
    long long* rowidx = reinterpret_cast<long long*>(wscr + (ENC_THREADS / 32) * 32);   // [Rmax]
    // CTA-private gradient record in global memory: [gW nchunks*C3p8*Dl | gWo Dp8*Il | small 3*Dp8]
    float* rec = a.partials + (size_t)blockIdx.x * p.psize;
    float* gW = rec;
    float* gWo = gW + (size_t)p.nchunks * p.C3p8 * Dl;
    stage_qkv_chunks(a.Wq, a.Wk, a.Wv, D, p, Wc);
    stage_padded(a.Wo, D, a.I, Dp8, Il, WoN);
    zero_floats(g3, 3 * Dp8);
    zero_floats(as, (size_t)Rmax * (3 * Dl + 2 * C3l + 2 * Cql));
    for (int i = threadIdx.x; i < p.psize; i += blockDim.x) rec[i] = 0.f;
    const long long ntiles = (a.nseq + p.SPT - 1) / p.SPT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * p.SPT;
        const int nseq_t = (int)min((long long)p.SPT, a.nseq - s0);
        const int R = nseq_t * S, R16 = pad16(R), R8 = pad8(R);
        __syncthreads();
        fill_rowidx(rowidx, a.g, s0, R);
        __syncthreads();
        ln_rows_to_smem(a.x, rowidx, R, D, Dp8, a.ln_w, a.ln_b, as, Dl, p.lg, stats);
        zero_floats(da, (size_t)R * Dl);
        {
            const int lg = p.lg, groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
            for (int r = gi; r < R; r += groups) {
                const float* src = a.dout + rowidx[r] * D;
                for (int d = li; d < Dp8; d += lg) dys[(size_t)r * Dl + d] = d < D ? a.alpha * src[d] : 0.f;
            }
        }
        zero_rows(as, Dl, R, R16);
        zero_rows(da, Dl, R, R16);
        zero_rows(dys, Dl, R, R16);
        zero_rows(os, Cql, R, R16);
        zero_rows(dqkv, C3l, R, R16);
        for (int ch = 0; ch < p.nchunks; ++ch) {
            const float* W = Wc + (size_t)ch * p.C3p8 * Dl;
            const int col0 = ch * p.Cq;
            __syncthreads();
            // (1) qkv[r][c] = sum_d as[r][d] W[c][d] ; (3) dos[r][c] = sum_d dys[r][d] Wo[d][col0+c]
            tc_gemm<MMA, 4>(as, Dl, 1, W, 1, Dl, qkv, C3l, R, p.C3p8, Dp8, false, EpiNone2());
            tc_gemm<MMA, 3>(dys, Dl, 1, WoN + col0, Il, 1, dos, Cql, R, p.Cq8, Dp8, false, EpiNone2());
            __syncthreads();
            const bool tc_attn = MMA && S <= 16;
            // (2) attention forward recompute -> os (+ lse for the SIMT backward core)
            if (tc_attn) attn_fwd_mma<DH>(qkv, C3l, p.Cq, os, Cql, nullptr, nseq_t, S, p.hc, a.scale);
            else attn_core<DH>(qkv, C3l, p.Cq, os, Cql, lse, nseq_t, S, p.hc, p.lpt, a.scale);
            __syncthreads();
            // delta[r][hl] = do . o   (SIMT core only; the tensor-core core derives it from P and dP)
            if (!tc_attn) for (int i = threadIdx.x; i < R * p.hc; i += blockDim.x) {
                const int r = i / p.hc, hl = i - r * p.hc;
                const float* o = os + (size_t)r * Cql + hl * DH;
                const float* dd = dos + (size_t)r * Cql + hl * DH;
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) s = fmaf(o[d], dd[d], s);
                delta[i] = s;
            }
            // (4) gWo[d][col0+c] += sum_r dys[r][d] os[r][c]
            tc_gemm<MMA, 3>(dys, 1, Dl, os, Cql, 1, gWo + col0, Il, Dp8, p.Cq8, R8, true, EpiNone2());
            __syncthreads();
            // (5) attention backward -> dqkv
            if (tc_attn) attn_bwd_mma<DH>(qkv, dqkv, C3l, p.Cq, dos, Cql, nseq_t, S, p.hc, a.scale, wscr);
            else attn_bwd_core<DH>(qkv, dqkv, C3l, p.Cq, dos, Cql, lse, delta, nseq_t, S, p.hc, p.lpt, a.scale);
            __syncthreads();
            // (6) gW[ch][c][d] += sum_r dqkv[r][c] as[r][d] ; (7) da[r][d] += sum_c dqkv[r][c] W[c][d]
            tc_gemm<MMA, 3>(dqkv, 1, C3l, as, Dl, 1, gW + (size_t)ch * p.C3p8 * Dl, Dl, p.C3p8, Dp8, R8, true, EpiNone2());
            tc_gemm<MMA, 3>(dqkv, C3l, 1, W, Dl, 1, da, Dl, R, Dp8, p.C3p8, true, EpiNone2());
        }
        __syncthreads();
        ln_bwd_rows(a.x, a.base, a.dx, rowidx, R, D, Dp8, a.ln_w, da, Dl, stats, dys, Dl, p.lg, scratch, g3);
    }
    __syncthreads();
    float* small = gWo + (size_t)Dp8 * Il;
    for (int i = threadIdx.x; i < 3 * Dp8; i += blockDim.x) small[i] = g3[i];     // dgamma | dbeta | dbo
}

// out[...] (+)= sum over CTA records, fixed order.  One thread per parameter-gradient element.
struct AttnReduceArgs {
    const float* partials; int nparts;
    float* dWq; float* dWk; float* dWv; float* dWo; float* dbo; float* dln_w; float* dln_b;
    int accumulate_wq, D, I;
    AttnPlan p;
};
__global__ void k_reduce_attn(AttnReduceArgs a) {
    const AttnPlan& p = a.p;
    const int D = a.D, I = a.I;
    const int total = 4 * I * D + 3 * D;
    const size_t off_wo = (size_t)p.nchunks * p.C3p8 * p.Dl, off_small = off_wo + (size_t)p.Dp8 * p.Il;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        size_t src;
        float* dst;
        bool acc = false;
        if (i < 3 * I * D) {
            const int which = i / (I * D), rem = i - which * (I * D);
            const int row = rem / D, d = rem - row * D;
            const int ch = row / p.Cq, c = row - ch * p.Cq;
            src = ((size_t)ch * p.C3p8 + which * p.Cq + c) * p.Dl + d;
            dst = (which == 0 ? a.dWq : which == 1 ? a.dWk : a.dWv);
            acc = which == 0 && a.accumulate_wq;
            if (dst) dst += rem;
        } else if (i < 4 * I * D) {
            const int rem = i - 3 * I * D;
            const int d = rem / I, col = rem - d * I;
            src = off_wo + (size_t)d * p.Il + col;
            dst = a.dWo ? a.dWo + rem : nullptr;
        } else {
            const int rem = i - 4 * I * D;
            const int which = rem / D, d = rem - which * D;        // 0: dbo, 1: dgamma, 2: dbeta
            src = off_small + (size_t)(which == 0 ? 2 : which == 1 ? 0 : 1) * p.Dp8 + d;
            float* base = which == 0 ? a.dbo : which == 1 ? a.dln_w : a.dln_b;
            dst = base ? base + d : nullptr;
        }
        if (!dst) continue;
        float s = 0.f;
        for (int c = 0; c < a.nparts; ++c) s += a.partials[(size_t)c * p.psize + src];
        *dst = acc ? *dst + s : s;
    }
}

// ------------------------------------------------------------------------------------------------------------
struct FFBwdArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* W1; const float* b1; const float* W2;
    float* partials;
    long long rows;
    int D, M;
    FFPlan p;
};

struct EpiBias2 {
    const float* b;
    __device__ __forceinline__ void operator()(int, int c, float& v0, float& v1) const { v0 += b[c]; v1 += b[c + 1]; }
};
// v = dh ; hs holds the pre-activation: hs <- gelu(pre), v <- dh * gelu'(pre)
struct EpiGeluBwd2 {
    float* hs; int ld;
    __device__ __forceinline__ void operator()(int r, int c, float& v0, float& v1) const {
        float2* hp = reinterpret_cast<float2*>(hs + (size_t)r * ld + c);
        const float2 pre = *hp;
        v0 *= gelu_erf_grad(pre.x);
        v1 *= gelu_erf_grad(pre.y);
        *hp = make_float2(gelu_erf(pre.x), gelu_erf(pre.y));
    }
};

template <bool MMA>
__global__ void __launch_bounds__(ENC_THREADS, 1) k_ff_bwd(FFBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const FFPlan& p = a.p;
    const int D = a.D, M = a.M, Dl = p.Dl, Ml = p.Ml, Dp8 = p.Dp8, Mp8 = p.Mp8;
    float* W1n = smem;                                  // [Mp8][Dl]  natural W1 [M,D]
    float* W2n = W1n + (size_t)Mp8 * Dl;                // [Dp8][Ml]  natural W2 [D,M]
    float* b1s = W2n + (size_t)Dp8 * Ml;                // [Mp8]
    float* gb1 = b1s + Mp8;                             // [Mp8]
    float* g3 = gb1 + Mp8;                              // [3][Dp8]  dgamma, dbeta, db2
    float* scratch = g3 + 3 * Dp8;                      // max(groups*3*Dp8, ENC_THREADS)
    float* stats = scratch + max((ENC_THREADS / p.lg) * 3 * Dp8, ENC_THREADS);   // [RPT][2]
    float* xs = stats + 2 * p.RPT;                      // [RPT][Dl]  FF input (LN(x) or x)
    float* dys = xs + (size_t)p.RPT * Dl;               // [RPT][Dl]  dout
    float* ys = dys + (size_t)p.RPT * Dl;               // [RPT][Dl]  grad wrt FF input
    float* hs = ys + (size_t)p.RPT * Dl;                // [RPT][Ml]  pre -> h
    float* dhs = hs + (size_t)p.RPT * Ml;               // [RPT][Ml]  dh -> dpre
    long long* rowidx = reinterpret_cast<long long*>(dhs + (size_t)p.RPT * Ml);   // [RPT]
    // CTA-private gradient record: [gW1 Mp8*Dl | gW2 Dp8*Ml | gb1 Mp8 | small 3*Dp8]
    float* rec = a.partials + (size_t)blockIdx.x * p.psize;
    float* gW1 = rec;
    float* gW2 = gW1 + (size_t)Mp8 * Dl;
    stage_padded(a.W1, M, D, Mp8, Dl, W1n);
    stage_padded(a.W2, D, M, Dp8, Ml, W2n);
    for (int i = threadIdx.x; i < Mp8; i += blockDim.x) { b1s[i] = i < M ? a.b1[i] : 0.f; gb1[i] = 0.f; }
    zero_floats(g3, 3 * Dp8);
    zero_floats(xs, (size_t)p.RPT * (3 * Dl + 2 * Ml));
    for (int i = threadIdx.x; i < p.psize; i += blockDim.x) rec[i] = 0.f;
    const long long ntiles = (a.rows + p.RPT - 1) / p.RPT;
    SeqGeom flat{1, 0, 1, 1};
    const int lg = p.lg, groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * p.RPT;
        const int R = (int)min((long long)p.RPT, a.rows - r0);
        const int R16 = pad16(R), R8 = pad8(R);
        __syncthreads();
        fill_rowidx(rowidx, flat, r0, R);
        __syncthreads();
        if (a.ln_w) {
            ln_rows_to_smem(a.x, rowidx, R, D, Dp8, a.ln_w, a.ln_b, xs, Dl, lg, stats);
        } else {
            for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
                const int r = i / D, d = i - r * D;
                xs[(size_t)r * Dl + d] = a.x[(r0 + r) * D + d];
            }
        }
        for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
            const int r = i / D, d = i - r * D;
            dys[(size_t)r * Dl + d] = a.dout[(r0 + r) * D + d];
        }
        zero_rows(xs, Dl, R, R16);
        zero_rows(dys, Dl, R, R16);
        __syncthreads();
        // pre[r][m] = sum_d x[r][d] W1[m][d] + b1[m]
        tc_gemm<MMA, 4>(xs, Dl, 1, W1n, 1, Dl, hs, Ml, R, Mp8, Dp8, false, EpiBias2{b1s});
        __syncthreads();
        // dpre[r][m] = (sum_d dy[r][d] W2[d][m]) * gelu'(pre) ; hs <- gelu(pre)
        tc_gemm<MMA, 4>(dys, Dl, 1, W2n, Ml, 1, dhs, Ml, R, Mp8, Dp8, false, EpiGeluBwd2{hs, Ml});
        zero_rows(hs, Ml, R, R16);          // pad rows were never touched by the epilogue; keep them zero
        zero_rows(dhs, Ml, R, R16);
        __syncthreads();
        // dW2[d][m] += sum_r dy[r][d] h[r][m] ; dW1[m][d] += sum_r dpre[r][m] x[r][d] ; dxa[r][d] = sum_m dpre[r][m] W1[m][d]
        tc_gemm<MMA, 4>(dys, 1, Dl, hs, Ml, 1, gW2, Ml, Dp8, Mp8, R8, true, EpiNone2());
        tc_gemm<MMA, 3>(dhs, 1, Ml, xs, Dl, 1, gW1, Dl, Mp8, Dp8, R8, true, EpiNone2());
        tc_gemm<MMA, 3>(dhs, Ml, 1, W1n, Dl, 1, ys, Dl, R, Dp8, Mp8, false, EpiNone2());
        tile_colsum_acc2(dhs, Ml, gb1, R, Mp8, scratch);              // db1 (two barriers inside)
        if (a.ln_w) {
            ln_bwd_rows(a.x, a.base, a.dx, rowidx, R, D, Dp8, a.ln_w, ys, Dl, stats, dys, Dl, lg, scratch, g3);
        } else {
            // dx = base + dxa ; db2 += colsum(dy)  (deterministic per-group partials)
            float pe[4] = {0.f, 0.f, 0.f, 0.f};
            for (int rr = 0; rr < R; rr += groups) {
                const int r = rr + gi;
                if (r < R) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int d = li + k * lg;
                        if (d < D) {
                            float v = ys[(size_t)r * Dl + d];
                            if (a.base) v += a.base[(r0 + r) * D + d];
                            a.dx[(r0 + r) * D + d] = v;
                            pe[k] += dys[(size_t)r * Dl + d];
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int d = li + k * lg; if (d < D) scratch[gi * Dp8 + d] = pe[k]; }
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += blockDim.x) {
                float s = 0.f;
                for (int q = 0; q < groups; ++q) s += scratch[q * Dp8 + d];
                g3[2 * Dp8 + d] += s;
            }
        }
    }
    __syncthreads();
    float* sm_out = gW2 + (size_t)Dp8 * Ml;
    for (int i = threadIdx.x; i < Mp8; i += blockDim.x) sm_out[i] = gb1[i];
    for (int i = threadIdx.x; i < 3 * Dp8; i += blockDim.x) sm_out[Mp8 + i] = g3[i];
}

struct FFReduceArgs {
    const float* partials; int nparts;
    float* dW1; float* db1; float* dW2; float* db2; float* dln_w; float* dln_b;
    int D, M;
    FFPlan p;
};
__global__ void k_reduce_ff(FFReduceArgs a) {
    const FFPlan& p = a.p;
    const int D = a.D, M = a.M;
    const int total = 2 * M * D + M + 3 * D;
    const size_t off_w2 = (size_t)p.Mp8 * p.Dl, off_b1 = off_w2 + (size_t)p.Dp8 * p.Ml, off_small = off_b1 + p.Mp8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        size_t src;
        float* dst;
        if (i < M * D) { const int m = i / D, d = i - m * D; src = (size_t)m * p.Dl + d; dst = a.dW1 ? a.dW1 + i : nullptr; }
        else if (i < 2 * M * D) { const int rem = i - M * D; const int d = rem / M, m = rem - d * M;
                                  src = off_w2 + (size_t)d * p.Ml + m; dst = a.dW2 ? a.dW2 + rem : nullptr; }
        else if (i < 2 * M * D + M) { const int m = i - 2 * M * D; src = off_b1 + m; dst = a.db1 ? a.db1 + m : nullptr; }
        else {
            const int rem = i - 2 * M * D - M;
            const int which = rem / D, d = rem - which * D;       // 0: db2, 1: dgamma, 2: dbeta
            src = off_small + (size_t)(which == 0 ? 2 : which == 1 ? 0 : 1) * p.Dp8 + d;
            float* base = which == 0 ? a.db2 : which == 1 ? a.dln_w : a.dln_b;
            dst = base ? base + d : nullptr;
        }
        if (!dst) continue;
        float s = 0.f;
        for (int c = 0; c < a.nparts; ++c) s += a.partials[(size_t)c * p.psize + src];
        *dst = s;
    }
}

// generic: out[i] = sum_c partials[c][off + i]  (final LayerNorm gradients)
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
int plan_attn_bwd(int S, int D, int H, int dh, AttnPlan* out) {
    const size_t bud = (size_t)(max_smem_optin() - 2048) / 4;
    const int cap_rows = 192;
    AttnPlan best{};
    int bestR = 0;
    for (int hc = H; hc >= 1; --hc) {
        if (H % hc) continue;
        AttnPlan c{};
        fill_attn_plan(S, D, H, dh, hc, &c);
        const size_t fixed = (size_t)c.nchunks * c.C3p8 * c.Dl + (size_t)c.Dp8 * c.Il + 3 * c.Dp8 +
                             (size_t)(ENC_THREADS / c.lg) * 3 * c.Dp8 + ENC_THREADS;
        const size_t per_row = 3 * (size_t)c.Dl + 2 * c.C3l + 2 * c.Cql + 2 + 2 * hc + 2;
        if (fixed + per_row * pad16(S) > bud) continue;
        int spt = (int)min((size_t)max(1, cap_rows / S), (bud - fixed) / (per_row * S));
        while (spt > 1 && fixed + per_row * pad16(spt * S) > bud) --spt;
        if (spt < 1) continue;
        const int R = spt * S;
        if (R > bestR) {
            bestR = R; best = c; best.SPT = spt; best.Rmax16 = pad16(R);
            best.smem_bytes = (fixed + per_row * pad16(R)) * 4;
        }
        if (R >= min(96, max(1, cap_rows / S) * S)) break;
    }
    if (!bestR) return RAT_ESMEM;
    best.psize = (int)((size_t)best.nchunks * best.C3p8 * best.Dl + (size_t)best.Dp8 * best.Il + 3 * best.Dp8);
    best.psize = (best.psize + 3) & ~3;
    *out = best;
    return RAT_OK;
}

int plan_ff_bwd(int D, int M, FFPlan* out) {
    FFPlan p{};
    fill_ff_plan(D, M, &p);
    const size_t bud = (size_t)(max_smem_optin() - 2048) / 4;
    const size_t fixed = (size_t)p.Mp8 * p.Dl + (size_t)p.Dp8 * p.Ml + 2 * p.Mp8 + 3 * p.Dp8 +
                         (size_t)max((ENC_THREADS / p.lg) * 3 * p.Dp8, ENC_THREADS);
    const size_t per_row = 3 * (size_t)p.Dl + 2 * p.Ml + 2 + 2;
    int rpt = 192;
    while (rpt >= 16 && fixed + per_row * rpt > bud) rpt -= 16;
    if (rpt < 16) return RAT_ESMEM;
    p.RPT = rpt;
    p.smem_bytes = (fixed + per_row * rpt) * 4;
    p.psize = (int)((size_t)p.Mp8 * p.Dl + (size_t)p.Dp8 * p.Ml + p.Mp8 + 3 * p.Dp8);
    p.psize = (p.psize + 3) & ~3;
    *out = p;
    return RAT_OK;
}

static int bwd_grid(long long ntiles) { return (int)min(ntiles, (long long)num_sms()); }

}  // namespace rat

using namespace rat;

size_t ff_bwd_tc_workspace_bytes(long long rows, int D, int M);
size_t attn_bwd_tc_workspace_bytes(int B, int T, int N, int D, int heads, int dh, int mode);
int attn_bwd_tc_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                         const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                         float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq,
                         int B, int T, int N, int D, int heads, int dh, float scale, float alpha, int mode,
                         const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, float out_drop_p,
                         unsigned long long seed, unsigned int rng_stream, cudaStream_t st);
int attn_bwd_rr_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                         const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                         float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq,
                         int B, int T, int N, int D, int heads, int dh, float scale, float alpha, int mode,
                         const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, float out_drop_p,
                         unsigned long long seed, unsigned int rng_stream, cudaStream_t st);
size_t attn_bwd_rr_workspace_bytes(int B, int T, int N, int D, int heads, int dh, int mode);
bool rr_enabled();
int ff_bwd_rr_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* W1, const float* b1,
                       const float* W2, float* dW1, float* db1, float* dW2, float* db2, long long rows, int D, int M,
                       const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, cudaStream_t st);
size_t ff_bwd_rr_workspace_bytes(long long rows, int D, int M);
int ff_bwd_tc_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                       const float* W1, const float* b1, const float* W2, float* dW1, float* db1, float* dW2, float* db2,
                       long long rows, int D, int M, const float* dout_amax, float* dx_amax, float* workspace,
                       size_t workspace_bytes, cudaStream_t st);

extern "C" size_t rat_attn_bwd_workspace_bytes(int B, int T, int N, int D, int heads, int dim_head, int mode) {
    AttnPlan p{};
    const int S = mode == 0 ? N : T;
    if (plan_attn_bwd(S, D, heads, dim_head, &p) != RAT_OK) return 0;
    const long long nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    const long long ntiles = (nseq + p.SPT - 1) / p.SPT;
    return std::max({(size_t)bwd_grid(ntiles) * p.psize * sizeof(float), attn_bwd_tc_workspace_bytes(B, T, N, D, heads, dim_head, mode),
                     attn_bwd_rr_workspace_bytes(B, T, N, D, heads, dim_head, mode)});
}

template <int DH, bool MMA>
static int launch_attn_bwd(const AttnBwdArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_bwd<DH, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_bwd)");
        attr_set = true;
    }
    k_attn_bwd<DH, MMA><<<grid, ENC_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_bwd");
    return RAT_OK;
}
template <int DH>
static int launch_attn_bwd_p(const AttnBwdArgs& a, int grid, cudaStream_t st) {
    return precision_mode() ? launch_attn_bwd<DH, true>(a, grid, st) : launch_attn_bwd<DH, false>(a, grid, st);
}

extern "C" int rat_attn_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                            const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                            float* dWq, float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b,
                            int accumulate_wq, int B, int T, int N, int D, int heads, int dim_head, float scale,
                            float alpha, int mode, const float* dout_amax, float* dx_amax, float* workspace,
                            size_t workspace_bytes, void* stream) {
    return rat_attn_bwd_dropout(x, dout, base, dx, ln_w, ln_b, Wq, Wk, Wv, Wo, dWq, dWk, dWv, dWo, dbo, dln_w, dln_b, accumulate_wq,
                                B, T, N, D, heads, dim_head, scale, alpha, mode, dout_amax, dx_amax, workspace, workspace_bytes,
                                0.0f, 0ull, 0u, stream);
}

extern "C" int rat_dropout_bwd(float* grad, long long n, float p, unsigned long long seed, unsigned int rng_stream, void* stream);

extern "C" int rat_attn_bwd_dropout(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                                    const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                                    float* dWq, float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b,
                                    int accumulate_wq, int B, int T, int N, int D, int heads, int dim_head, float scale,
                                    float alpha, int mode, const float* dout_amax, float* dx_amax, float* workspace,
                                    size_t workspace_bytes, float out_drop_p, unsigned long long seed, unsigned int rng_stream,
                                    void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && N > 0 && D > 0 && heads > 0, "rat_attn_bwd: bad shape");
    RAT_REQUIRE(D <= 128, "rat_attn_bwd: D=%d > 128 not supported", D);
    RAT_REQUIRE(out_drop_p >= 0.f && out_drop_p < 1.f, "rat_attn_bwd_dropout: dropout p=%f", out_drop_p);
    reduce_ws_acquire((cudaStream_t)stream, workspace);     // deferred record reductions (rat_set_reduce_stream)
    if (precision_mode() == 2 && rr_enabled() && !getenv("RAT_RR_BWD_OFF")) {
        const int rc3 = attn_bwd_rr_dispatch(x, dout, base, dx, ln_w, ln_b, Wq, Wk, Wv, Wo, dWq, dWk, dWv, dWo, dbo, dln_w,
                                             dln_b, accumulate_wq, B, T, N, D, heads, dim_head, scale, alpha, mode, dout_amax, dx_amax,
                                             workspace, workspace_bytes, out_drop_p, seed, rng_stream, (cudaStream_t)stream);
        if (rc3 <= 0) return rc3;
    }
    if (precision_mode() == 2) {
        const int rc2 = attn_bwd_tc_dispatch(x, dout, base, dx, ln_w, ln_b, Wq, Wk, Wv, Wo, dWq, dWk, dWv, dWo, dbo, dln_w,
                                             dln_b, accumulate_wq, B, T, N, D, heads, dim_head, scale, alpha, mode, dout_amax, dx_amax,
                                             workspace, workspace_bytes, out_drop_p, seed, rng_stream, (cudaStream_t)stream);
        if (rc2 <= 0) return rc2;
    }
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
        case 4: rc = launch_attn_bwd_p<4>(a, grid, st); break;
        case 8: rc = launch_attn_bwd_p<8>(a, grid, st); break;
        case 10: rc = launch_attn_bwd_p<10>(a, grid, st); break;
        case 16: rc = launch_attn_bwd_p<16>(a, grid, st); break;
        case 20: rc = launch_attn_bwd_p<20>(a, grid, st); break;
        case 32: rc = launch_attn_bwd_p<32>(a, grid, st); break;
        default: set_error("rat_attn_bwd: dim_head=%d not instantiated", dim_head); return RAT_EINVAL;
    }
    if (rc != RAT_OK) return rc;
    AttnReduceArgs r{workspace, grid, dWq, dWk, dWv, dWo, dbo, dln_w, dln_b, accumulate_wq, D, a.I, a.p};
    const int total = 4 * a.I * D + 3 * D;
    k_reduce_attn<<<max(1, min(ceil_div(total, 256), 1024)), 256, 0, st>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_attn");
    if (out_drop_p > 0.f) {  // these kernels do not fuse the output mask: same mask as a separate pass
        rc = rat_dropout_bwd(dx, (long long)B * T * N * D, out_drop_p, seed, rng_stream, stream);
        if (rc != RAT_OK) return rc;
    }
    if (dx_amax)            // fp32 / tf32 kernels do not track the maximum themselves
        return rat_absmax(dx, (long long)B * T * N, D, D, dx_amax, stream);
    return RAT_OK;
}

extern "C" size_t rat_ff_bwd_workspace_bytes(long long rows, int D, int M) {
    FFPlan p{};
    if (plan_ff_bwd(D, M, &p) != RAT_OK) return 0;
    const long long ntiles = (rows + p.RPT - 1) / p.RPT;
    return std::max({(size_t)bwd_grid(ntiles) * p.psize * sizeof(float), ff_bwd_tc_workspace_bytes(rows, D, M),
                     ff_bwd_rr_workspace_bytes(rows, D, M)});
}

template <bool MMA>
static int launch_ff_bwd(const FFBwdArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_bwd<MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_bwd)");
        attr_set = true;
    }
    k_ff_bwd<MMA><<<grid, ENC_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_ff_bwd");
    return RAT_OK;
}

extern "C" int rat_ff_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                          const float* ln_b, const float* W1, const float* b1, const float* W2, float* dW1, float* db1,
                          float* dW2, float* db2, float* dln_w, float* dln_b, long long rows, int D, int M,
                          const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && M > 0, "rat_ff_bwd: bad shape");
    RAT_REQUIRE(D <= 128 && pad8(M) <= ENC_THREADS, "rat_ff_bwd: D=%d (<=128) M=%d (<=%d) not supported", D, M, ENC_THREADS);
    reduce_ws_acquire((cudaStream_t)stream, workspace);
    if (precision_mode() == 2 && rr_enabled() && ln_w == nullptr && !getenv("RAT_RR_FF_OFF")) {
        const int rc3 = ff_bwd_rr_dispatch(x, dout, base, dx, W1, b1, W2, dW1, db1, dW2, db2, rows, D, M, dout_amax, dx_amax, workspace,
                                           workspace_bytes, (cudaStream_t)stream);
        if (rc3 <= 0) return rc3;
    }
    if (precision_mode() == 2) {
        const int rc2 = ff_bwd_tc_dispatch(x, dout, base, dx, ln_w, W1, b1, W2, dW1, db1, dW2, db2, rows, D, M, dout_amax, dx_amax,
                                           workspace, workspace_bytes, (cudaStream_t)stream);
        if (rc2 <= 0) return rc2;
    }
    FFBwdArgs a{};
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.W2 = W2;
    a.rows = rows; a.D = D; a.M = M;
    int rc = plan_ff_bwd(D, M, &a.p);
    if (rc != RAT_OK) { set_error("rat_ff_bwd: D=%d M=%d does not fit in shared memory", D, M); return rc; }
    const long long ntiles = (rows + a.p.RPT - 1) / a.p.RPT;
    const int grid = bwd_grid(ntiles);
    RAT_REQUIRE(workspace && workspace_bytes >= (size_t)grid * a.p.psize * sizeof(float), "rat_ff_bwd: workspace too small");
    a.partials = workspace;
    cudaStream_t st = (cudaStream_t)stream;
    rc = precision_mode() ? launch_ff_bwd<true>(a, grid, st) : launch_ff_bwd<false>(a, grid, st);
    if (rc != RAT_OK) return rc;
    FFReduceArgs r{workspace, grid, dW1, db1, dW2, db2, dln_w, dln_b, D, M, a.p};
    const int total = 2 * M * D + M + 3 * D;
    k_reduce_ff<<<max(1, min(ceil_div(total, 256), 1024)), 256, 0, st>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_ff");
    if (dx_amax) return rat_absmax(dx, rows, D, D, dx_amax, stream);
    return RAT_OK;
}

extern "C" size_t rat_layernorm_bwd_workspace_bytes(long long rows, int D) {
    (void)rows;
    return (size_t)num_sms() * 4 * 2 * D * sizeof(float);
}

extern "C" int rat_layernorm_bwd(const float* x, const float* dout, float* dx, const float* w, float* dw, float* db,
                                 long long rows, int D, float* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && D <= 128, "rat_layernorm_bwd: bad shape");
    const int lg = min(32, next_pow2(D));
    const int groups = 256 / lg;
    const int grid = (int)min((rows + groups - 1) / groups, (long long)num_sms() * 4);
    RAT_REQUIRE(workspace && workspace_bytes >= (size_t)grid * 2 * D * sizeof(float), "rat_layernorm_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    reduce_ws_acquire(st, workspace);
    k_ln_bwd<<<grid, 256, (size_t)groups * 2 * D * sizeof(float), st>>>(x, dout, dx, w, rows, D, lg, workspace);
    RAT_CHECK_LAUNCH("k_ln_bwd");
    ReduceArgs r{};
    r.partials = workspace; r.nparts = grid; r.psize = 2 * D; r.nseg = 2;
    r.seg[0] = ReduceSeg{dw, 0, D, 0};
    r.seg[1] = ReduceSeg{db, D, D, 0};
    k_reduce_partials<<<1, 256, 0, st>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_partials");
    return RAT_OK;
}
