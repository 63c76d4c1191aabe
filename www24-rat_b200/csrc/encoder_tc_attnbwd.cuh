// K5 (Blackwell path): attention backward kernel + host planning (included once per head width, see encoder_tc_attnbwd_dh*.cu).
#pragma once
#include "encoder_tc.cuh"

namespace rat {

// ------------------------------------------------------------------------------------------------ Attention backward
// Backward of  out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo ):  dx = base + dLN(...), and all parameter gradients.
// One team of 16 warps per CTA (the weight-gradient accumulators of all heads live in registers, 5-6 16x16 tiles per
// warp).  Per 128-row tile and per chunk of hc heads:
//   tcgen05: q|k|v = LN(x) Wqkv^T (recompute, padded heads)  and  dO = (alpha dout) Wo_chunk (padded heads)  -> TMEM
//   TMEM -> bf16 tiles ; one warp per (sequence, head): recompute P, O ; dP = dO V^T ; dS = P o (dP - rowsum(P o dP)) ;
//     dQ = scale dS K ; dK = dS^T Q ; dV = P^T dO  (mma.sync bf16, fragment transposes by movmatrix) -> COMPACT bf16
//     dq|dk|dv tile (hc*dh columns per part, no head padding) and compact O tile
//   tcgen05: dA[128 x Kp] += dqkv_c . Wqkv_c  (accumulated over chunks in TMEM)
//   mma.sync weight-gradient jobs (token reduction): gWqkv_c += dqkv_c^T LN(x) ; gWo_c^T += O_c^T (alpha dout)
// tile end: LayerNorm backward from dA (fp32), dx written once; dgamma/dbeta/dbo as ones-row jobs on bf16 tiles.
struct AttnBwdTcArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo;
    float* partials;         // [gridDim.x][psize]
    const float* dout_amax;  // device: max|dout| (nullptr: no gradient scaling)
    float* dx_amax;          // device: receives max|dx| (nullptr: not wanted)
    long long nseq;
    SeqGeom g;
    int D, H, I;
    float scale, alpha;
    int Kp, hc, nchunks;
    int NCq;                 // 3 * hc * DHP   padded q|k|v columns per chunk (N of the recompute GEMM)
    int NCc;                 // pad16(3 * hc * dh) compact dq|dk|dv columns per chunk (K of the dA GEMM)
    int NDo;                 // hc * DHP        padded dO columns per chunk
    int Cc;                  // pad16(hc * dh)  compact O columns per chunk
    int SPT, njobs, psize, smem_bytes;
    // optional mask on dx: the backward of nn.Dropout(emb_dropout) (RAT_m2.py:135) fused into the LAST backward kernel of the
    // encoder, with the mask of rat_gather_fwd (same seed / stream / element index), so the segment reduce reads a ready gradient
    float out_drop_p; unsigned long long seed; unsigned int rng_stream; const unsigned int* rng_step;
};

__device__ __forceinline__ uint32_t movm_t(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

template <int DH>
__device__ __forceinline__ void attn_bwd_core_bf16(const unsigned char* __restrict__ QKVt, const unsigned char* __restrict__ dOt,
                                                   unsigned char* __restrict__ dQc, unsigned char* __restrict__ Ot, int hc,
                                                   int nseq_t, int S, float scale, int warp, int nwarps, int lane,
                                                   const CoreLane& cl, const uint32_t* __restrict__ coltab) {
    constexpr int DHP = (DH + 15) / 16 * 16, KS = DHP / 16, ND = (DH + 7) / 8;
    constexpr uint32_t HEAD = (DHP / 8) * tc5::TILE_CHUNK;
    const int t = lane & 3;
    const int ntasks = (cl.packed ? (nseq_t + 1) >> 1 : nseq_t) * hc;
    const uint32_t qkv_s = tc5::smem_u32(QKVt), do_s = tc5::smem_u32(dOt);
    const uint32_t part = (uint32_t)hc * HEAD;
    const float inv_log2e = 0.6931471805599453f;
    int sp = warp / hc, hl = warp - sp * hc;
    const int dsp = nwarps / hc, dhl = nwarps - dsp * hc;
    for (int task = warp; task < ntasks; task += nwarps) {
        const int seq0 = cl.packed ? 2 * sp : sp;
        const uint32_t tb = (uint32_t)(seq0 * S) * 16u;
        const uint32_t qa = qkv_s + tb + (uint32_t)hl * HEAD;         // q of this head; k at +part, v at +2 part
        const uint32_t da = do_s + tb + (uint32_t)hl * HEAD;
        // ---- S = Q K^T (log2 domain, scale folded into Q) and dP = dO V^T
        float sc[2][4] = {}, dp[2][4] = {};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a[4], b[4], a2[4], b2[4];
            ldsm_x4(a, qa + cl.a_off + 2 * ks * tc5::TILE_CHUNK);
            ldsm_x4(b, qa + part + cl.b_off + 2 * ks * tc5::TILE_CHUNK);
            ldsm_x4(a2, da + cl.a_off + 2 * ks * tc5::TILE_CHUNK);
            ldsm_x4(b2, qa + 2 * part + cl.b_off + 2 * ks * tc5::TILE_CHUNK);
            mma_h_16x8x16(sc[0], a, b[0], b[1]);
            mma_h_16x8x16(sc[1], a, b[2], b[3]);
            mma_h_16x8x16(dp[0], a2, b2[0], b2[1]);
            mma_h_16x8x16(dp[1], a2, b2[2], b2[3]);
        }
        const bool has2 = !cl.packed || (seq0 + 1 < nseq_t);
        const bool vlo = cl.lo_rel >= 0, vhi = cl.hi_rel >= 0 && has2;
        float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] += cl.madd[nt][e];
                if (e < 2) mlo = fmaxf(mlo, sc[nt][e]); else mhi = fmaxf(mhi, sc[nt][e]);
            }
        mlo = qmax(mlo); mhi = qmax(mhi);
        if (mlo == -INFINITY) mlo = 0.f;
        if (mhi == -INFINITY) mhi = 0.f;
        float llo = 0.f, lhi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] = ex2f(sc[nt][e] - ((e < 2) ? mlo : mhi));
                if (e < 2) llo += sc[nt][e]; else lhi += sc[nt][e];
            }
        llo = qsum(llo); lhi = qsum(lhi);
        // rows that do not exist (or belong to a missing second sequence) get P = 0: they must not leak into dK / dV
        const float ilo = vlo ? rcp_fast(llo) : 0.f, ihi = vhi ? rcp_fast(lhi) : 0.f;
        float dlo = 0.f, dhi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[nt][e] *= (e < 2) ? ilo : ihi;                                    // P
                if (e < 2) dlo = fmaf(sc[nt][e], dp[nt][e], dlo); else dhi = fmaf(sc[nt][e], dp[nt][e], dhi);
            }
        dlo = qsum(dlo); dhi = qsum(dhi);                                            // delta_i = sum_j P_ij dP_ij
        uint32_t pa[4], sa[4];
        pa[0] = pack_h2(sc[0][0], sc[0][1]); pa[1] = pack_h2(sc[0][2], sc[0][3]);
        pa[2] = pack_h2(sc[1][0], sc[1][1]); pa[3] = pack_h2(sc[1][2], sc[1][3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) dp[nt][e] = sc[nt][e] * (dp[nt][e] - ((e < 2) ? dlo : dhi));   // dS
        sa[0] = pack_h2(dp[0][0], dp[0][1]); sa[1] = pack_h2(dp[0][2], dp[0][3]);
        sa[2] = pack_h2(dp[1][0], dp[1][1]); sa[3] = pack_h2(dp[1][2], dp[1][3]);
        // transposed A fragments (P^T, dS^T): transpose each 8x8 block and swap the off-diagonal blocks
        uint32_t pt[4], st[4];
        pt[0] = movm_t(pa[0]); pt[1] = movm_t(pa[2]); pt[2] = movm_t(pa[1]); pt[3] = movm_t(pa[3]);
        st[0] = movm_t(sa[0]); st[1] = movm_t(sa[2]); st[2] = movm_t(sa[1]); st[3] = movm_t(sa[3]);
        // ---- O = P V ; dQ = dS K ; dK = dS^T Q ; dV = P^T dO      (B operands: ldmatrix.trans of [row][d] tiles)
        float o[2 * KS][4] = {}, dq[2 * KS][4] = {}, dk[2 * KS][4] = {}, dv[2 * KS][4] = {};
#pragma unroll
        for (int pp = 0; pp < KS; ++pp) {
            uint32_t bv[4], bk[4], bq[4], bo[4];
            const uint32_t ko = cl.a_off + 2 * pp * tc5::TILE_CHUNK;
            ldsm_x4_t(bv, qa + 2 * part + ko);          // V
            ldsm_x4_t(bk, qa + part + ko);              // K
            ldsm_x4_t(bq, qa + ko);                     // Q (scaled)
            ldsm_x4_t(bo, da + ko);                     // dO
            mma_h_16x8x16(o[2 * pp], pa, bv[0], bv[1]);
            mma_h_16x8x16(dq[2 * pp], sa, bk[0], bk[1]);
            mma_h_16x8x16(dk[2 * pp], st, bq[0], bq[1]);
            mma_h_16x8x16(dv[2 * pp], pt, bo[0], bo[1]);
            if (2 * pp + 1 < ND) {
                mma_h_16x8x16(o[2 * pp + 1], pa, bv[2], bv[3]);
                mma_h_16x8x16(dq[2 * pp + 1], sa, bk[2], bk[3]);
                mma_h_16x8x16(dk[2 * pp + 1], st, bq[2], bq[3]);
                mma_h_16x8x16(dv[2 * pp + 1], pt, bo[2], bo[3]);
            }
        }
        // ---- compact stores: columns [q: hl*DH + d | k: hc*DH + hl*DH + d | v: 2*hc*DH + hl*DH + d], O: hl*DH + d
        const uint32_t rlo = tb + (uint32_t)(cl.lo_rel * 16), rhi = tb + (uint32_t)(cl.hi_rel * 16);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
            const int d = 8 * nd + 2 * t;
            if (d < DH) {
                // byte offset of compact column c inside a tile row: (c >> 3) * TILE_CHUNK + (c & 7) * 2, from a table
                const int cq = hl * DH + d;
                const uint32_t oq = coltab[cq], ok = coltab[cq + hc * DH], ov = coltab[cq + 2 * hc * DH];
                if (vlo) {
                    *reinterpret_cast<uint32_t*>(dQc + rlo + oq) = pack_h2(dq[nd][0] * scale, dq[nd][1] * scale);
                    *reinterpret_cast<uint32_t*>(dQc + rlo + ok) = pack_h2(dk[nd][0] * inv_log2e, dk[nd][1] * inv_log2e);
                    *reinterpret_cast<uint32_t*>(dQc + rlo + ov) = pack_h2(dv[nd][0], dv[nd][1]);
                    *reinterpret_cast<uint32_t*>(Ot + rlo + oq) = pack_h2(o[nd][0], o[nd][1]);
                }
                if (vhi) {
                    *reinterpret_cast<uint32_t*>(dQc + rhi + oq) = pack_h2(dq[nd][2] * scale, dq[nd][3] * scale);
                    *reinterpret_cast<uint32_t*>(dQc + rhi + ok) = pack_h2(dk[nd][2] * inv_log2e, dk[nd][3] * inv_log2e);
                    *reinterpret_cast<uint32_t*>(dQc + rhi + ov) = pack_h2(dv[nd][2], dv[nd][3]);
                    *reinterpret_cast<uint32_t*>(Ot + rhi + oq) = pack_h2(o[nd][2], o[nd][3]);
                }
            }
        }
        sp += dsp; hl += dhl;
        if (hl >= hc) { hl -= hc; ++sp; }
    }
}

template <int DH, int KCH, bool VEC4, int JW>
__global__ void __launch_bounds__(TC_THREADS, 1) k_attn_bwd_tc(AttnBwdTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int DHP = (DH + 15) / 16 * 16;
    constexpr int NW = TC_THREADS / 32;
    const int D = a.D, Kp = a.Kp, hc = a.hc, NCq = a.NCq, NCc = a.NCc, NDo = a.NDo, Cc = a.Cc, S = a.g.S;
    const int KC1 = Kp >> 3, KCq = NCq >> 3, KCc = NCc >> 3, KCd = NDo >> 3, KCo = Cc >> 3;
    unsigned char* Wqkv_i = smem_raw;                                        // [nchunks*NCq x Kp]  recompute (q rows scaled)
    unsigned char* WqkvT_i = Wqkv_i + (size_t)a.nchunks * NCq * Kp * 2;      // [nchunks][Kp x NCc] dA GEMM  (n = d, k = compact col)
    unsigned char* WoT_i = WqkvT_i + (size_t)a.nchunks * Kp * NCc * 2;       // [nchunks][NDo x Kp] dO GEMM  (n = padded c, k = d)
    float* stats = reinterpret_cast<float*>(WoT_i + (size_t)a.nchunks * NDo * Kp * 2);   // [128][2] mean, rstd
    float* parts = stats + 2 * TILE_M;                                       // [128][4][2] LayerNorm-backward row partials
    unsigned char* At = reinterpret_cast<unsigned char*>(parts + 8 * TILE_M);            // [128 x Kp]  LN(x)
    unsigned char* DYt = At + (size_t)TILE_M * Kp * 2;                       // [128 x Kp]  alpha * dout
    unsigned char* QKVt = DYt + (size_t)TILE_M * Kp * 2;                     // [128 x NCq] ; after the chunk loop: G1 | G2
    unsigned char* dQc = QKVt + (size_t)TILE_M * NCq * 2;                    // [128 x NCc] compact dq|dk|dv
    unsigned char* Ot = dQc + (size_t)TILE_M * NCc * 2;                      // [128 x Cc]  compact O
    unsigned char* dOt = Ot + (size_t)TILE_M * Cc * 2;                       // [128 x NDo] padded dO
    unsigned char* G1t = QKVt;                                               // [128 x Kp]  g = dA          (aliases QKVt)
    unsigned char* G2t = QKVt + (size_t)TILE_M * Kp * 2;                     // [128 x Kp]  g * xhat
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t coltab[272];                                         // compact column -> byte offset in a tile row
    for (int c = threadIdx.x; c < 272; c += blockDim.x) coltab[c] = (uint32_t)(c >> 3) * tc5::TILE_CHUNK + (uint32_t)(c & 7) * 2u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gs = tc_grad_scale(a.dout_amax), inv_gs = 1.0f / gs;
    float dx_max = 0.f;

    // ---- resident weight images
    {
        const int rows_img = a.nchunks * NCq;
        for (int i = threadIdx.x; i < rows_img * KC1; i += blockDim.x) {
            const int n = i % rows_img, kc = i / rows_img;
            const int ch = n / NCq, rem = n - ch * NCq;
            const int w = rem / (hc * DHP), rem2 = rem - w * (hc * DHP);
            const int hl = rem2 / DHP, d = rem2 - hl * DHP;
            const float* W = w == 0 ? a.Wq : w == 1 ? a.Wk : a.Wv;
            const float mul = w == 0 ? a.scale * 1.4426950408889634f : 1.0f;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = kc * 8 + k;
                v[k] = (d < DH && c < D) ? mul * __ldg(W + (size_t)((ch * hc + hl) * DH + d) * D + c) : 0.f;
            }
            sts128(Wqkv_i + (size_t)ch * NCq * Kp * 2 + tc5::kmajor_off(rem, kc, NCq), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                   pack_h2(v[6], v[7]));
        }
        const int perT = Kp * KCc;
        for (int i = threadIdx.x; i < a.nchunks * perT; i += blockDim.x) {
            const int ch = i / perT, rem = i - ch * perT;
            const int d = rem % Kp, kc = rem / Kp;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int nc = kc * 8 + k;                       // compact column: [q | k | v] x [hl][dd]
                const int w = nc / (hc * DH), rem2 = nc - w * (hc * DH);
                const float* W = w == 0 ? a.Wq : w == 1 ? a.Wk : a.Wv;
                v[k] = (w < 3 && d < D) ? __ldg(W + (size_t)(ch * hc * DH + rem2) * D + d) : 0.f;
            }
            sts128(WqkvT_i + (size_t)ch * Kp * NCc * 2 + tc5::kmajor_off(d, kc, Kp), pack_h2(v[0], v[1]),
                   pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        const int perO = NDo * KC1;
        for (int i = threadIdx.x; i < a.nchunks * perO; i += blockDim.x) {
            const int ch = i / perO, rem = i - ch * perO;
            const int n = rem % NDo, kc = rem / NDo;
            const int hl = n / DHP, dd = n - hl * DHP;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = kc * 8 + k;
                v[k] = (dd < DH && d < D) ? __ldg(a.Wo + (size_t)d * a.I + (ch * hc + hl) * DH + dd) : 0.f;
            }
            sts128(WoT_i + (size_t)ch * NDo * Kp * 2 + tc5::kmajor_off(n, kc, NDo), pack_h2(v[0], v[1]),
                   pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        // activation tiles start as zeros (pad columns of the compact tiles are never written)
        const size_t tile_bytes = (size_t)TILE_M * (2 * Kp + NCq + NCc + Cc + NDo) * 2;
        for (int i = threadIdx.x; i < (int)(tile_bytes / 16); i += blockDim.x)
            reinterpret_cast<uint4*>(At)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) { tc5::mbar_init(&mbar, 1); tc5::fence_mbar_init(); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_Q = tmem_base_s;                             // [0, NCq)
    const uint32_t tmem_O = tmem_Q + NCq;                            // [NCq, NCq + NDo)
    const uint32_t tmem_A = tmem_O + NDo;                            // [.., + Kp)
    const uint32_t idesc_q = tc5::instr_desc(TC_FMT, TILE_M, NCq);
    const uint32_t idesc_o = tc5::instr_desc(TC_FMT, TILE_M, NDo);
    const uint32_t idesc_a = tc5::instr_desc(TC_FMT, TILE_M, Kp);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int part = warp >> 2;                                      // 0..3: column share of this warp inside its lane quadrant
    const uint32_t tmem_P = tmem_A + Kp + (uint32_t)(part * JW * 8); // parking columns of this warp's accumulators
    const int row_e = (warp & 3) * 32 + lane;
    uint32_t phase = 0;
    const CoreLane cl = make_core_lane(S, lane);
    const uint32_t At_s = tc5::smem_u32(At), DYt_s = tc5::smem_u32(DYt), dQc_s = tc5::smem_u32(dQc);
    const uint32_t Wq_s = tc5::smem_u32(Wqkv_i), WqT_s = tc5::smem_u32(WqkvT_i), WoT_s = tc5::smem_u32(WoT_i);

    // weight-gradient jobs (16 x 16 output tiles), job id = warp + 16 j :
    //   per chunk: [gWqkv: MTq x NP]  A = dqkv_c (m-tile), B = LN(x) (n-pair) ; [gWoT: MTo x NP]  A = O_c, B = alpha*dout
    //   tile end : NP jobs each for dbo (ones x alpha*dout), dbeta (ones x G1), dgamma (ones x G2)
    const int MTq = NCc >> 4, MTo = Cc >> 4, NP = KC1 >> 1;
    const int per_chunk = (MTq + MTo) * NP;
    const int njobs = a.nchunks * per_chunk + 3 * NP;
    float acc[JW][2][4];
#pragma unroll
    for (int j = 0; j < JW; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) acc[j][q][0] = acc[j][q][1] = acc[j][q][2] = acc[j][q][3] = 0.f;

    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * a.SPT;
        const int nseq_t = (int)min((long long)a.SPT, a.nseq - s0);
        const int R = nseq_t * S;
        // ---- staging: threads 0..255 LN(x) -> At (+ row statistics), threads 256..511 alpha*dout -> DYt
        {
            const int tid2 = threadIdx.x & 255;
            const int row = tid2 >> 1, h = tid2 & 1;
            const bool valid = row < R;
            const int ls = row / S, pos = row - ls * S;
            const long long gr = valid ? a.g.grow(s0 + ls, pos) : 0;
            if (threadIdx.x < 256) stage_row_h<KCH, VEC4>(a.x + gr * D, valid, D, KC1, row, h, a.ln_w, a.ln_b, At, -1, stats);
            else stage_row_h<KCH, VEC4>(a.dout + gr * D, valid, D, KC1, row, h, nullptr, nullptr, DYt, -1, nullptr, a.alpha * gs);
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc5::fence_after_sync();
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_Q, tc5::kdesc(At_s, TILE_M, k), tc5::kdesc(Wq_s, NCq, k),
                             idesc_q, k > 0);
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_O, tc5::kdesc(DYt_s, TILE_M, k), tc5::kdesc(WoT_s, NDo, k),
                             idesc_o, k > 0);
            tc5::mma_commit(&mbar);
        }
        for (int ch = 0; ch < a.nchunks; ++ch) {
            tc5::mbar_wait(&mbar, phase);
            phase ^= 1;
            tc5::fence_after_sync();
            // ---- epilogue 1: q|k|v and dO accumulators -> bf16 tiles (16-column groups dealt round-robin to the 4 warps of a quadrant)
            {
                const int ngq = NCq >> 4, ngo = NDo >> 4;
                for (int gq = part; gq < ngq + ngo; gq += 4) {
                    float v[16];
                    const bool isq = gq < ngq;
                    const int gg = isq ? gq : gq - ngq;
                    tc5::tmem_ld16((isq ? tmem_Q : tmem_O) + lane_base + gg * 16, v);
                    tc5::tmem_ld_wait();
                    unsigned char* dst = isq ? QKVt : dOt;
                    const int KCx = isq ? KCq : KCd;
                    sts128(dst + tc5::toff(row_e, 2 * gg), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                           pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    sts128(dst + tc5::toff(row_e, 2 * gg + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]),
                           pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
                }
            }
            // park the weight-gradient accumulators (8 JW registers per thread, dead during the core) in spare TMEM
            // columns: the core is the register-hungry phase and the kernel sits at the 128-register cap
#pragma unroll
            for (int j = 0; j < JW; ++j) {
                const float t8[8] = {acc[j][0][0], acc[j][0][1], acc[j][0][2], acc[j][0][3],
                                     acc[j][1][0], acc[j][1][1], acc[j][1][2], acc[j][1][3]};
                tc5::tmem_st8(tmem_P + lane_base + j * 8, t8);
            }
            tc5::tmem_st_wait();
            tc5::fence_before_sync();
            __syncthreads();
            attn_bwd_core_bf16<DH>(QKVt, dOt, dQc, Ot, hc, nseq_t, S, a.scale, warp, NW, lane, cl, coltab);
            tc5::fence_proxy_async();
            __syncthreads();
#pragma unroll
            for (int j = 0; j < JW; ++j) {
                float t8[8];
                tc5::tmem_ld8(tmem_P + lane_base + j * 8, t8);
                tc5::tmem_ld_wait();
                acc[j][0][0] = t8[0]; acc[j][0][1] = t8[1]; acc[j][0][2] = t8[2]; acc[j][0][3] = t8[3];
                acc[j][1][0] = t8[4]; acc[j][1][1] = t8[5]; acc[j][1][2] = t8[6]; acc[j][1][3] = t8[7];
            }
            if (threadIdx.x == 0) {
                tc5::fence_after_sync();
                const uint32_t wt = WqT_s + (uint32_t)ch * Kp * NCc * 2;
                for (int k = 0; k < NCc / 16; ++k)
                    tc5::mma_f16(tmem_A, tc5::kdesc(dQc_s, TILE_M, k), tc5::kdesc(wt, Kp, k),
                                 idesc_a, (ch > 0 || k > 0) ? 1u : 0u);
                if (ch + 1 < a.nchunks) {
                    const uint32_t wq = Wq_s + (uint32_t)(ch + 1) * NCq * Kp * 2, wo = WoT_s + (uint32_t)(ch + 1) * NDo * Kp * 2;
                    for (int k = 0; k < Kp / 16; ++k)
                        tc5::mma_f16(tmem_Q, tc5::kdesc(At_s, TILE_M, k), tc5::kdesc(wq, NCq, k),
                                     idesc_q, k > 0);
                    for (int k = 0; k < Kp / 16; ++k)
                        tc5::mma_f16(tmem_O, tc5::kdesc(DYt_s, TILE_M, k), tc5::kdesc(wo, NDo, k),
                                     idesc_o, k > 0);
                }
                tc5::mma_commit(&mbar);
            }
            // ---- weight-gradient jobs of this chunk (overlap the dA MMA)
#pragma unroll
            for (int j = 0; j < JW; ++j) {
                const int job = warp + NW * j;
                const int rel = job - ch * per_chunk;
                if (rel >= 0 && rel < per_chunk && job < njobs) {
                    if (rel < MTq * NP) wgrad_job(dQc, KCc, 2 * (rel / NP), false, At, KC1, 2 * (rel % NP), lane, acc[j]);
                    else { const int r2 = rel - MTq * NP; wgrad_job(Ot, KCo, 2 * (r2 / NP), false, DYt, KC1, 2 * (r2 % NP), lane, acc[j]); }
                }
            }
        }
        tc5::mbar_wait(&mbar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 2: LayerNorm backward.  g = dA ; gg = g*gamma ; dx = base + rstd*(gg - mean(gg) - xhat*mean(gg*xhat))
        {
            const int ls = row_e / S, pos = row_e - ls * S;
            const bool valid = row_e < R;
            const long long gr = valid ? a.g.grow(s0 + ls, pos) : 0;
            const float mean = stats[2 * row_e], rstd = stats[2 * row_e + 1];
            const int ngr = (D + 7) >> 3;
            float gg[2][8], xh[2][8];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int gq = part + 4 * u;
#pragma unroll
                for (int k = 0; k < 8; ++k) { gg[u][k] = 0.f; xh[u][k] = 0.f; }
                if (gq < ngr) {
                    float v[8], xv[8];
                    tc5::tmem_ld8(tmem_A + lane_base + gq * 8, v);
                    if (valid) load8<VEC4>(a.x + gr * D, gq * 8, D, xv);
                    tc5::tmem_ld_wait();
                    float g1[8], g2[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int c = gq * 8 + k;
                        const bool okc = valid && c < D;
                        const float gv = okc ? v[k] : 0.f;
                        xh[u][k] = okc ? (xv[k] - mean) * rstd : 0.f;
                        gg[u][k] = okc ? gv * __ldg(a.ln_w + c) : 0.f;
                        s1 += gg[u][k];
                        s2 = fmaf(gg[u][k], xh[u][k], s2);
                        g1[k] = gv; g2[k] = gv * xh[u][k];
                    }
                    sts128(G1t + tc5::toff(row_e, gq), pack_h2(g1[0], g1[1]), pack_h2(g1[2], g1[3]),
                           pack_h2(g1[4], g1[5]), pack_h2(g1[6], g1[7]));
                    sts128(G2t + tc5::toff(row_e, gq), pack_h2(g2[0], g2[1]), pack_h2(g2[2], g2[3]),
                           pack_h2(g2[4], g2[5]), pack_h2(g2[6], g2[7]));
                } else if (gq < KC1) {
                    sts128(G1t + tc5::toff(row_e, gq), 0u, 0u, 0u, 0u);
                    sts128(G2t + tc5::toff(row_e, gq), 0u, 0u, 0u, 0u);
                }
            }
            parts[(row_e * 4 + part) * 2] = s1;
            parts[(row_e * 4 + part) * 2 + 1] = s2;
            tc5::fence_before_sync();
            __syncthreads();
            const float invD = 1.0f / (float)D;
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { t1 += parts[(row_e * 4 + q) * 2]; t2 += parts[(row_e * 4 + q) * 2 + 1]; }
            t1 *= invD; t2 *= invD;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int gq = part + 4 * u;
                if (gq < ngr && valid) {
                    float bv[8], ov[8];
                    if (a.base) load8<VEC4>(a.base + gr * D, gq * 8, D, bv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        ov[k] = (rstd * inv_gs) * (gg[u][k] - t1 - xh[u][k] * t2);
                        if (a.base) ov[k] += bv[k];
                    }
                    if (a.out_drop_p > 0.f) {
                        const unsigned long long e0 = (unsigned long long)gr * D + gq * 8;
                        const uint32_t strm = rng_stream_of_step(a.rng_stream, a.rng_step);
                        const float inv_keep = 1.0f / (1.0f - a.out_drop_p);
                        if ((e0 & 7ull) == 0ull) {                       // D % 8 == 0: the 8 decisions of one counter
                            const uint4 bits = dropout_bits8(a.seed, strm, e0 >> 3);
                            const uint32_t thr = dropout_threshold(a.out_drop_p);
#pragma unroll
                            for (int k = 0; k < 8; ++k) ov[k] *= dropout_lane16(bits, k) < thr ? 0.f : inv_keep;
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (gq * 8 + k < D) ov[k] *= dropout_scale(a.seed, strm, e0 + k, a.out_drop_p, inv_keep);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) dx_max = fmaxf(dx_max, fabsf(ov[k]));
                    store8<VEC4>(a.dx + gr * D, gq * 8, D, ov);
                }
            }
        }
        // ---- pull the rows the NEXT tile will stage towards L2 (x was written a whole forward pass ago: DRAM) while the
        //      tile-end jobs run; no registers are held
        if (tile + gridDim.x < ntiles) {
            const long long s1 = (tile + gridDim.x) * a.SPT;
            const int R1 = (int)min((long long)a.SPT, a.nseq - s1) * S;
            const int tid2 = threadIdx.x & 255, row = tid2 >> 1;
            if (row < R1) {
                const int ls = row / S, pos = row - ls * S;
                const float* p = (threadIdx.x < 256 ? a.x : a.dout) + a.g.grow(s1 + ls, pos) * D + (tid2 & 1) * 32;
                if ((tid2 & 1) * 32 < D) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
        // ---- tile-end jobs: dbo = colsum(alpha*dout), dbeta = colsum(g), dgamma = colsum(g*xhat)
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            const int job = warp + NW * j;
            const int rel = job - a.nchunks * per_chunk;
            if (rel >= 0 && job < njobs) {
                const int which = rel / NP, np = rel - which * NP;
                wgrad_job(At, KC1, 0, true, which == 0 ? DYt : which == 1 ? G1t : G2t, KC1, 2 * np, lane, acc[j]);
            }
        }
        __syncthreads();                    // all warps are done with this tile's tiles before the next staging
    }
    // ---- per-CTA gradient record: [nchunks][NCc][Kp] gWqkv_c | [nchunks][Cc][Kp] gWoT_c | [3][Kp] dbo, dbeta, dgamma
    {
        float* rec = a.partials + (size_t)blockIdx.x * a.psize;
        float* recO = rec + (size_t)a.nchunks * NCc * Kp;
        float* recS = recO + (size_t)a.nchunks * Cc * Kp;
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            const int job = warp + NW * j;
            if (job >= njobs) continue;
            if (job < a.nchunks * per_chunk) {
                const int ch = job / per_chunk, rel = job - ch * per_chunk;
                if (rel < MTq * NP) wgrad_store(rec + (size_t)ch * NCc * Kp, Kp, 16 * (rel / NP), 16 * (rel % NP), lane, acc[j], false, inv_gs);
                else { const int r2 = rel - MTq * NP; wgrad_store(recO + (size_t)ch * Cc * Kp, Kp, 16 * (r2 / NP), 16 * (r2 % NP), lane, acc[j], false, inv_gs); }
            } else {
                const int rel = job - a.nchunks * per_chunk;
                wgrad_store(recS + (size_t)(rel / NP) * Kp, Kp, 0, 16 * (rel % NP), lane, acc[j], true, inv_gs);
            }
        }
    }
    publish_amax_block(a.dx_amax, dx_max);
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tmem_base_s, 512);
}

struct AttnReduceTcArgs {
    const float* partials; int nparts, psize;
    float* dWq; float* dWk; float* dWv; float* dWo; float* dbo; float* dln_w; float* dln_b;
    int accumulate_wq, D, I, dh, hc, nchunks, Kp, NCc, Cc;
    float qmul, kmul;        // multipliers of the dWq / dWk records (the register-resident kernel stores dq / scale, dk / ln2)
};
static __global__ void k_reduce_attn_tc(AttnReduceTcArgs a) {
    pdl_launch_dependents();            // the next backward kernel may run its prologue under this reduction (common.cuh)
    const int D = a.D, I = a.I;
    const int total = 4 * I * D + 3 * D;
    const size_t offO = (size_t)a.nchunks * a.NCc * a.Kp, offS = offO + (size_t)a.nchunks * a.Cc * a.Kp;
    for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
        const int i = min(base + (int)threadIdx.x, total - 1);
        const bool active = base + (int)threadIdx.x < total;
        size_t src;
        float* dst;
        bool accf = false;
        if (i < 3 * I * D) {
            const int which = i / (I * D), rem = i - which * (I * D);
            const int row = rem / D, d = rem - row * D;               // row = head * dh + dd
            const int ch = row / (a.hc * a.dh), rl = row - ch * (a.hc * a.dh);
            src = ((size_t)ch * a.NCc + which * a.hc * a.dh + rl) * a.Kp + d;
            dst = which == 0 ? a.dWq : which == 1 ? a.dWk : a.dWv;
            accf = which == 0 && a.accumulate_wq;
            if (dst) dst += rem;
        } else if (i < 4 * I * D) {
            const int rem = i - 3 * I * D;
            const int d = rem / I, col = rem - d * I;
            const int ch = col / (a.hc * a.dh), cl = col - ch * (a.hc * a.dh);
            src = offO + ((size_t)ch * a.Cc + cl) * a.Kp + d;
            dst = a.dWo ? a.dWo + rem : nullptr;
        } else {
            const int rem = i - 4 * I * D;
            const int which = rem / D, d = rem - which * D;            // 0: dbo, 1: dgamma, 2: dbeta
            src = offS + (size_t)(which == 0 ? 0 : which == 1 ? 2 : 1) * a.Kp + d;
            float* b = which == 0 ? a.dbo : which == 1 ? a.dln_w : a.dln_b;
            dst = b ? b + d : nullptr;
        }
        float s = record_sum_sliced(a.partials, a.psize, a.nparts, src, active && dst != nullptr);
        if (i < I * D) s *= a.qmul; else if (i < 2 * I * D) s *= a.kmul;
        if (active && dst != nullptr && threadIdx.y == 0) *dst = accf ? *dst + s : s;
    }
}


}  // namespace rat
