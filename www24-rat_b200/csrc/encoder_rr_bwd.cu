// K5 backward, register-resident generation (see encoder_rr.cuh).  Backward of
//     out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )        (reference: autograd of models/RAT_m2.py:176-236)
// dx = base + dLN(...) and all parameter gradients, for sequences <= 16 tokens, head width <= 16 (even), D <= 48.
//
// 8 warps per CTA, one CTA per SM (255 registers per thread), tiles of 8 warp tasks (128 fragment rows).
//   compute warp, per task (one sequence, or two of <= 8 tokens) -- everything in registers, mma.sync fragments:
//     LN(x), alpha*gs*dout -> A fragments ; per head: q, k, v^T, dO projections ; S, P ; dP = dO V^T ; dS ; O = P V ;
//     dV = P^T dO ; dQ = dS K ; dK = dS^T Q   (operand transposes by movmatrix on packed 8x8 blocks)
//     -> fp16 rows [LN(x)] [alpha*gs*dout] [dq|dk|dv compact] [O compact] of the CTA's 128-row token tile in shared memory
//        (UMMA canonical chunk-major layout), then one atomic arrival per warp -- no CTA barrier.
//   the LAST warp to arrive issues the tile's tcgen05 products (accumulators in TMEM) and moves on:
//     dA[128 x Kp]  = [dq|dk|dv] . Wqkv            (tile K-major, resident weight image)
//     gWqkv^T      += [dq|dk|dv]^T . LN(x)         (both operands MN-major views of the same token tile, K = 128 tokens)
//     gWo^T        += O^T . (alpha*gs*dout)        -- accumulated over ALL tiles of the CTA, written once at the end
//   compute warps, one tile later (deferred, so the MMA latency hides under the next tile's head loop):
//     dA from TMEM (thread = token row) -> LayerNorm backward in fp32 -> dx ; dgamma / dbeta / dbo partial sums in registers.
// Per-CTA gradient records are summed in fixed order by k_reduce_attn_tc => bitwise run-to-run deterministic.
#include "encoder_rr.cuh"
#include "encoder_tc_attnbwd.cuh"     // AttnReduceTcArgs / k_reduce_attn_tc (record layout shared with the tile kernel)

namespace rat {

struct AttnBwdRRArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo;
    float* partials;
    const float* dout_amax; float* dx_amax;
    long long nseq;
    SeqGeom g;
    int D, H, I, dh;
    float scale, alpha;
    int hc, nchunks;         // heads per chunk (one round of tile -> MMA), chunks per tile
    int NCc;                 // round_up(3 * hc * dh, 128): compact dq|dk|dv columns of the tile (zero padded)
    int Cc;                  // round_up(hc * dh, 128):     compact O columns, stored after the dq|dk|dv columns
    int psize, smem_bytes, tmem_cols;
    float out_drop_p; unsigned long long seed; unsigned int rng_stream; const unsigned int* rng_step;
};

constexpr int RRB_CWARPS = 8;
constexpr int RRB_THREADS = RRB_CWARPS * 32;

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 5, 256;" ::: "memory"); }

template <int KS, int NTO, bool VEC4, bool F16P, bool F16C, int UNR>
__global__ void __launch_bounds__(RRB_THREADS, 1) k_attn_bwd_rr(AttnBwdRRArgs a) {
    extern __shared__ __align__(128) unsigned char rrb_smem[];
    constexpr int Kp = 16 * KS, KC1 = 2 * KS;
    constexpr int GH = (NTO + 1) / 2;                    // 8-column groups per epilogue half
    const int H = a.H, D = a.D, dh = a.dh, hc = a.hc, NCc = a.NCc, Cc = a.Cc;
    const int MB = (NCc + Cc) >> 7, MBq = NCc >> 7;      // 128-column blocks of the weight-gradient products
    const int ksA = (3 * hc * dh + 15) >> 4;             // k-steps of the dA product (real compact columns)
    // ---- shared memory carve-up
    uint4* Wq_i = reinterpret_cast<uint4*>(rrb_smem);    // fragment-order images [H][KS][32]
    uint4* Wk_i = Wq_i + H * KS * 32;
    uint4* Wv_i = Wk_i + H * KS * 32;
    uint4* Wd_i = Wv_i + H * KS * 32;                    // dO = dY . Wo_h   (n = head column, k = model column)
    unsigned char* WT_i = reinterpret_cast<unsigned char*>(Wd_i + H * KS * 32);      // [nchunks][Kp x NCc] canonical K-major
    unsigned char* Xt = WT_i + (size_t)a.nchunks * Kp * NCc * 2;                     // [128 x Kp]   LN(x)
    unsigned char* DYt = Xt + (size_t)KC1 * tc5::TILE_CHUNK;                          // [128 x Kp]   alpha*gs*dout
    unsigned char* Gt = DYt + (size_t)KC1 * tc5::TILE_CHUNK;                          // [128 x (NCc + Cc)]
    float* lnw_s = reinterpret_cast<float*>(Gt + (size_t)((NCc + Cc) >> 3) * tc5::TILE_CHUNK);   // [Kp]
    float* lnb_s = lnw_s + Kp;
    // row statistics / global row index of the tile rows, THREE buffers (tile % 3): a warp may already record tile i+1 while
    // a slower warp still runs the deferred epilogue of tile i-1 (warps are coupled only through the two mbarriers)
    float* stats = lnb_s + Kp;                           // [3][128][2] mean, rstd
    float* parts = stats + 3 * 128 * 2;                  // [128][2][2] LayerNorm-backward row partials of the two column halves
    long long* growS = reinterpret_cast<long long*>(parts + 128 * 4);                 // [3][128] global row (-1: absent)
    uint32_t* coltab = reinterpret_cast<uint32_t*>(growS + 3 * 128);                  // [NCc + Cc] column -> byte offset in a tile row
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int arrive_cnt;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    pdl_launch_dependents();
    float gs = 1.0f, inv_gs = 1.0f;                       // gradient scale: read after pdl_wait (the slot is published by the previous kernel)
    // ---- images (raw weights first, with coalesced loads, into the token-tile region; then fragment order from shared memory)
    {
        float* raw = reinterpret_cast<float*>(Xt);        // Wq | Wk | Wv [3][I][D], Wo [D][I]: 16 I D bytes <= the tile region (plan)
        const int nw = a.I * D;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            raw[i] = __ldg(a.Wq + i); raw[nw + i] = __ldg(a.Wk + i); raw[2 * nw + i] = __ldg(a.Wv + i); raw[3 * nw + i] = __ldg(a.Wo + i);
        }
        for (int i = threadIdx.x; i < Kp; i += blockDim.x) {
            lnw_s[i] = i < D ? a.ln_w[i] : 0.f;
            lnb_s[i] = i < D ? a.ln_b[i] : 0.f;
        }
        for (int c = threadIdx.x; c < NCc + Cc; c += blockDim.x) coltab[c] = (uint32_t)(c >> 3) * tc5::TILE_CHUNK + (uint32_t)(c & 7) * 2u;
        __syncthreads();
        const int nqkv = H * KS * 32;
        for (int i = threadIdx.x; i < 4 * nqkv; i += blockDim.x) {
            const int w = i / nqkv, r = i - w * nqkv;
            const int h = r / (KS * 32), ks = (r >> 5) % KS, ln = r & 31;
            if (w < 3) {
                const float* W = raw + w * nw + h * dh * D;
                const float mul = w == 0 ? a.scale * 1.4426950408889634f : 1.0f;
                Wq_i[i] = frag_pair_entry(ln, 0, 16 * ks, [&](int n, int k) { return (n < dh && k < D) ? mul * W[n * D + k] : 0.f; });
            } else {
                const float* W = raw + 3 * nw + h * dh;
                Wq_i[i] = frag_pair_entry(ln, 0, 16 * ks, [&](int dd, int c) { return (dd < dh && c < D) ? W[c * a.I + dd] : 0.f; });
            }
        }
        const int KCc = NCc >> 3, perT = Kp * KCc;
        for (int i = threadIdx.x; i < a.nchunks * perT; i += blockDim.x) {
            const int ch = i / perT, rem = i - ch * perT;
            const int d = rem % Kp, kc = rem / Kp;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int nc = kc * 8 + k;                       // compact column: [q | k | v] x [hl][dd]
                const int w = nc / (hc * dh), rem2 = nc - w * (hc * dh);
                const float mul = w == 0 ? a.scale : w == 1 ? 0.6931471805599453f : 1.0f;     // dq = scale dQ', dk = ln2 dK'
                v[k] = (w < 3 && d < D) ? mul * raw[w * nw + (ch * hc * dh + rem2) * D + d] : 0.f;
            }
            sts128(WT_i + (size_t)ch * Kp * NCc * 2 + tc5::kmajor_off(d, kc, Kp), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                   pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        __syncthreads();
        const int tile16 = (2 * KC1 + ((NCc + Cc) >> 3)) * (int)tc5::TILE_CHUNK / 16;   // pad columns are never written: zero once
        for (int i = threadIdx.x; i < tile16; i += blockDim.x) reinterpret_cast<uint4*>(Xt)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) { tc5::mbar_init(&done_bar, 1); tc5::fence_mbar_init(); arrive_cnt = 0u; }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    pdl_wait();                                           // x / dout / base / the amax slot come from the previous kernels
    gs = tc_grad_scale(a.dout_amax);
    inv_gs = 1.0f / gs;
    const uint32_t tmem_W = tmem_base_s;                                   // [nchunks][MB] blocks of Kp columns
    const uint32_t tmem_A = tmem_W + (uint32_t)(a.nchunks * MB * Kp);      // dA, Kp columns

    const int S = a.g.S;
    const bool packed = S <= 8;
    const long long ntasks = packed ? (a.nseq + 1) >> 1 : a.nseq;
    const long long ntiles = (ntasks + RRB_CWARPS - 1) / RRB_CWARPS;

    const uint32_t Xt_s = tc5::smem_u32(Xt), DYt_s = tc5::smem_u32(DYt), Gt_s = tc5::smem_u32(Gt), WT_s = tc5::smem_u32(WT_i);
    const uint32_t idesc_a = tc5::instr_desc(TC_FMT, 128, Kp);
    const uint32_t idesc_w = tc5::instr_desc(TC_FMT, 128, Kp, 1, 1);
    // the tile's tensor-core products (whole converged warp, one elected lane issues)
    auto issue_products = [&](int ch, bool first) {
        const uint32_t wt = WT_s + (uint32_t)ch * Kp * NCc * 2;
        for (int k = 0; k < ksA; ++k)
            tc5::mma_f16_w(tmem_A, tc5::kdesc(Gt_s, 128, k), tc5::kdesc(wt, Kp, k), idesc_a, (ch > 0 || k > 0) ? 1u : 0u);
        for (int b = 0; b < MB; ++b) {
            const uint32_t bt = b < MBq ? Xt_s : DYt_s;
            const uint32_t at = Gt_s + (uint32_t)b * 16u * tc5::TILE_CHUNK;
            const uint32_t td = tmem_W + (uint32_t)((ch * MB + b) * Kp);
            for (int j = 0; j < 8; ++j)              // K = 128 token rows in steps of 16 (two 128-byte core matrices)
                tc5::mma_f16_w(td, tc5::smem_desc(at + j * 256, 128u, tc5::TILE_CHUNK), tc5::smem_desc(bt + j * 256, 128u, tc5::TILE_CHUNK),
                               idesc_w, (!first || j > 0) ? 1u : 0u);
        }
        tc5::mma_commit_w(&done_bar);
    };
    float dx_max = 0.f;
    float acc_bo[NTO][2];            // dbo partial: columns 8 nt + 2t, +1 over this thread's rows
    float acc_g[GH][8], acc_b[GH][8];   // dgamma / dbeta partials: row = thread, columns of this thread's epilogue half
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) acc_bo[nt][0] = acc_bo[nt][1] = 0.f;
#pragma unroll
    for (int u = 0; u < GH; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc_g[u][k] = acc_b[u][k] = 0.f;
    const int erow = (warp & 3) * 32 + lane, ehalf = (warp >> 2) & 1;       // epilogue: token row of the tile, column half
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

    // LayerNorm backward of the tile whose rows were recorded in buffer pp (dA is complete in TMEM)
    auto epilogue = [&](int pp) {
        const long long gr = growS[pp * 128 + erow];
        const bool valid = gr >= 0;
        const float mean = stats[(pp * 128 + erow) * 2], rstd = stats[(pp * 128 + erow) * 2 + 1];
        float gg[GH][8], xh[GH][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < GH; ++u) {
            const int gq = ehalf * GH + u;
#pragma unroll
            for (int k = 0; k < 8; ++k) { gg[u][k] = 0.f; xh[u][k] = 0.f; }
            if (gq < NTO) {
                float v[8], xv[8];
                tc5::tmem_ld8(tmem_A + lane_base + gq * 8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) xv[k] = 0.f;
                if (valid) load8<VEC4>(a.x + gr * D, gq * 8, D, xv);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = gq * 8 + k;
                    const bool okc = valid && c < D;
                    const float gv = okc ? v[k] : 0.f;
                    xh[u][k] = okc ? (xv[k] - mean) * rstd : 0.f;
                    gg[u][k] = gv * lnw_s[c];
                    s1 += gg[u][k];
                    s2 = fmaf(gg[u][k], xh[u][k], s2);
                    acc_b[u][k] += gv;
                    acc_g[u][k] = fmaf(gv, xh[u][k], acc_g[u][k]);
                }
            }
        }
        parts[(erow * 2 + ehalf) * 2] = s1;
        parts[(erow * 2 + ehalf) * 2 + 1] = s2;
        tc5::fence_before_sync();
        pair_sync(1 + (warp & 3));
        const float invD = 1.0f / (float)D;
        const float t1 = (parts[erow * 4] + parts[erow * 4 + 2]) * invD, t2 = (parts[erow * 4 + 1] + parts[erow * 4 + 3]) * invD;
        pair_sync(1 + (warp & 3));                       // parts are rewritten by the next epilogue
#pragma unroll
        for (int u = 0; u < GH; ++u) {
            const int gq = ehalf * GH + u;
            if (gq < NTO && valid) {
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
                    if ((e0 & 7ull) == 0ull) {
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
                for (int k = 0; k < 8; ++k)
                    if (gq * 8 + k < D) dx_max = fmaxf(dx_max, fabsf(ov[k]));
                store8<VEC4>(a.dx + gr * D, gq * 8, D, ov);
            }
        }
    };

    {
        const RRLane cl = make_rr_lane(S, lane);
        const float invD_ = 1.0f / (float)D;
        const uint32_t row_lo = (uint32_t)(warp * 16 + g) * 16u, row_hi = row_lo + 128u;     // byte offsets of the tile rows
        uint32_t dph = 0;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int pp = it % 3;
            const long long task = tile * RRB_CWARPS + warp;
            const bool tv = task < ntasks;
            const long long seq0 = packed ? 2 * task : task;
            const bool vlo = tv && cl.lo_pos >= 0, vhi = tv && cl.hi_pos >= 0 && seq0 + cl.hi_sq < a.nseq;
            const long long rlo = vlo ? a.g.grow(seq0, cl.lo_pos) : 0, rhi = vhi ? a.g.grow(seq0 + cl.hi_sq, cl.hi_pos) : 0;
            // ---- LN(x) and alpha*gs*dout -> A fragments
            uint32_t xa[KS][4], da[KS][4];
            {
                float2 xl[NTO], xh2[NTO];
                rr_load_rows<NTO>(a.x + rlo * D, a.x + rhi * D, vlo, vhi, D, t, xl, xh2);
                float2 dl[NTO], dh2[NTO];
                rr_load_rows<NTO>(a.dout + rlo * D, a.dout + rhi * D, vlo, vhi, D, t, dl, dh2);
                // pull the rows of this warp's NEXT tile towards L2 (x was written a whole forward pass ago)
                if (tile + gridDim.x < ntiles) {
                    const long long task1 = task + (long long)gridDim.x * RRB_CWARPS;
                    const long long seq1 = packed ? 2 * task1 : task1;
                    const int r16 = lane & 15;
                    const int p_sq = packed ? r16 >> 3 : 0, p_pos = packed ? r16 & 7 : r16;
                    if (p_pos < S && seq1 + p_sq < a.nseq) {
                        const float* p = (lane < 16 ? a.x : a.dout) + a.g.grow(seq1 + p_sq, p_pos) * D;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                        if (D > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
                    }
                }
                float ml, rl, mh, rh;
                rr_row_stats<NTO>(xl, D, invD_, t, ml, rl);
                rr_row_stats<NTO>(xh2, D, invD_, t, mh, rh);
                if (t == 0) {
                    stats[(pp * 128 + warp * 16 + g) * 2] = ml; stats[(pp * 128 + warp * 16 + g) * 2 + 1] = rl;
                    stats[(pp * 128 + warp * 16 + g + 8) * 2] = mh; stats[(pp * 128 + warp * 16 + g + 8) * 2 + 1] = rh;
                    growS[pp * 128 + warp * 16 + g] = vlo ? rlo : -1;
                    growS[pp * 128 + warp * 16 + g + 8] = vhi ? rhi : -1;
                }
                const float dmul = a.alpha * gs;
#pragma unroll
                for (int nt = 0; nt < 2 * KS; ++nt) {
                    uint32_t lo = 0u, hi = 0u, dlo = 0u, dhi = 0u;
                    if (nt < NTO) {
                        const float2 w = *reinterpret_cast<const float2*>(lnw_s + 8 * nt + 2 * t);
                        const float2 b = *reinterpret_cast<const float2*>(lnb_s + 8 * nt + 2 * t);
                        lo = vlo ? pack_h2(fmaf((xl[nt].x - ml) * rl, w.x, b.x), fmaf((xl[nt].y - ml) * rl, w.y, b.y)) : 0u;
                        hi = vhi ? pack_h2(fmaf((xh2[nt].x - mh) * rh, w.x, b.x), fmaf((xh2[nt].y - mh) * rh, w.y, b.y)) : 0u;
                        const float a0 = dl[nt].x * dmul, a1 = dl[nt].y * dmul, b0 = dh2[nt].x * dmul, b1 = dh2[nt].y * dmul;
                        dlo = pack_h2(a0, a1); dhi = pack_h2(b0, b1);
                        acc_bo[nt][0] += a0 + b0; acc_bo[nt][1] += a1 + b1;
                    }
                    xa[nt >> 1][(nt & 1) * 2] = lo; xa[nt >> 1][(nt & 1) * 2 + 1] = hi;
                    da[nt >> 1][(nt & 1) * 2] = dlo; da[nt >> 1][(nt & 1) * 2 + 1] = dhi;
                }
            }
            // ---- previous tile: its products are complete -> LayerNorm backward; the token tile is free again
            if (it > 0) {
                tc5::mbar_wait(&done_bar, dph);
                dph ^= 1;
                tc5::fence_after_sync();
                epilogue((it - 1) % 3);
            }
            // ---- LN(x) and dY rows of the token tile
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint32_t c0 = (uint32_t)(2 * ks) * tc5::TILE_CHUNK + 4u * t, c1 = c0 + tc5::TILE_CHUNK;
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c0 + row_lo), "r"(xa[ks][0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c0 + row_hi), "r"(xa[ks][1]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c1 + row_lo), "r"(xa[ks][2]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c1 + row_hi), "r"(xa[ks][3]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c0 + row_lo), "r"(da[ks][0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c0 + row_hi), "r"(da[ks][1]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c1 + row_lo), "r"(da[ks][2]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c1 + row_hi), "r"(da[ks][3]) : "memory");
            }
            for (int ch = 0; ch < a.nchunks; ++ch) {
                if (ch > 0) {                               // the dq|dk|dv|O columns are reused by every chunk
                    tc5::mbar_wait(&done_bar, dph);
                    dph ^= 1;
                }
                const uint4* wq = Wq_i + (size_t)ch * hc * KS * 32 + lane;
                const uint4* wk = Wk_i + (size_t)ch * hc * KS * 32 + lane;
                const uint4* wv = Wv_i + (size_t)ch * hc * KS * 32 + lane;
                const uint4* wd = Wd_i + (size_t)ch * hc * KS * 32 + lane;
#pragma unroll UNR
                for (int hl = 0; hl < hc; ++hl) {
                    // the core's outputs (one k-step each) accumulate in fp16: packed accumulators are tile rows as they are;
                    // the projections (three k-steps) do the same when F16P
                    ProjAcc<F16P> q, k, vt, dO;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint4 fq = wq[ks * 32], fk = wk[ks * 32], fv = wv[ks * 32], fd = wd[ks * 32];
                        q.mma(0, xa[ks], fq.x, fq.y);
                        q.mma(1, xa[ks], fq.z, fq.w);
                        k.mma(0, xa[ks], fk.x, fk.y);
                        k.mma(1, xa[ks], fk.z, fk.w);
                        const uint32_t av[4] = {fv.x, fv.z, fv.y, fv.w};
                        vt.mma(0, av, xa[ks][0], xa[ks][2]);
                        vt.mma(1, av, xa[ks][1], xa[ks][3]);
                        dO.mma(0, da[ks], fd.x, fd.y);
                        dO.mma(1, da[ks], fd.z, fd.w);
                    }
                    wq += KS * 32; wk += KS * 32; wv += KS * 32; wd += KS * 32;
                    // 8x8 blocks {rows g / cols 0-7, rows g+8 / cols 0-7, rows g / cols 8-15, rows g+8 / cols 8-15} of q, k, dO, v^T
                    uint32_t qa[4], ka[4], doa[4], va[4];
                    q.frag(qa); k.frag(ka); dO.frag(doa); vt.frag(va);
                    const uint32_t v00 = va[0], v10 = va[1], v01 = va[2], v11 = va[3];     // [dd block][token block]
                    // ---- S = q k^T (keys 0..7: rows g of k, keys 8..15: rows g + 8) ; dP = dO v^T   (fp32)
                    float sc[2][4] = {}, dp[2][4] = {};
                    rr_mma(sc[0], qa, ka[0], ka[2]);
                    rr_mma(sc[1], qa, ka[1], ka[3]);
                    rr_mma(dp[0], doa, rr_movm(v00), rr_movm(v10));
                    rr_mma(dp[1], doa, rr_movm(v01), rr_movm(v11));
                    rr_softmax(sc, cl, vlo, vhi);                                             // sc = P
                    float dlo = 0.f, dhi = 0.f;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (e < 2) dlo = fmaf(sc[nt][e], dp[nt][e], dlo); else dhi = fmaf(sc[nt][e], dp[nt][e], dhi);
                        }
                    dlo = qsum(dlo); dhi = qsum(dhi);
                    uint32_t pa[4], sa[4];
                    c_to_a(sc, pa);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) dp[nt][e] = sc[nt][e] * (dp[nt][e] - ((e < 2) ? dlo : dhi));    // dS
                    c_to_a(dp, sa);
                    const uint32_t pt[4] = {rr_movm(pa[0]), rr_movm(pa[2]), rr_movm(pa[1]), rr_movm(pa[3])};             // P^T
                    const uint32_t st[4] = {rr_movm(sa[0]), rr_movm(sa[2]), rr_movm(sa[1]), rr_movm(sa[3])};             // dS^T
                    // ---- O = P v ; dV = P^T dO ; dQ' = dS k ; dK' = dS^T q'   (dq = scale dQ', dk = ln2 dK': folded into the
                    //      dA weight image and the record reduction, so the packed accumulators go to the tile as they are)
                    ProjAcc<F16C> o, dv, dq, dk;
                    o.mma(0, pa, v00, v01);
                    o.mma(1, pa, v10, v11);
                    dv.mma(0, pt, rr_movm(doa[0]), rr_movm(doa[1]));
                    dv.mma(1, pt, rr_movm(doa[2]), rr_movm(doa[3]));
                    dq.mma(0, sa, rr_movm(ka[0]), rr_movm(ka[1]));
                    dq.mma(1, sa, rr_movm(ka[2]), rr_movm(ka[3]));
                    dk.mma(0, st, rr_movm(qa[0]), rr_movm(qa[1]));
                    dk.mma(1, st, rr_movm(qa[2]), rr_movm(qa[3]));
                    uint32_t of[4], dvf[4], dqf[4], dkf[4];          // {rows g / d 0-7, rows g+8 / d 0-7, rows g / d 8-15, rows g+8 / d 8-15}
                    o.frag(of); dv.frag(dvf); dq.frag(dqf); dk.frag(dkf);
                    // ---- compact fp16 rows of the token tile: dq | dk | dv at columns [part * hc*dh + hl*dh + d], O after them
#pragma unroll
                    for (int nd = 0; nd < 2; ++nd) {
                        const int d = 8 * nd + 2 * t;
                        if (d < dh) {
                            const int cq = hl * dh + d;
                            const uint32_t oq = Gt_s + coltab[cq], ok = Gt_s + coltab[cq + hc * dh], ov = Gt_s + coltab[cq + 2 * hc * dh];
                            const uint32_t oo = Gt_s + coltab[NCc + cq];
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(oq + row_lo), "r"(dqf[2 * nd]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(oq + row_hi), "r"(dqf[2 * nd + 1]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ok + row_lo), "r"(dkf[2 * nd]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ok + row_hi), "r"(dkf[2 * nd + 1]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ov + row_lo), "r"(dvf[2 * nd]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ov + row_hi), "r"(dvf[2 * nd + 1]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(oo + row_lo), "r"(of[2 * nd]) : "memory");
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(oo + row_hi), "r"(of[2 * nd + 1]) : "memory");
                        }
                    }
                }
                // ---- this warp's 16 rows of the tile are complete; the last of the 8 warps to get here issues the products
                //      (rounds cannot overlap: nobody arrives for the next round before this round's commit is waited on)
                tc5::fence_proxy_async();
                tc5::fence_before_sync();
                __threadfence_block();
                __syncwarp();
                unsigned int cnt = 0u;
                if (lane == 0) cnt = atomicAdd(&arrive_cnt, 1u);
                cnt = __shfl_sync(0xffffffffu, cnt, 0);
                if ((cnt & (RRB_CWARPS - 1)) == RRB_CWARPS - 1) {
                    __threadfence_block();
                    tc5::fence_after_sync();
                    issue_products(ch, it == 0);
                }
            }
        }
        if (it > 0) {                                       // the last tile of this CTA
            tc5::mbar_wait(&done_bar, dph);
            dph ^= 1;
            tc5::fence_after_sync();
            epilogue((it - 1) % 3);
        }
        // ---- per-CTA gradient record: [nchunks][NCc][Kp] gWqkv_c | [nchunks][Cc][Kp] gWoT_c | [3][Kp] dbo, dbeta, dgamma
        float* rec = a.partials + (size_t)blockIdx.x * a.psize;
        float* recO = rec + (size_t)a.nchunks * NCc * Kp;
        float* recS = recO + (size_t)a.nchunks * Cc * Kp;
        for (int ch = 0; ch < a.nchunks; ++ch)
            for (int b = 0; b < MB; ++b) {
                float* dst = b < MBq ? rec + ((size_t)ch * NCc + b * 128 + erow) * Kp : recO + ((size_t)ch * Cc + (b - MBq) * 128 + erow) * Kp;
#pragma unroll
                for (int u = 0; u < KS; ++u) {
                    const int gq = ehalf * KS + u;          // all Kp / 8 = 2 KS column groups, split over the two halves
                    float v[8];
                    tc5::tmem_ld8(tmem_W + lane_base + (uint32_t)((ch * MB + b) * Kp + gq * 8), v);
                    tc5::tmem_ld_wait();
                    *reinterpret_cast<float4*>(dst + gq * 8) = make_float4(v[0] * inv_gs, v[1] * inv_gs, v[2] * inv_gs, v[3] * inv_gs);
                    *reinterpret_cast<float4*>(dst + gq * 8 + 4) = make_float4(v[4] * inv_gs, v[5] * inv_gs, v[6] * inv_gs, v[7] * inv_gs);
                }
            }
        // ---- dbo / dbeta / dgamma: fixed-order reduction over rows (lanes, then warps) through shared memory
        float* red = reinterpret_cast<float*>(Gt);          // the token tile is free now: [8 warps][3][Kp]
        compute_sync();                                      // every warp is past its last tile reads
#pragma unroll
        for (int u = 0; u < GH; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float vb = acc_b[u][k], vg = acc_g[u][k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { vb += __shfl_xor_sync(0xffffffffu, vb, o); vg += __shfl_xor_sync(0xffffffffu, vg, o); }
                const int c = (ehalf * GH + u) * 8 + k;
                if (lane == 0 && c < Kp) { red[(warp * 3 + 1) * Kp + c] = vb; red[(warp * 3 + 2) * Kp + c] = vg; }
            }
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float v = acc_bo[nt][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0) red[(warp * 3) * Kp + 8 * nt + 2 * t + e] = v;
            }
        compute_sync();
        if (threadIdx.x < Kp) {
            const int c = threadIdx.x;
            float sbo = 0.f, sb = 0.f, sg = 0.f;
            const int hf = (c >> 3) / GH;                   // the epilogue half that owns column c
#pragma unroll
            for (int w = 0; w < RRB_CWARPS; ++w) {
                if (c < 8 * NTO) sbo += red[(w * 3) * Kp + c];
                if ((w >> 2) == hf && c < 8 * NTO) { sb += red[(w * 3 + 1) * Kp + c]; sg += red[(w * 3 + 2) * Kp + c]; }
            }
            recS[c] = sbo * inv_gs; recS[Kp + c] = sb * inv_gs; recS[2 * Kp + c] = sg * inv_gs;
        }
    }
    publish_amax_block(a.dx_amax, dx_max);
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem_base_s, (uint32_t)a.tmem_cols);
}

template <int KS, int NTO, bool VEC4, bool F16P, bool F16C, int UNR>
static int launch_attn_bwd_rr_v(const AttnBwdRRArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_bwd_rr<KS, NTO, VEC4, F16P, F16C, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 2048);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_bwd_rr)");
        attr_set = true;
    }
    if (launch_pdl(k_attn_bwd_rr<KS, NTO, VEC4, F16P, F16C, UNR>, dim3(grid), dim3(RRB_THREADS), (size_t)a.smem_bytes, st, a) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "k_attn_bwd_rr");
    RAT_CHECK_LAUNCH("k_attn_bwd_rr");
    return RAT_OK;
}
template <int KS, int NTO, bool VEC4>
static int launch_attn_bwd_rr(const AttnBwdRRArgs& a, int grid, cudaStream_t st) {
    static int variant = -1;     // RAT_RR_BWD_VARIANT (tuning aid): 1 = fp16 accumulators for everything but S and dP (see the forward)
    if (variant < 0) { const char* e = getenv("RAT_RR_BWD_VARIANT"); variant = e ? atoi(e) : 0; }
    if (variant == 1) return launch_attn_bwd_rr_v<KS, NTO, VEC4, true, true, 1>(a, grid, st);
    if (variant == 2) return launch_attn_bwd_rr_v<KS, NTO, VEC4, false, false, 2>(a, grid, st);
    return launch_attn_bwd_rr_v<KS, NTO, VEC4, false, false, 1>(a, grid, st);
}

static bool attn_bwd_rr_plan(int S, int D, int heads, int dh, AttnBwdRRArgs* a) {
    if (S < 1 || S > 16 || dh < 2 || dh > 16 || (dh & 1) || D < 2 || (D & 1) || D > 48) return false;
    const int Kp = pad16(D), KS = Kp / 16, KC1 = 2 * KS;
    a->D = D; a->H = heads; a->I = heads * dh; a->dh = dh;
    for (int hc = heads; hc >= 1; --hc) {
        if (heads % hc) continue;
        const int nch = heads / hc;
        const int NCc = round_up(3 * hc * dh, 128), Cc = round_up(hc * dh, 128);
        const int MB = (NCc + Cc) / 128;
        int cols = nch * MB * Kp + Kp, alloc = 32;
        while (alloc < cols) alloc <<= 1;
        if (alloc > 512) continue;
        const size_t smem = (size_t)4 * heads * KS * 32 * 16 + (size_t)nch * Kp * NCc * 2 +
                            (size_t)(2 * KC1 + (NCc + Cc) / 8) * tc5::TILE_CHUNK + (size_t)2 * Kp * 4 + 3 * 128 * 2 * 4 + 128 * 4 * 4 +
                            3 * 128 * 8 + (size_t)(NCc + Cc) * 4;
        if (smem > (size_t)max_smem_optin() - 4096) continue;
        if ((size_t)RRB_CWARPS * 3 * Kp * 4 > (size_t)((NCc + Cc) / 8) * tc5::TILE_CHUNK) continue;
        if ((size_t)16 * heads * dh * D > (size_t)(2 * KC1 + (NCc + Cc) / 8) * tc5::TILE_CHUNK) continue;     // raw weights staged in the tile region
        a->hc = hc; a->nchunks = nch; a->NCc = NCc; a->Cc = Cc; a->tmem_cols = alloc; a->smem_bytes = (int)smem;
        a->psize = nch * (NCc + Cc) * Kp + 3 * Kp;
        return true;
    }
    return false;
}
static int attn_bwd_rr_grid(int S, long long nseq) {
    const long long ntasks = S <= 8 ? (nseq + 1) / 2 : nseq;
    return (int)std::min<long long>((ntasks + RRB_CWARPS - 1) / RRB_CWARPS, (long long)num_sms());
}

}  // namespace rat

using namespace rat;

size_t attn_bwd_rr_workspace_bytes(int B, int T, int N, int D, int heads, int dh, int mode) {
    AttnBwdRRArgs a{};
    const int S = mode == 0 ? N : T;
    if (!attn_bwd_rr_plan(S, D, heads, dh, &a)) return 0;
    const long long nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    return (size_t)attn_bwd_rr_grid(S, nseq) * a.psize * sizeof(float);
}

// returns 1 when the shape is outside this kernel's envelope (the caller falls back to the tile kernels)
int attn_bwd_rr_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                         const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                         float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq,
                         int B, int T, int N, int D, int heads, int dh, float scale, float alpha, int mode,
                         const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, float out_drop_p,
                         unsigned long long seed, unsigned int rng_stream, cudaStream_t st) {
    AttnBwdRRArgs a{};
    const int S = mode == 0 ? N : T;
    if (!attn_bwd_rr_plan(S, D, heads, dh, &a)) return 1;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(base) |
          reinterpret_cast<uintptr_t>(dx)) & ((D % 4) == 0 ? 15 : 7)) != 0) return 1;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    const int grid = attn_bwd_rr_grid(S, a.nseq);
    if (!workspace || workspace_bytes < (size_t)grid * a.psize * sizeof(float)) return 1;
    reduce_ws_acquire(st, workspace);       // a deferred reduction may still be reading the records of an earlier call
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo;
    a.partials = workspace; a.dout_amax = dout_amax; a.dx_amax = dx_amax;
    a.g.S = S; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.scale = scale; a.alpha = alpha;
    a.out_drop_p = out_drop_p; a.seed = seed; a.rng_stream = rng_stream; a.rng_step = rng_step_ptr();
    const int NTO = (D + 7) / 8;
    const bool v4 = (D % 4) == 0;
    int rc;
#define RAT_RRB(KS_, NTO_) (v4 ? launch_attn_bwd_rr<KS_, NTO_, true>(a, grid, st) : launch_attn_bwd_rr<KS_, NTO_, false>(a, grid, st))
    switch (NTO) {
        case 1: rc = RAT_RRB(1, 1); break;
        case 2: rc = RAT_RRB(1, 2); break;
        case 3: rc = RAT_RRB(2, 3); break;
        case 4: rc = RAT_RRB(2, 4); break;
        case 5: rc = RAT_RRB(3, 5); break;
        case 6: rc = RAT_RRB(3, 6); break;
        default: return 1;
    }
#undef RAT_RRB
    if (rc != RAT_OK) return rc;
    AttnReduceTcArgs r{workspace, grid, a.psize, dWq, dWk, dWv, dWo, dbo, dln_w, dln_b, accumulate_wq, D, a.I, dh, a.hc,
                       a.nchunks, pad16(D), a.NCc, a.Cc, scale, 0.6931471805599453f};
    const int total = 4 * a.I * D + 3 * D;
    cudaStream_t rs = reduce_fork(st, workspace);
    k_reduce_attn_tc<<<std::max(1, std::min((total + 31) / 32, 1024)), dim3(32, 8), 0, rs>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_attn_tc");
    reduce_forked(rs, st, workspace);
    return RAT_OK;
}
