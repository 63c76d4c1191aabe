// K2 (Blackwell path, second generation): attention sub-block forward with every product on tcgen05.
//
//   out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )   (PreNorm RAT_m2.py:155-161, Attention :176-202, residual :224/:231)
//
// See encoder_tc2.cuh for the tile geometry and the thread organisation.  Per 128-row tile and per chunk of hc heads:
//   group g : q|k|v of its heads = LN(x)[128 x Kp] . Wqkv_g^T   (tcgen05 M=128, N = HPG*3*DHP)     -> TMEM -> fp16 tiles
//             per head: S = q k^T (2 x M=64,N=64) -> thread-per-row softmax -> block-diagonal P tile -> O = P v (2 x M=64,
//             N=DHP, K=64, v read MN-major) -> fp16 o tile
//   CTA     : y[128 x Np] (+)= o[128 x hc*DHP] . Wo_chunk^T     (tcgen05, accumulated over chunks in TMEM)
// epilogue: out = res + alpha * (y + bo), fp32.  The next tile's rows are staged while the out-projection runs.
#include "encoder_tc2.cuh"
#include <cstdlib>

namespace rat {

struct AttnTc2Args {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo; const float* bo;
    long long nseq;
    SeqGeom g;
    int D, H, dh, I;
    float qscale, alpha;        // qscale = softmax scale * log2(e), folded into the Wq rows
    int Kp, Np;                 // pad16(D): K of the q|k|v GEMM, N of the out-projection
    int hc, nchunks, HPG;       // heads per chunk, chunks, heads per group and chunk
    int NG;                     // HPG * 3 * DHP: q|k|v accumulator columns of one group
    int GR;                     // TMEM columns per group region (>= NG, >= 64 + DHP)
    int KO;                     // hc * DHP: K of the out-projection per chunk
    int SPT;                    // sequences per tile
    int o_alias;                // the o tile aliases the q region (single chunk)
    int smem_bytes;
    int off_wqkv, off_wo, off_f32, off_x, off_q, off_k, off_v, off_o, off_p;
    long long* dbg;             // RAT_T2_DBG=1: per-phase clock totals of CTA 0
};

// weight images shared by the forward and backward kernels (all threads of the CTA):
//   Wqkv image of chunk ch: rows n = [group][u][q|k|v][DHP] (head hl = group + 4u), K-major, Kp columns:
//       [n][c] = mul * W[n][c] * gamma[c]  (c < D),   [n][D] = mul * sum_c W[n][c] beta[c]   (the LayerNorm affine, folded)
__device__ __forceinline__ void t2_stage_wqkv(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                              const float* __restrict__ Wv, const float* __restrict__ ln_w,
                                              const float* __restrict__ ln_b, float qscale, int D, int dh, int hc, int nchunks,
                                              int NG, int Kp, unsigned char* __restrict__ img) {
    constexpr int DHP = 16;
    const int RI = 4 * NG, KC1 = Kp >> 3;
    for (int i = threadIdx.x; i < nchunks * RI * KC1; i += blockDim.x) {
        const int n = i % RI, rest = i / RI, kc = rest % KC1, ch = rest / KC1;
        const int gr = n / NG, rem = n - gr * NG;
        const int u = rem / (3 * DHP), rem2 = rem - u * 3 * DHP;
        const int w = rem2 / DHP, dd = rem2 - w * DHP;
        const int hl = gr + 4 * u;
        const bool live = hl < hc && dd < dh;
        const float* W = (w == 0 ? Wq : w == 1 ? Wk : Wv) + (size_t)((ch * hc + hl) * dh + dd) * D;
        const float mul = w == 0 ? qscale : 1.0f;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = kc * 8 + k;
            v[k] = 0.f;
            if (live && c < D) v[k] = mul * __ldg(W + c) * (ln_w ? __ldg(ln_w + c) : 1.0f);
            else if (live && c == D && ln_b) {
                float acc = 0.f;
                for (int c2 = 0; c2 < D; ++c2) acc = fmaf(__ldg(W + c2), __ldg(ln_b + c2), acc);
                v[k] = mul * acc;
            }
        }
        sts128(img + (size_t)ch * RI * Kp * 2 + tc5::kmajor_off(n, kc, RI), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
               pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
}

#define T2_S_MMA(hl_)                                                                                              \
    do {                                                                                                           \
        _Pragma("unroll") for (int h2 = 0; h2 < 2; ++h2) {                                                          \
            const uint32_t o_ = (uint32_t)((((hl_) * DC) * 128 + 64 * h2) * 16);                                   \
            tc5::mma_f16_w(t_S + ((uint32_t)(16 * h2) << 16), tc5::smem_desc(Qs + o_, tc5::TILE_CHUNK, 128u),       \
                           tc5::smem_desc(Ks + o_, tc5::TILE_CHUNK, 128u), idesc_s, 0u);                           \
        }                                                                                                          \
        tc5::mma_commit_w(&bar_s[grp]);                                                                            \
    } while (0)

template <int HPG, int SL, int ST, bool VEC4>
__global__ void __launch_bounds__(T2_THREADS, 1) k_attn_fwd_tc2(AttnTc2Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_x, bar_q[4], bar_s[4], bar_o[4], bar_og, bar_y;
    __shared__ uint32_t tmem_base_s;
    constexpr int DHP = 16, DC = DHP / 8;             // padded head width, 16-byte chunks per head
    constexpr int SLSH = SL == 16 ? 4 : 3;
    constexpr int NV = ST > 0 ? ST : SL;              // keys visited by the softmax loops
    constexpr int NG = HPG * 3 * DHP, RI = 4 * NG;    // q|k|v accumulator columns of one group, rows of a chunk's weight image
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int grp = warp >> 2, q = warp & 3, gt = threadIdx.x & 127;
    const int D = a.D, Kp = a.Kp, Np = a.Np, hc = a.hc, KO = a.KO, S = a.g.S;
    unsigned char* Wqkv_i = smem + a.off_wqkv;        // [nchunks][RI x Kp]
    unsigned char* Wo_i = smem + a.off_wo;            // [nchunks][Np x KO]
    float* bos = reinterpret_cast<float*>(smem + a.off_f32);     // [Np]
    unsigned char* Xt = smem + a.off_x;               // [128 x Kp]      normalised x | 1
    unsigned char* Qt = smem + a.off_q;               // [128 x hc*DHP]  q (scaled)
    unsigned char* Kt = smem + a.off_k;
    unsigned char* Vt = smem + a.off_v;
    unsigned char* Ot = smem + a.off_o;               // [128 x hc*DHP]  o   (== Qt when o_alias)
    unsigned char* Pg = smem + a.off_p + grp * 2 * T2_HALF_BYTES;   // this group's block-diagonal P tile (2 halves)
    const int nh = hc > grp ? (hc - grp + 3) / 4 : 0; // heads of this group in a chunk: hl = grp + 4u

    // ---- resident weight images, zero-initialised activation tiles
    t2_stage_wqkv(a.Wq, a.Wk, a.Wv, a.ln_w, a.ln_b, a.qscale, D, a.dh, hc, a.nchunks, NG, Kp, Wqkv_i);
    {
        const int KCo = KO >> 3;
        for (int i = threadIdx.x; i < a.nchunks * Np * KCo; i += blockDim.x) {
            const int n = i % Np, rest = i / Np, kc = rest % KCo, ch = rest / KCo;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = kc * 8 + k, hl = c / DHP, dd = c - hl * DHP;
                v[k] = (n < D && dd < a.dh) ? __ldg(a.Wo + (size_t)n * a.I + (ch * hc + hl) * a.dh + dd) : 0.f;
            }
            sts128(Wo_i + (size_t)ch * Np * KO * 2 + tc5::kmajor_off(n, kc, Np), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                   pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        for (int i = threadIdx.x; i < Np; i += blockDim.x) bos[i] = i < D ? a.bo[i] : 0.f;
        // x tile (pad columns stay zero) ... P tiles (off-diagonal blocks stay zero): everything from off_x on
        for (int i = threadIdx.x; i < (a.smem_bytes - a.off_x) / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem + a.off_x)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) {
        tc5::mbar_init(&bar_x, 4); tc5::mbar_init(&bar_og, 4); tc5::mbar_init(&bar_y, 1);
        for (int i = 0; i < 4; ++i) { tc5::mbar_init(&bar_q[i], 1); tc5::mbar_init(&bar_s[i], 1); tc5::mbar_init(&bar_o[i], 1); }
        tc5::fence_mbar_init();
    }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t t_reg = tmem_base_s + (uint32_t)(grp * a.GR);         // this group's column region
    const uint32_t t_S = t_reg, t_O = t_reg + 64;                        // inside the region once q|k|v is evacuated
    const uint32_t t_Y = tmem_base_s + (uint32_t)(4 * a.GR);
    const uint32_t idesc_q = tc5::instr_desc(TC_FMT, 128, NG);
    const uint32_t idesc_s = tc5::instr_desc(TC_FMT, 64, 64);
    const uint32_t idesc_pv = tc5::instr_desc(TC_FMT, 64, DHP, 0, 1);
    const uint32_t idesc_y = tc5::instr_desc(TC_FMT, 128, Np);
    const uint32_t Xs = tc5::smem_u32(Xt), Qs = tc5::smem_u32(Qt), Ks = tc5::smem_u32(Kt), Vs = tc5::smem_u32(Vt);
    const uint32_t Os = tc5::smem_u32(Ot), Ps = tc5::smem_u32(Pg), Wqs = tc5::smem_u32(Wqkv_i), Wos = tc5::smem_u32(Wo_i);
    // rows of this thread: M=128 accumulators (q|k|v, y): row_e ; M=64 accumulators (scores, o): row_s
    const int row_e = q * 32 + lane;
    const int hf = lane >> 4, li = lane & 15;
    const int row_s = 64 * hf + 16 * q + li;
    const int sb = (li >> 3) & 1;                                       // sub-slot inside the row group (SL == 8)
    unsigned char* const p_row = Pg + hf * T2_HALF_BYTES + poff(16 * q + li, 2 * q + (SL == 8 ? sb : 0));
    uint32_t ph_x = 0, ph_q = 0, ph_s = 0, ph_o = 0, ph_og = 0, ph_y = 0;
    const int nck = (D + 7) >> 3;
    __shared__ long long tks[8][16];
    if (threadIdx.x < 128) tks[threadIdx.x >> 4][threadIdx.x & 15] = 0;
    const bool prof = a.dbg != nullptr && (gt == 0 || gt == 127);
    long long t_prev = clock64();
#define T2_TICK(i) do { if (prof) { const long long t_now = clock64(); tks[grp * 2 + (gt ? 1 : 0)][i] += t_now - t_prev; t_prev = t_now; } } while (0)

    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    XRegs<VEC4> xr;
    if ((long long)blockIdx.x < ntiles) {
        t2_rows_load<VEC4, SLSH>(a.x, a.g, (long long)blockIdx.x * a.SPT, a.nseq, D, grp, gt, xr);
        t2_rows_finish<VEC4>(xr, D, Xt, grp, gt, nullptr);
        tc5::fence_proxy_async();
        group_sync(grp);
        if (q == 0) tc5::mbar_arrive_w(&bar_x);
    }
    // epilogue of the PREVIOUS tile (runs under the q|k|v MMA of the current one): out = res + alpha * (y + bo)
    long long gr_prev = -1;                          // global row of accumulator row row_e in the previous tile (-1: none)
    int it = 0;
    auto y_epilogue = [&](int it_prev) {
        tc5::mbar_wait_sleep(&bar_y, ph_y);
        ph_y ^= 1;
        tc5::fence_after_sync();
        for (int c = (grp + 4 - (it_prev & 3)) & 3; c < nck; c += 4) {
            float v[8], rv[8];
            tc5::tmem_ld8(t_Y + lane_base + c * 8, v);
            if (gr_prev >= 0 && a.res) load8<VEC4>(a.res + gr_prev * D, c * 8, D, rv);
            tc5::tmem_ld_wait();
            if (gr_prev >= 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = a.alpha * (v[k] + bos[c * 8 + k]);
                    if (a.res) v[k] += rv[k];
                }
                store8<VEC4>(a.out + gr_prev * D, c * 8, D, v);
            }
        }
        tc5::fence_before_sync();
    };
    bool y_pending = false;                           // an out-projection whose completion has not been waited for yet
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long s0 = tile * a.SPT;
        const bool has_next = tile + gridDim.x < ntiles;
        for (int ch = 0; ch < a.nchunks; ++ch) {
            // ---- q|k|v of this group's heads
            if (nh > 0 && q == 0) {
                if (ch == 0) tc5::mbar_wait_sleep(&bar_x, ph_x);
                tc5::fence_after_sync();
                const uint32_t wq = Wqs + (uint32_t)ch * RI * Kp * 2 + (uint32_t)(grp * NG * 16);
                for (int k = 0; k < Kp / 16; ++k)
                    tc5::mma_f16_w(t_reg, tc5::kdesc(Xs, 128, k), tc5::smem_desc(wq + (uint32_t)(k * 2 * RI * 16), (uint32_t)RI * 16, 128u),
                                   idesc_q, k > 0);
                tc5::mma_commit_w(&bar_q[grp]);
            }
            T2_TICK(0);
            if (ch == 0 && y_pending) { y_epilogue(it - 1); y_pending = false; }
            T2_TICK(1);
            if (nh > 0) {
                tc5::mbar_wait_sleep(&bar_q[grp], ph_q);
                tc5::fence_after_sync();
                T2_TICK(2);
                if (a.o_alias && y_pending) { tc5::mbar_wait_sleep(&bar_y, ph_y); ph_y ^= 1; y_pending = false; }
                // accumulator -> fp16 q | k | v tiles (thread = row row_e)
#pragma unroll
                for (int u = 0; u < HPG; ++u) {
                    float v[3][16];
#pragma unroll
                    for (int w = 0; w < 3; ++w) tc5::tmem_ld16(t_reg + lane_base + (u * 3 + w) * 16, v[w]);
                    tc5::tmem_ld_wait();
                    const uint32_t ro = tc5::toff(row_e, (grp + 4 * u) * DC);
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        unsigned char* dst = (w == 0 ? Qt : w == 1 ? Kt : Vt) + ro;
                        sts128(dst, pack_h2(v[w][0], v[w][1]), pack_h2(v[w][2], v[w][3]), pack_h2(v[w][4], v[w][5]), pack_h2(v[w][6], v[w][7]));
                        sts128(dst + tc5::TILE_CHUNK, pack_h2(v[w][8], v[w][9]), pack_h2(v[w][10], v[w][11]), pack_h2(v[w][12], v[w][13]),
                               pack_h2(v[w][14], v[w][15]));
                    }
                }
                tc5::fence_proxy_async();
                tc5::fence_before_sync();
                group_sync(grp);
                T2_TICK(3);
                if (q == 0) {                         // scores of the first head
                    tc5::fence_after_sync();
                    T2_S_MMA(grp);
                }
            }
            ph_q ^= 1;
            // ---- the next tile's rows start their trip from HBM now and are consumed after the head loop
            if (ch == a.nchunks - 1 && has_next) t2_rows_load<VEC4, SLSH>(a.x, a.g, (tile + gridDim.x) * a.SPT, a.nseq, D, grp, gt, xr);
            T2_TICK(4);
            const long long seq_s = s0 + (row_s >> SLSH);
            const bool valid_s = (row_s & (SL - 1)) < S && seq_s < a.nseq;
            for (int u = 0; u < nh; ++u) {
                const int hl = grp + 4 * u;
                // ---- softmax of this thread's row
                tc5::mbar_wait_sleep(&bar_s[grp], ph_s);
                ph_s ^= 1;
                tc5::fence_after_sync();
                T2_TICK(5);
                float v[16];
                tc5::tmem_ld16(t_S + lane_base + 16 * q, v);
                tc5::tmem_ld_wait();
                float xs[SL];
#pragma unroll
                for (int j = 0; j < SL; ++j) xs[j] = slot_pick<SL>(v, sb, j);
                if (ST == 0) {
#pragma unroll
                    for (int j = 0; j < SL; ++j) xs[j] = j < S ? xs[j] : -INFINITY;
                }
                float m = xs[0];
#pragma unroll
                for (int j = 1; j < NV; ++j) m = fmaxf(m, xs[j]);
                float l = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) { xs[j] = ex2f(xs[j] - m); l += xs[j]; }
                const float inv = valid_s ? rcp_fast(l) : 0.f;
                uint32_t pk[SL / 2];
#pragma unroll
                for (int j = 0; j < SL / 2; ++j)
                    pk[j] = (2 * j < NV) ? pack_h2(xs[2 * j] * inv, (2 * j + 1 < NV) ? xs[2 * j + 1] * inv : 0.f) : 0u;
                sts128(p_row, pk[0], pk[1], pk[2], pk[3]);
                if (SL == 16) sts128(p_row + 64 * 16, pk[SL / 2 - 4], pk[SL / 2 - 3], pk[SL / 2 - 2], pk[SL / 2 - 1]);
                tc5::fence_proxy_async();
                tc5::fence_before_sync();
                group_sync(grp);
                T2_TICK(6);
                if (q == 0) {
                    tc5::fence_after_sync();
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc5::mma_f16_w(t_O + ((uint32_t)(16 * h2) << 16),
                                           tc5::smem_desc(Ps + h2 * T2_HALF_BYTES + k * 2 * 64 * 16, 64 * 16, 128u),
                                           tc5::smem_desc(Vs + (uint32_t)((hl * DC * 128 + 64 * h2 + 16 * k) * 16), 128u, tc5::TILE_CHUNK),
                                           idesc_pv, k > 0);
                    tc5::mma_commit_w(&bar_o[grp]);
                    if (u + 1 < nh) T2_S_MMA(hl + 4);   // scores of the next head run under this head's o evacuation
                }
                T2_TICK(7);
                // ---- o of this head -> fp16 o tile
                tc5::mbar_wait_sleep(&bar_o[grp], ph_o);
                ph_o ^= 1;
                tc5::fence_after_sync();
                T2_TICK(8);
                if (!a.o_alias && y_pending && u == 0) { tc5::mbar_wait_sleep(&bar_y, ph_y); ph_y ^= 1; y_pending = false; }
                {
                    float o[16];
                    tc5::tmem_ld16(t_O + lane_base, o);
                    tc5::tmem_ld_wait();
                    unsigned char* dst = Ot + tc5::toff(row_s, hl * DC);
                    sts128(dst, pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
                    sts128(dst + tc5::TILE_CHUNK, pack_h2(o[8], o[9]), pack_h2(o[10], o[11]), pack_h2(o[12], o[13]), pack_h2(o[14], o[15]));
                }
            }
            if (nh == 0 && y_pending && ch > 0) { tc5::mbar_wait_sleep(&bar_y, ph_y); ph_y ^= 1; y_pending = false; }
            tc5::fence_proxy_async();
            tc5::fence_before_sync();
            group_sync(grp);
            T2_TICK(9);
            // ---- out-projection of this chunk: issued by one group's warp 0 once all four groups have arrived
            if (q == 0) {
                tc5::mbar_arrive_w(&bar_og);
                if (grp == ((it + ch) & 3)) {
                    tc5::mbar_wait_sleep(&bar_og, ph_og);
                    tc5::fence_after_sync();
                    const uint32_t wo = Wos + (uint32_t)ch * Np * KO * 2;
                    for (int k = 0; k < KO / 16; ++k)
                        tc5::mma_f16_w(t_Y, tc5::kdesc(Os, 128, k), tc5::kdesc(wo, Np, k), idesc_y, (ch > 0 || k > 0) ? 1u : 0u);
                    tc5::mma_commit_w(&bar_y);
                }
            }
            ph_og ^= 1;
            y_pending = true;
            T2_TICK(10);
        }
        // ---- finish staging the next tile (the x tile is free once every group's q|k|v MMAs are done)
        if (has_next) {
            for (int g2 = 0; g2 < 4; ++g2)
                if (hc > g2 && g2 != grp) tc5::mbar_wait_sleep(&bar_q[g2], ph_q ^ 1);
            T2_TICK(11);
            t2_rows_finish<VEC4>(xr, D, Xt, grp, gt, nullptr);
            tc5::fence_proxy_async();
            group_sync(grp);
            if (q == 0) tc5::mbar_arrive_w(&bar_x);
        }
        ph_x ^= 1;
        T2_TICK(12);
        {   // global row of this thread's accumulator row, for the deferred epilogue
            const int slot = row_e >> SLSH, pos = row_e & (SL - 1);
            const long long seq = s0 + slot;
            gr_prev = (pos < S && seq < a.nseq) ? a.g.grow(seq, pos) : -1;
        }
    }
    if (y_pending) y_epilogue(it - 1);
    __syncthreads();
    if (a.dbg && blockIdx.x == 0 && threadIdx.x < 128) a.dbg[threadIdx.x] = (threadIdx.x & 15) == 15 ? it : tks[threadIdx.x >> 4][threadIdx.x & 15];
    if (warp == 0) tc5::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace rat

using namespace rat;

template <int HPG, int SL, int ST, bool VEC4>
static int launch_attn_fwd_tc2(const AttnTc2Args& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd_tc2<HPG, SL, ST, VEC4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 2048);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd_tc2)");
        attr_set = true;
    }
    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms());
    k_attn_fwd_tc2<HPG, SL, ST, VEC4><<<grid, T2_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd_tc2");
    return RAT_OK;
}

template <int HPG, bool VEC4>
static int launch_attn_fwd_tc2_s(const AttnTc2Args& a, cudaStream_t st) {
    const int S = a.g.S;
    if (S > 8) return S == 14 ? launch_attn_fwd_tc2<HPG, 16, 14, VEC4>(a, st) : launch_attn_fwd_tc2<HPG, 16, 0, VEC4>(a, st);
    return S == 6 ? launch_attn_fwd_tc2<HPG, 8, 6, VEC4>(a, st) : launch_attn_fwd_tc2<HPG, 8, 0, VEC4>(a, st);
}

// Shared planning of the second-generation attention kernels: head chunking, TMEM regions, shared-memory map.
// extra_tile_bytes: per-kernel additions after the common tiles (backward).  Returns false if the shape is not covered.
bool attn_tc2_plan(int S, int D, int heads, int dh, AttnTc2Args* a) {
    if (S > 16 || S < 1 || dh > 16 || dh < 2 || (dh & 1) || D < 2 || (D & 1) || D > 64) return false;
    if ((D % 4) != 0 && D > 32) return false;
    const int DHP = 16;
    a->D = D; a->H = heads; a->dh = dh; a->I = heads * dh;
    a->Kp = pad16(D + 1); a->Np = pad16(D);          // + 1: the column of ones that carries the LayerNorm beta
    const int SL = S > 8 ? 16 : 8;
    a->SPT = 128 / SL;
    for (int hc = 8; hc >= 1; hc >>= 1) {
        if (heads % hc) continue;
        const int HPG = (hc + 3) / 4, NG = HPG * 3 * DHP, GR = std::max(NG, 64 + DHP);
        if (4 * GR + a->Np > 512) continue;
        const int nchunks = heads / hc, KO = hc * DHP;
        const int alias = nchunks == 1;
        size_t off = 0;
        const int off_wqkv = (int)off; off += (size_t)nchunks * 4 * NG * a->Kp * 2;
        const int off_wo = (int)off; off += (size_t)nchunks * a->Np * KO * 2;
        const int off_f32 = (int)off; off += (size_t)a->Np * 4;
        off = (off + 127) & ~(size_t)127;
        const int off_x = (int)off; off += (size_t)128 * a->Kp * 2;
        const int off_q = (int)off; off += (size_t)128 * KO * 2;
        const int off_k = (int)off; off += (size_t)128 * KO * 2;
        const int off_v = (int)off; off += (size_t)128 * KO * 2;
        int off_o = off_q;
        if (!alias) { off_o = (int)off; off += (size_t)128 * KO * 2; }
        const int off_p = (int)off; off += (size_t)4 * 2 * T2_HALF_BYTES;
        if (off > (size_t)max_smem_optin() - 2048) continue;
        a->hc = hc; a->nchunks = nchunks; a->HPG = HPG; a->NG = NG; a->GR = GR; a->KO = KO; a->o_alias = alias;
        a->off_wqkv = off_wqkv; a->off_wo = off_wo; a->off_f32 = off_f32; a->off_x = off_x; a->off_q = off_q; a->off_k = off_k;
        a->off_v = off_v; a->off_o = off_o; a->off_p = off_p; a->smem_bytes = (int)off;
        return true;
    }
    return false;
}

// returns RAT_OK if launched, 1 if the shape is not covered (caller falls back to the first-generation kernel), <0 on error
int attn_fwd_tc2_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                          const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                          int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st) {
    const int S = mode == 0 ? N : T;
    AttnTc2Args a{};
    if (!attn_tc2_plan(S, D, heads, dh, &a)) return 1;
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo; a.bo = bo;
    a.g.S = S; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.qscale = scale * 1.4426950408889634f; a.alpha = alpha;
    static long long* dbg = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) { const char* e = getenv("RAT_T2_DBG"); dbg_on = (e && e[0] == '1') ? 1 : 0; if (dbg_on) cudaMalloc(&dbg, 128 * 8); }
    a.dbg = dbg_on ? dbg : nullptr;
    int rc;
    if (a.HPG == 2) rc = (D % 4) == 0 ? launch_attn_fwd_tc2_s<2, true>(a, st) : launch_attn_fwd_tc2_s<2, false>(a, st);
    else rc = (D % 4) == 0 ? launch_attn_fwd_tc2_s<1, true>(a, st) : launch_attn_fwd_tc2_s<1, false>(a, st);
    if (dbg_on && rc == RAT_OK) {
        long long h[128];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        static const char* names[13] = {"qkv issue(+wait x)", "y epilogue(prev)", "wait qkv MMA", "qkv evac+sync", "x prefetch issue", "wait S MMA",
                                        "softmax+P+sync", "PV/S issue", "wait PV MMA", "o evac+sync", "outproj issue", "peek qkv bars", "stage finish"};
        fprintf(stderr, "[t2 fwd dbg] S=%d mode=%d tiles(CTA0)=%lld ; cycles per tile, group: thread0 / thread127\n", S, mode, h[15]);
        for (int i = 0; i < 13; ++i) {
            fprintf(stderr, "  %-20s", names[i]);
            for (int g = 0; g < 4; ++g) fprintf(stderr, "  g%d %6.0f /%6.0f", g, (double)h[(g * 2) * 16 + i] / h[15], (double)h[(g * 2 + 1) * 16 + i] / h[15]);
            fprintf(stderr, "\n");
        }
    }
    return rc;
}
