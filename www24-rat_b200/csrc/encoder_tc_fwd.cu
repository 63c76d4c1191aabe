// K2 (Blackwell path): fused RAT-block forward kernels on the 5th-generation tensor cores.
//
//   k_ff_fwd_tc : out = res + W2 gelu(W1 [LN](x) + b1) + b2          (reference: FeedForward RAT_m2.py:163-174)
//
// Structure (precision mode "bf16"): one persistent CTA per SM, 512 threads = 2 independent TEAMS of 8 warps.  A team
// owns one 128-token tile at a time: its threads LayerNorm / convert the tile to bf16 and write it to shared memory in
// the UMMA canonical K-major layout (tc5.cuh); ONE thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM,
// M=128) against the resident weight image; the accumulator comes back through tcgen05.ld (thread = token row) for
// the bias/GELU/residual epilogues.  While one team waits for its MMAs the other team runs its SIMT phases, so the
// tensor pipe, the LSU and the FMA pipe overlap without warp specialisation.  The residual stream, LayerNorm
// statistics, GELU and all accumulation are fp32; only the MMA operands are rounded to bf16.
#include "encoder_tc.cuh"

namespace rat {

// ------------------------------------------------------------------------------------------------ FeedForward
struct FFTcArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* W1; const float* b1; const float* W2; const float* b2;
    long long rows;
    int D, M;
    int Kp;      // pad16(D): K of GEMM1, KC1 = Kp/8
    int Mp;      // pad16(M): N of GEMM1 = K of GEMM2
    int Np;      // pad16(D): N of GEMM2
    int smem_bytes;
};

template <int KCH, bool VEC4>
__global__ void __launch_bounds__(TC_THREADS, 1) k_ff_fwd_tc(FFTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = a.D, M = a.M, Kp = a.Kp, Mp = a.Mp, Np = a.Np;
    const int KC1 = Kp >> 3, KC2 = Mp >> 3;
    unsigned char* W1i = smem_raw;                                   // [Mp x Kp] bf16 image
    unsigned char* W2i = W1i + (size_t)Mp * Kp * 2;                  // [Np x Mp] bf16 image
    float* b1s = reinterpret_cast<float*>(W2i + (size_t)Np * Mp * 2);   // [Mp]
    float* b2s = b1s + Mp;                                           // [Np]
    unsigned char* team_base = reinterpret_cast<unsigned char*>(b2s + Np);
    const size_t team_bytes = (size_t)TILE_M * Kp * 2 + (size_t)TILE_M * Mp * 2;
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int team = threadIdx.x / TEAM_THREADS, tid2 = threadIdx.x % TEAM_THREADS;
    const int warp2 = tid2 >> 5, lane = tid2 & 31;
    unsigned char* At = team_base + team * team_bytes;               // [128 x Kp] bf16
    unsigned char* Ht = At + (size_t)TILE_M * Kp * 2;                // [128 x Mp] bf16

    stage_weight_image(a.W1, M, D, Mp, Kp, W1i);
    stage_weight_image(a.W2, D, M, Np, Mp, W2i);
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) b1s[i] = i < M ? a.b1[i] : 0.f;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) b2s[i] = i < D ? a.b2[i] : 0.f;
    if (threadIdx.x == 0) { tc5::mbar_init(&mbar[0], 1); tc5::mbar_init(&mbar[1], 1); tc5::fence_mbar_init(); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&tmem_base_s, 256);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_H = tmem_base_s + team * 128;                // columns [0, Mp)
    const uint32_t tmem_Y = tmem_H + Mp;                             // columns [Mp, Mp+Np)
    const uint32_t idesc1 = tc5::instr_desc(TC_FMT, TILE_M, Mp);
    const uint32_t idesc2 = tc5::instr_desc(TC_FMT, TILE_M, Np);
    const uint32_t lane_base = (uint32_t)((warp2 & 3) * 32) << 16;   // this warp's TMEM lane quadrant
    const int chalf = warp2 >> 2;                                    // column half handled by this warp
    const int row_e = (warp2 & 3) * 32 + lane;                       // accumulator row of this thread
    uint32_t phase = 0;
    uint64_t* bar = &mbar[team];

    const long long ntiles = (a.rows + TILE_M - 1) / TILE_M;
    for (long long tile = (long long)blockIdx.x * 2 + team; tile < ntiles; tile += (long long)gridDim.x * 2) {
        const long long r0 = tile * TILE_M;
        const int R = (int)min((long long)TILE_M, a.rows - r0);
        // ---- phase 1: x tile -> (LayerNorm) -> bf16 A tile
        {
            const int row = tid2 >> 1, h = tid2 & 1;
            stage_row_h<KCH, VEC4>(a.x + (r0 + row) * D, row < R, D, KC1, row, h, a.ln_w, a.ln_b, At);
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        // ---- GEMM1: H[128 x Mp] = A[128 x Kp] . W1^T
        if (tid2 == 0) {
            tc5::fence_after_sync();
            const uint32_t a0 = tc5::smem_u32(At), b0 = tc5::smem_u32(W1i);
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_H, tc5::kdesc(a0, TILE_M, k), tc5::kdesc(b0, Mp, k),
                             idesc1, k > 0);
            tc5::mma_commit(bar);
        }
        tc5::mbar_wait(bar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 1: h = gelu(H + b1) -> bf16 H tile (K-major operand of GEMM2)
        {
            const int ng = Mp >> 3, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
            for (int g = g0; g < g1; ++g) {
                float v[8];
                tc5::tmem_ld8(tmem_H + lane_base + g * 8, v);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; ++k) { float dg; gelu_fast(v[k] + b1s[g * 8 + k], v[k], dg); }
                // columns >= M: W1 image rows are zero and b1s is zero -> gelu(0) = 0
                sts128(Ht + tc5::toff(row_e, g), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                       pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            }
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        // ---- GEMM2: Y[128 x Np] = h[128 x Mp] . W2^T
        if (tid2 == 0) {
            tc5::fence_after_sync();
            const uint32_t a0 = tc5::smem_u32(Ht), b0 = tc5::smem_u32(W2i);
            for (int k = 0; k < Mp / 16; ++k)
                tc5::mma_f16(tmem_Y, tc5::kdesc(a0, TILE_M, k), tc5::kdesc(b0, Np, k),
                             idesc2, k > 0);
            tc5::mma_commit(bar);
        }
        tc5::mbar_wait(bar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 2: out = res + Y + b2
        {
            const int ng = Np >> 3, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
            for (int g = g0; g < g1; ++g) {
                if (g * 8 >= D) break;
                float v[8], rv[8];
                tc5::tmem_ld8(tmem_Y + lane_base + g * 8, v);
                if (row_e < R) load8<VEC4>(a.res + (r0 + row_e) * D, g * 8, D, rv);
                tc5::tmem_ld_wait();
                if (row_e < R) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += rv[k] + b2s[g * 8 + k];
                    store8<VEC4>(a.out + (r0 + row_e) * D, g * 8, D, v);
                }
            }
        }
        tc5::fence_before_sync();          // TMEM reads of this tile are ordered before the next tile's MMAs
    }
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tmem_base_s, 256);
}

// ------------------------------------------------------------------------------------------------ Attention
// out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )    (PreNorm RAT_m2.py:155-161, Attention :176-202, residual :224/:231)
//
// Per 128-row tile (whole sequences: SPT = 128 / S of them) and per chunk of hc heads:
//   GEMM  q|k|v[128 x 3*hc*DHP] = LN(x)[128 x Kp] . Wqkv_chunk^T     tcgen05, head width zero-padded to DHP = pad16(dh),
//                                                                    softmax scale * log2(e) folded into the Wq rows
//   TMEM -> bf16 q|k|v tile in shared memory (canonical core-matrix layout: also what ldmatrix wants)
//   softmax(q k^T) v per (sequence, head) on ONE WARP: ldmatrix + mma.sync.m16n8k16 bf16 (two sequences of S <= 8
//     share the m16 tile, off-diagonal blocks masked), fp32 scores / exp2 / sums, P rounded to bf16 for P.V
//   GEMM  y[128 x Np] (+)= o_chunk[128 x Cp] . Wo_chunk^T            tcgen05, accumulated over head chunks in TMEM
// epilogue: out = res + alpha * (y + bo), fp32.
struct AttnTcArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo; const float* bo;
    long long nseq;
    SeqGeom g;
    int D, H, I;
    float qscale, alpha;     // qscale = softmax scale * log2(e)
    int Kp;                  // pad16(D)
    int Np;                  // pad16(D): N of the out-projection
    int hc, nchunks;         // heads per chunk
    int NCq;                 // 3 * hc * DHP: N of the q|k|v GEMM
    int Cp;                  // pad16(hc * dh): K of the out-projection per chunk
    int SPT;                 // sequences per tile
    int smem_bytes;
};

template <int DH>
__device__ __forceinline__ void attn_core_bf16(const unsigned char* __restrict__ QKVt, int hc, unsigned char* __restrict__ Ot,
                                               int nseq_t, int S, int warp, int nwarps, int lane, const CoreLane& cl) {
    constexpr int DHP = (DH + 15) / 16 * 16, KS = DHP / 16, ND = (DH + 7) / 8;
    constexpr uint32_t HEAD = (DHP / 8) * tc5::TILE_CHUNK;            // byte stride between heads inside q / k / v
    const int t = lane & 3;
    const int ntasks = (cl.packed ? (nseq_t + 1) >> 1 : nseq_t) * hc;
    const uint32_t qkv_s = tc5::smem_u32(QKVt);
    const uint32_t part = (uint32_t)hc * HEAD;                        // q -> k -> v
    int sp = warp / hc, hl = warp - sp * hc;
    const int dsp = nwarps / hc, dhl = nwarps - dsp * hc;
    for (int task = warp; task < ntasks; task += nwarps) {
        const int seq0 = cl.packed ? 2 * sp : sp;
        const uint32_t tb = (uint32_t)(seq0 * S) * 16u;
        const uint32_t qa = qkv_s + tb + (uint32_t)hl * HEAD;
        float sc[2][4] = {};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a[4], b[4];
            ldsm_x4(a, qa + cl.a_off + 2 * ks * tc5::TILE_CHUNK);
            ldsm_x4(b, qa + part + cl.b_off + 2 * ks * tc5::TILE_CHUNK);
            mma_h_16x8x16(sc[0], a, b[0], b[1]);
            mma_h_16x8x16(sc[1], a, b[2], b[3]);
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
        if (mlo == -INFINITY) mlo = 0.f;                            // rows that do not exist: exp2(-inf) = 0, no NaN
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
        uint32_t pa[4];
        pa[0] = pack_h2(sc[0][0], sc[0][1]); pa[1] = pack_h2(sc[0][2], sc[0][3]);
        pa[2] = pack_h2(sc[1][0], sc[1][1]); pa[3] = pack_h2(sc[1][2], sc[1][3]);
        float o[2 * KS][4] = {};
#pragma unroll
        for (int pp = 0; pp < KS; ++pp) {
            uint32_t vb[4];
            ldsm_x4_t(vb, qa + 2 * part + cl.a_off + 2 * pp * tc5::TILE_CHUNK);
            mma_h_16x8x16(o[2 * pp], pa, vb[0], vb[1]);
            if (2 * pp + 1 < ND) mma_h_16x8x16(o[2 * pp + 1], pa, vb[2], vb[3]);
        }
        const float ilo = rcp_fast(llo), ihi = rcp_fast(lhi);
        unsigned char* olo = Ot + tb + (uint32_t)(cl.lo_rel * 16);
        unsigned char* ohi = Ot + tb + (uint32_t)(cl.hi_rel * 16);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
            const int d = 8 * nd + 2 * t;
            if (d < DH) {                                           // DH even: the (d, d+1) pair is valid as a whole
                const int col = hl * DH + d;
                const uint32_t co = (uint32_t)(col >> 3) * tc5::TILE_CHUNK + (uint32_t)(col & 7) * 2u;
                if (vlo) *reinterpret_cast<uint32_t*>(olo + co) = pack_h2(o[nd][0] * ilo, o[nd][1] * ilo);
                if (vhi) *reinterpret_cast<uint32_t*>(ohi + co) = pack_h2(o[nd][2] * ihi, o[nd][3] * ihi);
            }
        }
        sp += dsp; hl += dhl;
        if (hl >= hc) { hl -= hc; ++sp; }
    }
}

// Long sequences (16 < S <= 128: cross attention at K = 16..64 retrieved neighbours, RAT_m0's flat T*N sequence): a warp
// task is (sequence, head, block of 16 query rows).  All NKB = ceil(S/16) score blocks of the 16 rows stay in registers
// (fp32), the softmax statistics are quad reductions as in the short core, P is rounded to fp16 block by block for
// P.V.  ldmatrix rows beyond the sequence are clamped to its last row (always finite data); their scores are masked.
template <int DH, int NKB>
__device__ __forceinline__ void attn_core_long(const unsigned char* __restrict__ QKVt, int hc, unsigned char* __restrict__ Ot,
                                               int nseq_t, int S, int warp, int nwarps, int lane) {
    constexpr int DHP = (DH + 15) / 16 * 16, KS = DHP / 16, ND = (DH + 7) / 8;
    constexpr uint32_t HEAD = (DHP / 8) * tc5::TILE_CHUNK;
    const int g = lane >> 2, t = lane & 3;
    const int nrb = (S + 15) >> 4;
    const int ntasks = nseq_t * hc * nrb;
    const uint32_t qkv_s = tc5::smem_u32(QKVt);
    const uint32_t part = (uint32_t)hc * HEAD;
    const uint32_t a_chunk = (uint32_t)(lane >> 4) * tc5::TILE_CHUNK, b_chunk = (uint32_t)((lane >> 3) & 1) * tc5::TILE_CHUNK;
    const int a_row = lane & 15, b_row = (lane & 7) + (lane >> 4) * 8;
    for (int task = warp; task < ntasks; task += nwarps) {
        const int rb = task % nrb, sh = task / nrb;
        const int hl = sh % hc, sq = sh / hc;
        const uint32_t tb = (uint32_t)(sq * S) * 16u;
        const uint32_t qa = qkv_s + tb + (uint32_t)hl * HEAD;
        const uint32_t q_off = (uint32_t)min(rb * 16 + a_row, S - 1) * 16u + a_chunk;
        float sc[NKB][2][4];
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) sc[kb][nt][0] = sc[kb][nt][1] = sc[kb][nt][2] = sc[kb][nt][3] = 0.f;
            if (kb < nrb) {
                const uint32_t k_off = (uint32_t)min(kb * 16 + b_row, S - 1) * 16u + b_chunk;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t a[4], b[4];
                    ldsm_x4(a, qa + q_off + 2 * ks * tc5::TILE_CHUNK);
                    ldsm_x4(b, qa + part + k_off + 2 * ks * tc5::TILE_CHUNK);
                    mma_h_16x8x16(sc[kb][0], a, b[0], b[1]);
                    mma_h_16x8x16(sc[kb][1], a, b[2], b[3]);
                }
            }
        }
        float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = kb * 16 + 8 * nt + 2 * t + (e & 1);
                    if (j >= S) sc[kb][nt][e] = -INFINITY;
                    if (e < 2) mlo = fmaxf(mlo, sc[kb][nt][e]); else mhi = fmaxf(mhi, sc[kb][nt][e]);
                }
        mlo = qmax(mlo); mhi = qmax(mhi);
        float llo = 0.f, lhi = 0.f;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sc[kb][nt][e] = ex2f(sc[kb][nt][e] - ((e < 2) ? mlo : mhi));
                    if (e < 2) llo += sc[kb][nt][e]; else lhi += sc[kb][nt][e];
                }
        llo = qsum(llo); lhi = qsum(lhi);
        float o[2 * KS][4] = {};
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
            if (kb >= nrb) continue;
            uint32_t pa[4];
            pa[0] = pack_h2(sc[kb][0][0], sc[kb][0][1]); pa[1] = pack_h2(sc[kb][0][2], sc[kb][0][3]);
            pa[2] = pack_h2(sc[kb][1][0], sc[kb][1][1]); pa[3] = pack_h2(sc[kb][1][2], sc[kb][1][3]);
            const uint32_t v_off = (uint32_t)min(kb * 16 + a_row, S - 1) * 16u + a_chunk;
#pragma unroll
            for (int pp = 0; pp < KS; ++pp) {
                uint32_t vb[4];
                ldsm_x4_t(vb, qa + 2 * part + v_off + 2 * pp * tc5::TILE_CHUNK);
                mma_h_16x8x16(o[2 * pp], pa, vb[0], vb[1]);
                if (2 * pp + 1 < ND) mma_h_16x8x16(o[2 * pp + 1], pa, vb[2], vb[3]);
            }
        }
        const float ilo = rcp_fast(llo), ihi = rcp_fast(lhi);
        const int rlo = rb * 16 + g, rhi = rlo + 8;
        unsigned char* olo = Ot + tb + (uint32_t)(rlo * 16);
        unsigned char* ohi = Ot + tb + (uint32_t)(rhi * 16);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
            const int d = 8 * nd + 2 * t;
            if (d < DH) {
                const int col = hl * DH + d;
                const uint32_t co = (uint32_t)(col >> 3) * tc5::TILE_CHUNK + (uint32_t)(col & 7) * 2u;
                if (rlo < S) *reinterpret_cast<uint32_t*>(olo + co) = pack_h2(o[nd][0] * ilo, o[nd][1] * ilo);
                if (rhi < S) *reinterpret_cast<uint32_t*>(ohi + co) = pack_h2(o[nd][2] * ihi, o[nd][3] * ihi);
            }
        }
    }
}
template <int DH>
__device__ __forceinline__ void attn_core_long_any(const unsigned char* __restrict__ QKVt, int hc, unsigned char* __restrict__ Ot,
                                                   int nseq_t, int S, int warp, int nwarps, int lane) {
    if (S <= 48) attn_core_long<DH, 3>(QKVt, hc, Ot, nseq_t, S, warp, nwarps, lane);
    else if (S <= 80) attn_core_long<DH, 5>(QKVt, hc, Ot, nseq_t, S, warp, nwarps, lane);
    else attn_core_long<DH, 8>(QKVt, hc, Ot, nseq_t, S, warp, nwarps, lane);
}

// LONG = false: sequences of <= 16 tokens (short core, two sequences packed per m16 tile when S <= 8);
// LONG = true : 16 < S <= 128 (long core).  Separate instantiations keep the short kernel's register allocation intact.
template <int DH, int KCH, bool VEC4, bool LONG>
__global__ void __launch_bounds__(TC_THREADS, 1) k_attn_fwd_tc(AttnTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int DHP = (DH + 15) / 16 * 16;
    const int D = a.D, Kp = a.Kp, Np = a.Np, hc = a.hc, NCq = a.NCq, Cp = a.Cp, S = a.g.S;
    const int KC1 = Kp >> 3, KCq = NCq >> 3, KCo = Cp >> 3;
    unsigned char* Wqkv_i = smem_raw;                                        // [nchunks*NCq x Kp] bf16
    unsigned char* Wo_i = Wqkv_i + (size_t)a.nchunks * NCq * Kp * 2;         // [nchunks][Np x Cp] bf16
    float* bos = reinterpret_cast<float*>(Wo_i + (size_t)a.nchunks * Np * Cp * 2);   // [Np]
    unsigned char* team_base = reinterpret_cast<unsigned char*>(bos + Np);
    const size_t team_bytes = (size_t)TILE_M * (Kp + NCq + Cp) * 2;
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int team = threadIdx.x / TEAM_THREADS, tid2 = threadIdx.x % TEAM_THREADS;
    const int warp2 = tid2 >> 5, lane = tid2 & 31;
    unsigned char* At = team_base + team * team_bytes;                       // [128 x Kp]
    unsigned char* QKVt = At + (size_t)TILE_M * Kp * 2;                      // [128 x NCq]
    unsigned char* Ot = QKVt + (size_t)TILE_M * NCq * 2;                     // [128 x Cp]

    // ---- resident weight images
    {
        const int rows_img = a.nchunks * NCq, total = rows_img * KC1;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int n = i % rows_img, kc = i / rows_img;
            const int ch = n / NCq, rem = n - ch * NCq;
            const int w = rem / (hc * DHP), rem2 = rem - w * (hc * DHP);
            const int hl = rem2 / DHP, d = rem2 - hl * DHP;
            const float* W = w == 0 ? a.Wq : w == 1 ? a.Wk : a.Wv;
            const float mul = w == 0 ? a.qscale : 1.0f;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = kc * 8 + k;
                v[k] = (d < DH && c < D) ? mul * __ldg(W + (size_t)((ch * hc + hl) * DH + d) * D + c) : 0.f;
            }
            sts128(Wqkv_i + (size_t)ch * NCq * Kp * 2 + tc5::kmajor_off(rem, kc, NCq), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                   pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        const int per = Np * KCo;
        for (int i = threadIdx.x; i < a.nchunks * per; i += blockDim.x) {
            const int ch = i / per, rem = i - ch * per;
            const int n = rem % Np, kc = rem / Np;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = kc * 8 + k;
                v[k] = (n < D && c < hc * DH) ? __ldg(a.Wo + (size_t)n * a.I + ch * hc * DH + c) : 0.f;
            }
            sts128(Wo_i + (size_t)ch * Np * Cp * 2 + tc5::kmajor_off(n, kc, Np), pack_h2(v[0], v[1]),
                   pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        for (int i = threadIdx.x; i < Np; i += blockDim.x) bos[i] = i < D ? a.bo[i] : 0.f;
        // o tiles: pad columns [hc*dh, Cp) are never written by the attention core and must read as zero
        for (int i = threadIdx.x; i < (int)(2 * team_bytes / 16); i += blockDim.x)
            reinterpret_cast<uint4*>(team_base)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) { tc5::mbar_init(&mbar[0], 1); tc5::mbar_init(&mbar[1], 1); tc5::fence_mbar_init(); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_Q = tmem_base_s + team * 256;                // columns [0, NCq)
    const uint32_t tmem_Y = tmem_Q + NCq;                            // columns [NCq, NCq + Np)
    const uint32_t idesc_q = tc5::instr_desc(TC_FMT, TILE_M, NCq);
    const uint32_t idesc_o = tc5::instr_desc(TC_FMT, TILE_M, Np);
    const uint32_t lane_base = (uint32_t)((warp2 & 3) * 32) << 16;
    const int chalf = warp2 >> 2;
    const int row_e = (warp2 & 3) * 32 + lane;
    uint32_t phase = 0;
    uint64_t* bar = &mbar[team];
    const CoreLane cl = make_core_lane(S, lane);
    const uint32_t At_s = tc5::smem_u32(At), Ot_s = tc5::smem_u32(Ot), Wq_s = tc5::smem_u32(Wqkv_i), Wo_s = tc5::smem_u32(Wo_i);

    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    for (long long tile = (long long)blockIdx.x * 2 + team; tile < ntiles; tile += (long long)gridDim.x * 2) {
        const long long s0 = tile * a.SPT;
        const int nseq_t = (int)min((long long)a.SPT, a.nseq - s0);
        const int R = nseq_t * S;
        // ---- phase 1: LN(x) tile -> bf16 A tile
        {
            const int row = tid2 >> 1, h = tid2 & 1;
            const bool valid = row < R;
            const int ls = row / S, pos = row - ls * S;
            const long long gr = valid ? a.g.grow(s0 + ls, pos) : 0;
            stage_row_h<KCH, VEC4>(a.x + gr * D, valid, D, KC1, row, h, a.ln_w, a.ln_b, At);
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        if (tid2 == 0) {
            tc5::fence_after_sync();
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_Q, tc5::kdesc(At_s, TILE_M, k), tc5::kdesc(Wq_s, NCq, k), idesc_q, k > 0);
            tc5::mma_commit(bar);
        }
        for (int ch = 0; ch < a.nchunks; ++ch) {
            tc5::mbar_wait(bar, phase);
            phase ^= 1;
            tc5::fence_after_sync();
            // ---- epilogue 1: q|k|v accumulator -> bf16 tile
            {
                const int ng = NCq >> 4, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
                for (int gq = g0; gq < g1; ++gq) {
                    float v[16];
                    tc5::tmem_ld16(tmem_Q + lane_base + gq * 16, v);
                    tc5::tmem_ld_wait();
                    sts128(QKVt + tc5::toff(row_e, 2 * gq), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                           pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    sts128(QKVt + tc5::toff(row_e, 2 * gq + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]),
                           pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
                }
            }
            tc5::fence_before_sync();
            team_sync(team);
            // ---- softmax(q k^T) v per (sequence, head) -> bf16 o tile
            if (!LONG) attn_core_bf16<DH>(QKVt, hc, Ot, nseq_t, S, warp2, TEAM_THREADS / 32, lane, cl);
            else attn_core_long_any<DH>(QKVt, hc, Ot, nseq_t, S, warp2, TEAM_THREADS / 32, lane);
            tc5::fence_proxy_async();
            team_sync(team);
            if (tid2 == 0) {
                tc5::fence_after_sync();
                const uint32_t wo = Wo_s + (uint32_t)ch * Np * Cp * 2;
                for (int k = 0; k < Cp / 16; ++k)
                    tc5::mma_f16(tmem_Y, tc5::kdesc(Ot_s, TILE_M, k), tc5::kdesc(wo, Np, k), idesc_o, (ch > 0 || k > 0) ? 1u : 0u);
                if (ch + 1 < a.nchunks) {
                    const uint32_t wq = Wq_s + (uint32_t)(ch + 1) * NCq * Kp * 2;
                    for (int k = 0; k < Kp / 16; ++k)
                        tc5::mma_f16(tmem_Q, tc5::kdesc(At_s, TILE_M, k), tc5::kdesc(wq, NCq, k), idesc_q, k > 0);
                }
                tc5::mma_commit(bar);
            }
        }
        tc5::mbar_wait(bar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 2: out = res + alpha * (y + bo)
        {
            const int ls = row_e / S, pos = row_e - ls * S;
            const bool valid = row_e < R;
            const long long gr = valid ? a.g.grow(s0 + ls, pos) : 0;
            const int ng = Np >> 3, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
            for (int gq = g0; gq < g1; ++gq) {
                if (gq * 8 >= D) break;
                float v[8], rv[8];
                tc5::tmem_ld8(tmem_Y + lane_base + gq * 8, v);
                if (valid && a.res) load8<VEC4>(a.res + gr * D, gq * 8, D, rv);
                tc5::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        v[k] = a.alpha * (v[k] + bos[gq * 8 + k]);
                        if (a.res) v[k] += rv[k];
                    }
                    store8<VEC4>(a.out + gr * D, gq * 8, D, v);
                }
            }
        }
        tc5::fence_before_sync();
    }
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tmem_base_s, 512);
}


}  // namespace rat

using namespace rat;

template <int KCH, bool VEC4>
static int launch_ff_fwd_tc(const FFTcArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_fwd_tc<KCH, VEC4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin() - 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_fwd_tc)");
        attr_set = true;
    }
    const long long ntiles = (a.rows + TILE_M - 1) / TILE_M;
    const int grid = (int)std::min<long long>((ntiles + 1) / 2, (long long)num_sms());
    k_ff_fwd_tc<KCH, VEC4><<<grid, TC_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_ff_fwd_tc");
    return RAT_OK;
}

// returns RAT_OK if launched, 1 if this shape is not covered by the tcgen05 path (caller falls back to the mma.sync kernel)
int ff_fwd_tc_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                       const float* W1, const float* b1, const float* W2, const float* b2, long long rows, int D, int M,
                       cudaStream_t st) {
    if (!ff_tc_supported(D, M) || res == nullptr) return 1;
    FFTcArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.rows = rows; a.D = D; a.M = M; a.Kp = pad16(D); a.Mp = pad16(M); a.Np = pad16(D);
    const size_t fixed = (size_t)a.Mp * a.Kp * 2 + (size_t)a.Np * a.Mp * 2 + (size_t)(a.Mp + a.Np) * 4;
    const size_t team = (size_t)TILE_M * a.Kp * 2 + (size_t)TILE_M * a.Mp * 2;
    a.smem_bytes = (int)(fixed + 2 * team);
    if (a.smem_bytes > max_smem_optin() - 1024) return 1;
    const int kch = a.Kp / 16;       // chunks (of 8 columns) per half row
    const bool v4 = (D % 4) == 0;
#define RAT_FF_TC(K_) (v4 ? launch_ff_fwd_tc<K_, true>(a, st) : launch_ff_fwd_tc<K_, false>(a, st))
    switch (kch) {
        case 1: return RAT_FF_TC(1);
        case 2: return RAT_FF_TC(2);
        case 3: return RAT_FF_TC(3);
        case 4: return RAT_FF_TC(4);
        default: return 1;
    }
#undef RAT_FF_TC
}

template <int DH, int KCH, bool VEC4, bool LONG>
static int launch_attn_fwd_tc_l(const AttnTcArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd_tc<DH, KCH, VEC4, LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd_tc)");
        attr_set = true;
    }
    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    const int grid = (int)std::min<long long>((ntiles + 1) / 2, (long long)num_sms());
    k_attn_fwd_tc<DH, KCH, VEC4, LONG><<<grid, TC_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd_tc");
    return RAT_OK;
}
template <int DH, int KCH, bool VEC4>
static int launch_attn_fwd_tc(const AttnTcArgs& a, cudaStream_t st) {
    return a.g.S <= 16 ? launch_attn_fwd_tc_l<DH, KCH, VEC4, false>(a, st) : launch_attn_fwd_tc_l<DH, KCH, VEC4, true>(a, st);
}

template <int DH>
static int launch_attn_fwd_tc_dh(const AttnTcArgs& a, cudaStream_t st) {
    const int kch = a.Kp / 16;
    const bool v4 = (a.D % 4) == 0;
#define RAT_ATTN_TC(K_) (v4 ? launch_attn_fwd_tc<DH, K_, true>(a, st) : launch_attn_fwd_tc<DH, K_, false>(a, st))
    switch (kch) {
        case 1: return RAT_ATTN_TC(1);
        case 2: return RAT_ATTN_TC(2);
        case 3: return RAT_ATTN_TC(3);
        case 4: return RAT_ATTN_TC(4);
        default: return 1;
    }
#undef RAT_ATTN_TC
}

// returns RAT_OK if launched, 1 if the shape is not covered (caller falls back), <0 on error
int attn_fwd_tc_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                         const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                         int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st) {
    const int S = mode == 0 ? N : T;
    if (S > TILE_M || S < 1 || (dh != 10 && dh != 20 && dh != 8) || D < 2 || (D & 1) || D > 64) return 1;
    const int DHP = pad16(dh);
    AttnTcArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo; a.bo = bo;
    a.g.S = S; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.D = D; a.H = heads; a.I = heads * dh; a.qscale = scale * 1.4426950408889634f; a.alpha = alpha;
    a.Kp = pad16(D); a.Np = pad16(D);
    int hc = 0;
    for (int c = heads; c >= 1; --c) {
        if (heads % c) continue;
        const int ncq = 3 * c * DHP;
        if (ncq <= 256 && ncq + a.Np <= 256 && ((ncq >> 4) % 2 == 0)) { hc = c; break; }
    }
    if (!hc) return 1;
    a.hc = hc; a.nchunks = heads / hc; a.NCq = 3 * hc * DHP; a.Cp = pad16(hc * dh);
    a.SPT = TILE_M / S;
    if (S <= 8) a.SPT &= ~1;                    // sequences are processed in pairs
    const size_t fixed = (size_t)a.nchunks * a.NCq * a.Kp * 2 + (size_t)a.nchunks * a.Np * a.Cp * 2 + (size_t)a.Np * 4;
    const size_t team = (size_t)TILE_M * (a.Kp + a.NCq + a.Cp) * 2;
    a.smem_bytes = (int)(fixed + 2 * team);
    if (a.smem_bytes > max_smem_optin() - 1024) return 1;
    switch (dh) {
        case 8: return launch_attn_fwd_tc_dh<8>(a, st);
        case 10: return launch_attn_fwd_tc_dh<10>(a, st);
        case 20: return launch_attn_fwd_tc_dh<20>(a, st);
        default: return 1;
    }
}

