// K5 forward, register-resident generation (see encoder_rr.cuh):  out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )
// for sequences of <= 16 tokens, head width <= 16, D <= 48 (every RAT_m1/m2 configuration).  Replaces PreNorm + Attention +
// residual of the reference (models/RAT_m2.py:176-236; intra mode 0: sequence = the N tokens of one (b, t); cross mode 1:
// sequence = the T retrieved samples of one (b, n)).
#include "encoder_rr.cuh"

namespace rat {

struct AttnRRArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv; const float* Wo; const float* bo;
    long long nseq;
    SeqGeom g;
    int D, H, I, dh;
    float qscale, alpha;
};

// KS = pad16(D) / 16 k-steps of the projections, NTO = ceil(D / 8) n-tiles of a token row.  WARPS x CTAS warps per SM.
// BULK (D % 4 == 0): every warp keeps the token rows of its NEXT task in flight as 1-D bulk asynchronous copies
// (cp.async.bulk, one per row, completion counted on a per-warp mbarrier) into a private double buffer, so a task starts on
// rows that are already in shared memory and the residual add re-reads them there instead of from L2.
template <int KS, int NTO, int WARPS, int CTAS, bool BULK, bool F16P, bool F16C, int UNR>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_attn_fwd_rr(AttnRRArgs a) {
    extern __shared__ __align__(16) uint4 rr_smem[];
    constexpr int NP = (NTO + 1) / 2;                     // n-tile pairs of the out-projection
    const int H = a.H, D = a.D, dh = a.dh, I = a.I;
    uint4* Wq_i = rr_smem;                                // [H][KS][32]
    uint4* Wk_i = Wq_i + H * KS * 32;
    uint4* Wv_i = Wk_i + H * KS * 32;
    uint4* Wo_i = Wv_i + H * KS * 32;                     // [H][NP][32]
    float* lnw_s = reinterpret_cast<float*>(Wo_i + H * NP * 32);   // [KS * 16], zero padded
    float* lnb_s = lnw_s + KS * 16;
    float* bo_s = lnb_s + KS * 16;                        // [NP * 16]
    float* stage = bo_s + NP * 16;                        // BULK: [WARPS][2][16 rows][D] ; during setup: raw fp32 weights
    __shared__ __align__(8) uint64_t row_bar[WARPS][2];
    pdl_launch_dependents();
    {
        // raw weights with coalesced loads (one memory latency), then the fragment-order images from shared memory
        float* raw = stage;                               // Wq | Wk | Wv [3][I][D], Wo [D][I]
        const int nw = I * D;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            raw[i] = __ldg(a.Wq + i); raw[nw + i] = __ldg(a.Wk + i); raw[2 * nw + i] = __ldg(a.Wv + i); raw[3 * nw + i] = __ldg(a.Wo + i);
        }
        for (int i = threadIdx.x; i < KS * 16; i += blockDim.x) {
            lnw_s[i] = i < D ? a.ln_w[i] : 0.f;
            lnb_s[i] = i < D ? a.ln_b[i] : 0.f;
        }
        for (int i = threadIdx.x; i < NP * 16; i += blockDim.x) bo_s[i] = i < D ? a.bo[i] : 0.f;
        __syncthreads();
        const int nqkv = H * KS * 32;
        for (int i = threadIdx.x; i < 3 * nqkv; i += blockDim.x) {
            const int w = i / nqkv, r = i - w * nqkv;
            const int h = r / (KS * 32), ks = (r >> 5) % KS, ln = r & 31;
            const float* W = raw + w * nw + h * dh * D;
            const float mul = w == 0 ? a.qscale : 1.0f;
            rr_smem[i] = frag_pair_entry(ln, 0, 16 * ks, [&](int n, int k) { return (n < dh && k < D) ? mul * W[n * D + k] : 0.f; });
        }
        for (int i = threadIdx.x; i < H * NP * 32; i += blockDim.x) {
            const int h = i / (NP * 32), p = (i >> 5) % NP, ln = i & 31;
            const float* W = raw + 3 * nw + h * dh;
            Wo_i[i] = frag_pair_entry(ln, 16 * p, 0, [&](int c, int dd) { return (c < D && dd < dh) ? W[c * I + dd] : 0.f; });
        }
        if (threadIdx.x < WARPS * 2) tc5::mbar_init(&row_bar[0][0] + threadIdx.x, 1);
        tc5::fence_mbar_init();
        tc5::fence_proxy_async();        // the raw-weight region becomes the target of bulk copies
    }
    __syncthreads();
    pdl_wait();                          // x (and res) are written by the previous kernel of the stream
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    const int S = a.g.S;
    const float invD = 1.0f / (float)D;
    const RRLane cl = make_rr_lane(S, lane);
    const long long ntasks = cl.packed ? (a.nseq + 1) >> 1 : a.nseq;
    const long long wstride = (long long)gridDim.x * WARPS;
    const bool res_is_x = a.res == a.x;
    // bulk staging: lane i < 16 owns fragment row i of the task
    float* my_stage = stage + (size_t)warp * 2 * 16 * D;
    const int c_sq = cl.packed ? (lane >> 3) & 1 : 0, c_pos = cl.packed ? (lane & 7) : lane;
    const bool c_row = lane < 16 && c_pos < S;
    auto issue_rows = [&](long long task, int buf) {
        const long long seq0 = cl.packed ? 2 * task : task;
        const bool ok = c_row && seq0 + c_sq < a.nseq;
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) mbar_arrive_expect_tx(&row_bar[warp][buf], (uint32_t)(__popc(m) * D * 4));
        __syncwarp();
        if (ok) bulk_g2s(tc5::smem_u32(my_stage + (size_t)(buf * 16 + lane) * D), a.x + a.g.grow(seq0 + c_sq, c_pos) * D, (uint32_t)(D * 4),
                         &row_bar[warp][buf]);
    };
    long long task = (long long)blockIdx.x * WARPS + warp;
    uint32_t ph0 = 0, ph1 = 0;
    int buf = 0;
    if (BULK && task < ntasks) issue_rows(task, 0);
    for (; task < ntasks; task += wstride, buf ^= 1) {
        const long long seq0 = cl.packed ? 2 * task : task;
        const bool vlo = cl.lo_pos >= 0, vhi = cl.hi_pos >= 0 && seq0 + cl.hi_sq < a.nseq;
        const long long rlo = vlo ? a.g.grow(seq0, cl.lo_pos) : 0, rhi = vhi ? a.g.grow(seq0 + cl.hi_sq, cl.hi_pos) : 0;
        const float* slo = my_stage + (size_t)(buf * 16 + g) * D;          // staged rows g and g + 8
        const float* shi = slo + 8 * D;
        // ---- rows -> LayerNorm -> A fragments
        uint32_t xa[KS][4];
        {
            float2 xl[NTO], xh[NTO];
            if (BULK) {
                if (task + wstride < ntasks) issue_rows(task + wstride, buf ^ 1);     // the other buffer was consumed a task ago
                tc5::mbar_wait(&row_bar[warp][buf], buf ? ph1 : ph0);
                if (buf) ph1 ^= 1; else ph0 ^= 1;
                rr_load_rows<NTO>(slo, shi, vlo, vhi, D, t, xl, xh);
            } else {
                rr_load_rows<NTO>(a.x + rlo * D, a.x + rhi * D, vlo, vhi, D, t, xl, xh);
            }
            float ml, rl, mh, rh;
            rr_row_stats<NTO>(xl, D, invD, t, ml, rl);
            rr_row_stats<NTO>(xh, D, invD, t, mh, rh);
#pragma unroll
            for (int nt = 0; nt < 2 * KS; ++nt) {
                uint32_t lo = 0u, hi = 0u;
                if (nt < NTO) {
                    const float2 w = *reinterpret_cast<const float2*>(lnw_s + 8 * nt + 2 * t);
                    const float2 b = *reinterpret_cast<const float2*>(lnb_s + 8 * nt + 2 * t);
                    lo = vlo ? pack_h2(fmaf((xl[nt].x - ml) * rl, w.x, b.x), fmaf((xl[nt].y - ml) * rl, w.y, b.y)) : 0u;
                    hi = vhi ? pack_h2(fmaf((xh[nt].x - mh) * rh, w.x, b.x), fmaf((xh[nt].y - mh) * rh, w.y, b.y)) : 0u;
                }
                xa[nt >> 1][(nt & 1) * 2] = lo;
                xa[nt >> 1][(nt & 1) * 2 + 1] = hi;
            }
        }
        float acc[NTO][4];
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const uint4* wq = Wq_i + lane;
        const uint4* wk = Wk_i + lane;
        const uint4* wv = Wv_i + lane;
        const uint4* wo = Wo_i + lane;
#pragma unroll UNR
        for (int h = 0; h < H; ++h) {
            // O (one k-step) accumulates in fp16: its packed accumulators ARE the out-projection's operand fragment; the
            // projections (three k-steps) do the same when F16P
            ProjAcc<F16P> q, k, vt;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 fq = wq[ks * 32], fk = wk[ks * 32], fv = wv[ks * 32];
                q.mma(0, xa[ks], fq.x, fq.y);
                q.mma(1, xa[ks], fq.z, fq.w);
                k.mma(0, xa[ks], fk.x, fk.y);
                k.mma(1, xa[ks], fk.z, fk.w);
                const uint32_t av[4] = {fv.x, fv.z, fv.y, fv.w};
                vt.mma(0, av, xa[ks][0], xa[ks][2]);                 // tokens 0..7  (fragment rows g)
                vt.mma(1, av, xa[ks][1], xa[ks][3]);                 // tokens 8..15 (fragment rows g + 8)
            }
            wq += KS * 32; wk += KS * 32; wv += KS * 32;
            uint32_t qa[4], ka[4], va[4];                            // 8x8 blocks {lo/0-7, hi/0-7, lo/8-15, hi/8-15}
            q.frag(qa); k.frag(ka); vt.frag(va);
            float sc[2][4] = {};
            rr_mma(sc[0], qa, ka[0], ka[2]);                  // keys 0..7  (rows g of k)
            rr_mma(sc[1], qa, ka[1], ka[3]);                  // keys 8..15 (rows g + 8)
            rr_softmax(sc, cl, vlo, vhi);
            uint32_t pa[4];
            c_to_a(sc, pa);
            ProjAcc<F16C> o;
            o.mma(0, pa, va[0], va[2]);                              // d 0..7
            o.mma(1, pa, va[1], va[3]);                              // d 8..15
            uint32_t oa[4];
            o.frag(oa);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint4 f = wo[p * 32];
                rr_mma(acc[2 * p], oa, f.x, f.y);
                if (2 * p + 1 < NTO) rr_mma(acc[2 * p + 1], oa, f.z, f.w);
            }
            wo += NP * 32;
        }
        // ---- out = res + alpha * (acc + bo)
        {
            const bool from_stage = BULK && res_is_x;
            const float* rl_p = from_stage ? slo : a.res ? a.res + rlo * D : nullptr;
            const float* rh_p = from_stage ? shi : a.res ? a.res + rhi * D : nullptr;
            float* ol = a.out + rlo * D;
            float* oh = a.out + rhi * D;
#pragma unroll
            for (int nt = 0; nt < NTO; ++nt) {
                const int c = 8 * nt + 2 * t;
                if (c < D) {
                    const float2 b = *reinterpret_cast<const float2*>(bo_s + c);
                    if (vlo) {
                        float2 r = rl_p ? *reinterpret_cast<const float2*>(rl_p + c) : make_float2(0.f, 0.f);
                        r.x = fmaf(a.alpha, acc[nt][0] + b.x, r.x); r.y = fmaf(a.alpha, acc[nt][1] + b.y, r.y);
                        *reinterpret_cast<float2*>(ol + c) = r;
                    }
                    if (vhi) {
                        float2 r = rh_p ? *reinterpret_cast<const float2*>(rh_p + c) : make_float2(0.f, 0.f);
                        r.x = fmaf(a.alpha, acc[nt][2] + b.x, r.x); r.y = fmaf(a.alpha, acc[nt][3] + b.y, r.y);
                        *reinterpret_cast<float2*>(oh + c) = r;
                    }
                }
            }
        }
        if (BULK) __syncwarp();          // every lane is done with this buffer before it is refilled (next iteration's issue)
    }
}

template <int KS, int NTO, int WARPS, int CTAS, bool BULK, bool F16P, bool F16C, int UNR>
static int launch_attn_fwd_rr_v(const AttnRRArgs& a, cudaStream_t st) {
    constexpr int NP = (NTO + 1) / 2;
    const size_t stage_bytes = std::max((size_t)(BULK ? WARPS * 2 * 16 * a.D * 4 : 0), (size_t)4 * a.I * a.D * 4);
    const size_t smem = ((size_t)3 * a.H * KS * 32 + (size_t)a.H * NP * 32) * sizeof(uint4) + (size_t)(2 * KS * 16 + NP * 16) * 4 + stage_bytes;
    if (smem > (size_t)max_smem_optin() / CTAS - 2048) return 1;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd_rr<KS, NTO, WARPS, CTAS, BULK, F16P, F16C, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd_rr)");
        attr_smem = smem;
    }
    const long long ntasks = a.g.S <= 8 ? (a.nseq + 1) / 2 : a.nseq;
    const long long nblk = (ntasks + WARPS - 1) / WARPS;
    const int grid = (int)std::min<long long>(nblk, (long long)CTAS * num_sms());
    if (launch_pdl(k_attn_fwd_rr<KS, NTO, WARPS, CTAS, BULK, F16P, F16C, UNR>, dim3(grid), dim3(WARPS * 32), smem, st, a) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "k_attn_fwd_rr");
    RAT_CHECK_LAUNCH("k_attn_fwd_rr");
    return RAT_OK;
}
// RAT_RR_FWD_VARIANT (tuning aid): 0 = bulk row staging, fp32 accumulators (default) ; 1 = fp16 accumulators for q, k, v^T, O
// (4 % faster, but the 15-step AUC / logloss comparison with the fp32 oracle drifts past the 1e-3 bar) ; 2 = direct loads
template <int KS, int NTO>
static int launch_attn_fwd_rr(const AttnRRArgs& a, cudaStream_t st) {
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("RAT_RR_FWD_VARIANT"); variant = e ? atoi(e) : 0; }
    const bool bulk_ok = (a.D % 4) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    int rc = 1;
    if (bulk_ok && variant == 0) rc = launch_attn_fwd_rr_v<KS, NTO, 16, 1, true, false, false, 1>(a, st);
    if (bulk_ok && variant == 1) rc = launch_attn_fwd_rr_v<KS, NTO, 16, 1, true, true, true, 1>(a, st);
    if (bulk_ok && variant == 4) rc = launch_attn_fwd_rr_v<KS, NTO, 16, 1, true, false, false, 2>(a, st);
    if (bulk_ok && variant == 5) rc = launch_attn_fwd_rr_v<KS, NTO, 12, 1, true, false, false, 2>(a, st);
    if (bulk_ok && variant == 6) rc = launch_attn_fwd_rr_v<KS, NTO, 8, 1, true, false, false, 4>(a, st);
    if (rc == 1) rc = launch_attn_fwd_rr_v<KS, NTO, 16, 1, false, false, false, 1>(a, st);
    return rc;
}

}  // namespace rat

using namespace rat;

// returns 1 when the shape is outside this kernel's envelope (the caller falls back to the tile kernels)
int attn_fwd_rr_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                         const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                         int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st) {
    const int S = mode == 0 ? N : T;
    if (S < 1 || S > 16 || dh < 1 || dh > 16 || D < 2 || (D & 1) || D > 48) return 1;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 7) != 0) return 1;
    AttnRRArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo; a.bo = bo;
    a.g.S = S; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.D = D; a.H = heads; a.I = heads * dh; a.dh = dh; a.qscale = scale * 1.4426950408889634f; a.alpha = alpha;
    const int NTO = (D + 7) / 8;
    switch (NTO) {
        case 1: return launch_attn_fwd_rr<1, 1>(a, st);
        case 2: return launch_attn_fwd_rr<1, 2>(a, st);
        case 3: return launch_attn_fwd_rr<2, 3>(a, st);
        case 4: return launch_attn_fwd_rr<2, 4>(a, st);
        case 5: return launch_attn_fwd_rr<3, 5>(a, st);
        case 6: return launch_attn_fwd_rr<3, 6>(a, st);
        default: return 1;
    }
}
