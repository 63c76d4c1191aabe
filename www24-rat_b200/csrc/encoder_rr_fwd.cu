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

// KS = pad16(D) / 16 k-steps of the projections, NTO = ceil(D / 8) n-tiles of a token row.
// WARPS x CTAS warps per SM and UNR heads in flight per warp trade registers for latency hiding (HMMA latency on B200 is
// ~100 cycles: tools/hmma_probe.cu, so one head at a time leaves a warp mostly waiting on its own dependency chain).
template <int KS, int NTO, int WARPS, int CTAS, int UNR>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_attn_fwd_rr(AttnRRArgs a) {
    constexpr int RR_FWD_THREADS = WARPS * 32;
    extern __shared__ __align__(16) uint4 rr_smem[];
    constexpr int NP = (NTO + 1) / 2;                     // n-tile pairs of the out-projection
    const int H = a.H, D = a.D, dh = a.dh;
    uint4* Wq_i = rr_smem;                                // [H][KS][32]
    uint4* Wk_i = Wq_i + H * KS * 32;
    uint4* Wv_i = Wk_i + H * KS * 32;
    uint4* Wo_i = Wv_i + H * KS * 32;                     // [H][NP][32]
    float* lnw_s = reinterpret_cast<float*>(Wo_i + H * NP * 32);   // [KS * 16], zero padded
    float* lnb_s = lnw_s + KS * 16;
    float* bo_s = lnb_s + KS * 16;                        // [NP * 16]
    {
        const int nqkv = H * KS * 32;
        for (int i = threadIdx.x; i < 3 * nqkv; i += blockDim.x) {
            const int w = i / nqkv, r = i - w * nqkv;
            const int h = r / (KS * 32), ks = (r >> 5) % KS, ln = r & 31;
            const float* W = (w == 0 ? a.Wq : w == 1 ? a.Wk : a.Wv) + (size_t)h * dh * D;
            const float mul = w == 0 ? a.qscale : 1.0f;
            rr_smem[i] = frag_pair_entry(ln, 0, 16 * ks, [&](int n, int k) {
                return (n < dh && k < D) ? mul * __ldg(W + (size_t)n * D + k) : 0.f; });
        }
        for (int i = threadIdx.x; i < H * NP * 32; i += blockDim.x) {
            const int h = i / (NP * 32), p = (i >> 5) % NP, ln = i & 31;
            const float* W = a.Wo + h * dh;
            Wo_i[i] = frag_pair_entry(ln, 16 * p, 0, [&](int c, int dd) {
                return (c < D && dd < dh) ? __ldg(W + (size_t)c * a.I + dd) : 0.f; });
        }
        for (int i = threadIdx.x; i < KS * 16; i += blockDim.x) {
            lnw_s[i] = i < D ? a.ln_w[i] : 0.f;
            lnb_s[i] = i < D ? a.ln_b[i] : 0.f;
        }
        for (int i = threadIdx.x; i < NP * 16; i += blockDim.x) bo_s[i] = i < D ? a.bo[i] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
    const int S = a.g.S;
    const RRLane cl = make_rr_lane(S, lane);
    const long long ntasks = cl.packed ? (a.nseq + 1) >> 1 : a.nseq;
    const long long wstride = (long long)gridDim.x * (RR_FWD_THREADS / 32);
    for (long long task = (long long)blockIdx.x * (RR_FWD_THREADS / 32) + warp; task < ntasks; task += wstride) {
        const long long seq0 = cl.packed ? 2 * task : task;
        const bool vlo = cl.lo_pos >= 0, vhi = cl.hi_pos >= 0 && seq0 + cl.hi_sq < a.nseq;
        const long long rlo = vlo ? a.g.grow(seq0, cl.lo_pos) : 0, rhi = vhi ? a.g.grow(seq0 + cl.hi_sq, cl.hi_pos) : 0;
        // ---- rows -> LayerNorm -> A fragments
        uint32_t xa[KS][4];
        {
            float2 xl[NTO], xh[NTO];
            rr_load_rows<NTO>(a.x + rlo * D, a.x + rhi * D, vlo, vhi, D, t, xl, xh);
            float ml, rl, mh, rh;
            rr_row_stats<NTO>(xl, D, t, ml, rl);
            rr_row_stats<NTO>(xh, D, t, mh, rh);
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
            float q[2][4] = {}, k[2][4] = {}, vt[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 fq = wq[ks * 32], fk = wk[ks * 32], fv = wv[ks * 32];
                mma_h_16x8x16(q[0], xa[ks], fq.x, fq.y);
                mma_h_16x8x16(q[1], xa[ks], fq.z, fq.w);
                mma_h_16x8x16(k[0], xa[ks], fk.x, fk.y);
                mma_h_16x8x16(k[1], xa[ks], fk.z, fk.w);
                const uint32_t av[4] = {fv.x, fv.z, fv.y, fv.w};
                mma_h_16x8x16(vt[0], av, xa[ks][0], xa[ks][2]);      // tokens 0..7  (fragment rows g)
                mma_h_16x8x16(vt[1], av, xa[ks][1], xa[ks][3]);      // tokens 8..15 (fragment rows g + 8)
            }
            wq += KS * 32; wk += KS * 32; wv += KS * 32;
            uint32_t qa[4];
            c_to_a(q, qa);
            float sc[2][4] = {};
            mma_h_16x8x16(sc[0], qa, pack_h2(k[0][0], k[0][1]), pack_h2(k[1][0], k[1][1]));     // keys 0..7
            mma_h_16x8x16(sc[1], qa, pack_h2(k[0][2], k[0][3]), pack_h2(k[1][2], k[1][3]));     // keys 8..15
            rr_softmax(sc, cl, vlo, vhi);
            uint32_t pa[4];
            c_to_a(sc, pa);
            float o[2][4] = {};
            mma_h_16x8x16(o[0], pa, pack_h2(vt[0][0], vt[0][1]), pack_h2(vt[1][0], vt[1][1]));  // d 0..7
            mma_h_16x8x16(o[1], pa, pack_h2(vt[0][2], vt[0][3]), pack_h2(vt[1][2], vt[1][3]));  // d 8..15
            uint32_t oa[4];
            c_to_a(o, oa);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint4 f = wo[p * 32];
                mma_h_16x8x16(acc[2 * p], oa, f.x, f.y);
                if (2 * p + 1 < NTO) mma_h_16x8x16(acc[2 * p + 1], oa, f.z, f.w);
            }
            wo += NP * 32;
        }
        // ---- out = res + alpha * (acc + bo)
        {
            const float* rl_p = a.res ? a.res + rlo * D : nullptr;
            const float* rh_p = a.res ? a.res + rhi * D : nullptr;
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
    }
}

template <int KS, int NTO, int WARPS, int CTAS, int UNR>
static int launch_attn_fwd_rr_v(const AttnRRArgs& a, cudaStream_t st) {
    constexpr int NP = (NTO + 1) / 2;
    const size_t smem = ((size_t)3 * a.H * KS * 32 + (size_t)a.H * NP * 32) * sizeof(uint4) + (size_t)(2 * KS * 16 + NP * 16) * 4;
    if (smem > (size_t)max_smem_optin() / CTAS - 2048) return 1;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd_rr<KS, NTO, WARPS, CTAS, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd_rr)");
        attr_smem = smem;
    }
    const long long ntasks = a.g.S <= 8 ? (a.nseq + 1) / 2 : a.nseq;
    const long long nblk = (ntasks + WARPS - 1) / WARPS;
    const int grid = (int)std::min<long long>(nblk, (long long)CTAS * num_sms());
    k_attn_fwd_rr<KS, NTO, WARPS, CTAS, UNR><<<grid, WARPS * 32, smem, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd_rr");
    return RAT_OK;
}
// RAT_RR_FWD_VARIANT (tuning aid): 0 = 8 warps x 2 CTAs, one head in flight ; 1 = 12 warps, 2 heads ; 2 = 8 warps, 4 heads
template <int KS, int NTO>
static int launch_attn_fwd_rr(const AttnRRArgs& a, cudaStream_t st) {
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("RAT_RR_FWD_VARIANT"); variant = e ? atoi(e) : 0; }
    switch (variant) {
        case 1: return launch_attn_fwd_rr_v<KS, NTO, 12, 1, 2>(a, st);
        case 2: return launch_attn_fwd_rr_v<KS, NTO, 8, 1, 4>(a, st);
        case 3: return launch_attn_fwd_rr_v<KS, NTO, 8, 2, 2>(a, st);
        case 4: return launch_attn_fwd_rr_v<KS, NTO, 16, 1, 1>(a, st);
        default: return launch_attn_fwd_rr_v<KS, NTO, 8, 2, 1>(a, st);
    }
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
