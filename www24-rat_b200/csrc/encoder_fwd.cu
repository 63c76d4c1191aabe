// K2: fused RAT-block forward kernels.
//
//   k_attn_fwd : out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )        PreNorm+Attention+residual
//                (reference: PreNorm RAT_m2.py:155-161, Attention RAT_m2.py:176-202, residual :224/:231;
//                 RAT_m3.Attention RAT_m3.py:164-196 with separate Wq/Wk/Wv pointers)
//   k_ff_fwd   : out = res + W2 gelu(W1 [LN](x) + b1) + b2                  FeedForward (+optional PreNorm)
//                (reference: FeedForward RAT_m2.py:163-174 ; PreNorm'd in RAT_m0.py:193-208)
//   k_ln_fwd   : out = LayerNorm(x)                                         final norm of RAT_m0/m1 Transformer
//
// One persistent CTA per SM (512 threads) keeps ALL weights of the sub-block resident in shared memory (natural
// torch layout, zero padded) and walks tiles of whole sequences (SPT sequences x S positions = R token rows).
// LayerNorm output, the per-head-chunk q|k|v, the attention output and the out-projection accumulator stay in
// shared memory.  The projections run on the tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate) or, with
// precision=fp32, on an exact SIMT twin; softmax(QK^T)V runs on the SIMT pipe with all S scores in registers.
// "Intra" attention (sequence = one sample row, positions = fields) and "cross" attention (sequence = one field
// over the 1+K retrieved rows) differ only in the row-index map SeqGeom::grow, i.e. the reference's
// reshape/transpose/flatten copies (RAT_m2.py:221-235) are pure indexing here.
#include "tile.cuh"
#include "encoder_common.cuh"
#include <cstdlib>
#include "attn_mma.cuh"
#include "../../include/rat_b200.h"

namespace rat {

struct AttnArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* Wq; const float* Wk; const float* Wv;   // [I, D] each (row-major, torch Linear layout)
    const float* Wo; const float* bo;                    // [D, I], [D]
    long long nseq;                                      // total sequences
    SeqGeom g;
    int D, H, I;
    float scale, alpha;
    AttnPlan p;
};

template <int DH, bool MMA>
__global__ void __launch_bounds__(ENC_THREADS, 1) k_attn_fwd(AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const AttnPlan& p = a.p;
    const int S = a.g.S, D = a.D;
    const int Dl = p.Dl, C3l = p.C3l, Cql = p.Cql, Il = p.Il;
    const int Rmax = p.Rmax16;
    float* Wc = smem;                                   // [nchunks][C3p8][Dl]   q|k|v rows of each head chunk
    float* WoN = Wc + (size_t)p.nchunks * p.C3p8 * Dl;  // [Dp8][Il]             natural Wo
    float* as = WoN + (size_t)p.Dp8 * Il;               // [Rmax][Dl]  LayerNorm(x)
    float* ys = as + (size_t)Rmax * Dl;                 // [Rmax][Dl]  out-projection accumulator
    float* qkv = ys + (size_t)Rmax * Dl;                // [Rmax][C3l]
    float* os = qkv + (size_t)Rmax * C3l;               // [Rmax][Cql]
    long long* rowidx = reinterpret_cast<long long*>(os + (size_t)Rmax * Cql);   // [Rmax]
    stage_qkv_chunks(a.Wq, a.Wk, a.Wv, D, p, Wc);
    stage_padded(a.Wo, D, a.I, p.Dp8, Il, WoN);
    zero_floats(as, (size_t)Rmax * (2 * Dl + C3l + Cql));
    const long long ntiles = (a.nseq + p.SPT - 1) / p.SPT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * p.SPT;
        const int nseq_t = (int)min((long long)p.SPT, a.nseq - s0);
        const int R = nseq_t * S;
        __syncthreads();                                // previous tile fully consumed (and staging visible)
        fill_rowidx(rowidx, a.g, s0, R);
        __syncthreads();
        ln_rows_to_smem(a.x, rowidx, R, D, p.Dp8, a.ln_w, a.ln_b, as, Dl, p.lg, nullptr);
        zero_rows(as, Dl, R, pad16(R));
        zero_rows(os, Cql, R, pad16(R));
        zero_floats(ys, (size_t)pad16(R) * Dl);
        for (int ch = 0; ch < p.nchunks; ++ch) {
            const float* W = Wc + (size_t)ch * p.C3p8 * Dl;
            __syncthreads();                            // LN / previous chunk's out-projection done
            // q|k|v[r][c] = sum_d as[r][d] * W[c][d]
            tc_gemm<MMA, 4>(as, Dl, 1, W, 1, Dl, qkv, C3l, R, p.C3p8, p.Dp8, false, EpiNone2());
            __syncthreads();
            if (MMA && S <= 16) attn_fwd_mma<DH>(qkv, C3l, p.Cq, os, Cql, nullptr, nseq_t, S, p.hc, a.scale);
            else attn_core<DH>(qkv, C3l, p.Cq, os, Cql, nullptr, nseq_t, S, p.hc, p.lpt, a.scale);
            __syncthreads();
            // ys[r][d] += sum_c os[r][c] * Wo[d][col0 + c]
            tc_gemm<MMA, 3>(os, Cql, 1, WoN + ch * p.Cq, 1, Il, ys, Dl, R, p.Dp8, p.Cq8, true, EpiNone2());
        }
        __syncthreads();
        {
            const int lg = p.lg, groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
            for (int r = gi; r < R; r += groups) {
                const long long gr = rowidx[r] * D;
                for (int d = li; d < D; d += lg) {
                    float v = a.alpha * (ys[(size_t)r * Dl + d] + a.bo[d]);
                    if (a.res) v += a.res[gr + d];
                    a.out[gr + d] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
struct FFArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;               // nullptr => no pre-norm (RAT_m2/m3)
    const float* W1; const float* b1; const float* W2; const float* b2;   // [M,D],[M],[D,M],[D]
    long long rows;
    int D, M;
    FFPlan p;
};

struct EpiBiasGelu2 {
    const float* b;
    __device__ __forceinline__ void operator()(int, int c, float& v0, float& v1) const {
        v0 = gelu_erf(v0 + b[c]);
        v1 = gelu_erf(v1 + b[c + 1]);
    }
};

template <bool MMA>
__global__ void __launch_bounds__(ENC_THREADS, 1) k_ff_fwd(FFArgs a) {
    extern __shared__ __align__(16) float smem[];
    const FFPlan& p = a.p;
    const int D = a.D, M = a.M, Dl = p.Dl, Ml = p.Ml;
    float* W1n = smem;                                  // [Mp8][Dl]  natural W1 [M,D]
    float* W2n = W1n + (size_t)p.Mp8 * Dl;              // [Dp8][Ml]  natural W2 [D,M]
    float* b1s = W2n + (size_t)p.Dp8 * Ml;              // [Mp8]
    float* xs = b1s + p.Mp8;                            // [RPT][Dl]
    float* hs = xs + (size_t)p.RPT * Dl;                // [RPT][Ml]
    float* ys = hs + (size_t)p.RPT * Ml;                // [RPT][Dl]
    long long* rowidx = reinterpret_cast<long long*>(ys + (size_t)p.RPT * Dl);   // [RPT]
    stage_padded(a.W1, M, D, p.Mp8, Dl, W1n);
    stage_padded(a.W2, D, M, p.Dp8, Ml, W2n);
    for (int i = threadIdx.x; i < p.Mp8; i += blockDim.x) b1s[i] = i < M ? a.b1[i] : 0.f;
    zero_floats(xs, (size_t)p.RPT * (2 * Dl + Ml));
    const long long ntiles = (a.rows + p.RPT - 1) / p.RPT;
    SeqGeom flat{1, 0, 1, 1};
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * p.RPT;
        const int R = (int)min((long long)p.RPT, a.rows - r0);
        __syncthreads();
        if (a.ln_w) {
            fill_rowidx(rowidx, flat, r0, R);
            __syncthreads();
            ln_rows_to_smem(a.x, rowidx, R, D, p.Dp8, a.ln_w, a.ln_b, xs, Dl, p.lg, nullptr);
        } else {
            for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
                const int r = i / D, d = i - r * D;
                xs[(size_t)r * Dl + d] = a.x[(r0 + r) * D + d];
            }
        }
        zero_rows(xs, Dl, R, pad16(R));
        __syncthreads();
        // h[r][m] = gelu(sum_d x[r][d] W1[m][d] + b1[m])
        tc_gemm<MMA, 4>(xs, Dl, 1, W1n, 1, Dl, hs, Ml, R, p.Mp8, p.Dp8, false, EpiBiasGelu2{b1s});
        __syncthreads();
        // y[r][d] = sum_m h[r][m] W2[d][m]
        tc_gemm<MMA, 3>(hs, Ml, 1, W2n, 1, Ml, ys, Dl, R, p.Dp8, p.Mp8, false, EpiNone2());
        __syncthreads();
        for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
            const int r = i / D, d = i - r * D;
            a.out[(r0 + r) * D + d] = a.res[(r0 + r) * D + d] + ys[(size_t)r * Dl + d] + a.b2[d];
        }
    }
}

__global__ void k_ln_fwd(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ w,
                         const float* __restrict__ b, long long rows, int D, int lg) {
    const int groups = blockDim.x / lg, gi = threadIdx.x / lg, li = threadIdx.x % lg;
    const float invD = 1.0f / (float)D;
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
        if (ok) for (int d = li; d < D; d += lg) out[r * D + d] = (src[d] - mean) * rstd * w[d] + b[d];
    }
}

// ---- host-side planning ---------------------------------------------------------------------------------
// fixed: resident weights ; per row: as + ys (2 Dl) + qkv (C3l) + os (Cql).  Prefer >=128 rows, then wide chunks.
int plan_attn(int S, int D, int H, int dh, AttnPlan* out) {
    const size_t bud = (size_t)(max_smem_optin() - 2048) / 4;
    const int cap_rows = 256;
    AttnPlan best{};
    int bestR = 0;
    for (int hc = H; hc >= 1; --hc) {
        if (H % hc) continue;
        AttnPlan c{};
        fill_attn_plan(S, D, H, dh, hc, &c);
        const size_t fixed = (size_t)c.nchunks * c.C3p8 * c.Dl + (size_t)c.Dp8 * c.Il;
        const size_t per_row = 2 * (size_t)c.Dl + c.C3l + c.Cql + 2;
        if (fixed + per_row * pad16(S) > bud) continue;
        int spt = (int)min((size_t)max(1, cap_rows / S), (bud - fixed) / (per_row * S));
        while (spt > 1 && fixed + per_row * pad16(spt * S) > bud) --spt;
        if (spt < 1) continue;
        const int R = spt * S;
        if (R > bestR) {
            bestR = R; best = c; best.SPT = spt; best.Rmax16 = pad16(R);
            best.smem_bytes = (fixed + per_row * pad16(R)) * 4;
        }
        if (R >= min(128, max(1, cap_rows / S) * S)) break;
    }
    if (!bestR) return RAT_ESMEM;
    *out = best;
    return RAT_OK;
}

int plan_ff(int D, int M, FFPlan* out) {
    FFPlan p{};
    fill_ff_plan(D, M, &p);
    const size_t bud = (size_t)(max_smem_optin() - 2048) / 4;
    const size_t fixed = (size_t)p.Mp8 * p.Dl + (size_t)p.Dp8 * p.Ml + p.Mp8;
    const size_t per_row = 2 * (size_t)p.Dl + p.Ml + 2;
    int rpt = 256;
    while (rpt >= 16 && fixed + per_row * rpt > bud) rpt -= 16;
    if (rpt < 16) return RAT_ESMEM;
    p.RPT = rpt;
    p.smem_bytes = (fixed + per_row * rpt) * 4;
    *out = p;
    return RAT_OK;
}

// 0 = exact fp32 SIMT ; 1 = tf32 mma.sync projections ; 2 = bf16 tcgen05 projections (shapes the tcgen05 kernels do
// not cover run the tf32 kernels)
static int g_precision = 2;     // default: tcgen05 path (fp16 operands, fp32 accumulate); 1 = mma.sync TF32, 0 = fp32 SIMT
int precision_mode() { return g_precision; }

}  // namespace rat

using namespace rat;

int ff_fwd_tc_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                       const float* W1, const float* b1, const float* W2, const float* b2, long long rows, int D, int M,
                       cudaStream_t st);

int attn_fwd_tc_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                         const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                         int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st);

int attn_fwd_tc2_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                          const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                          int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st);

// RAT_TC2=1 in the environment selects the second-generation attention forward (every product on tcgen05, one tile per
// 4-warp group; encoder_tc2_fwd.cu).  It is parity-green but measured no faster than the first generation on B200
// (kkbox B=4096: 128 / 144 us vs 129 / 135 us; DESIGN.md "Attention, second generation"), so it is opt-in.
int attn_fwd_rr_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                         const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                         int N, int D, int heads, int dh, float scale, float alpha, int mode, cudaStream_t st);

int ff_fwd_rr_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b, const float* W1,
                       const float* b1, const float* W2, const float* b2, long long rows, int D, int M, cudaStream_t st);

// The register-resident attention kernels (encoder_rr_*.cu) are the default of the fp16 mode for sequences <= 16 tokens;
// RAT_RR=0 in the environment selects the tile-based tcgen05 kernels for every shape (A/B measurements).
bool rr_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("RAT_RR"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

bool tc2_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("RAT_TC2"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}

extern "C" int rat_set_precision(int mode) {
    RAT_REQUIRE(mode >= 0 && mode <= 2, "rat_set_precision: mode must be 0 (fp32), 1 (tf32) or 2 (fp16 tcgen05)");
    g_precision = mode;
    return RAT_OK;
}
extern "C" int rat_get_precision(void) { return g_precision; }

template <int DH, bool MMA>
static int launch_attn_fwd(const AttnArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd<DH, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd)");
        attr_set = true;
    }
    const long long ntiles = (a.nseq + a.p.SPT - 1) / a.p.SPT;
    const int grid = (int)min(ntiles, (long long)num_sms());
    k_attn_fwd<DH, MMA><<<grid, ENC_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd");
    return RAT_OK;
}
template <int DH>
static int launch_attn_fwd_p(const AttnArgs& a, cudaStream_t st) {
    return g_precision ? launch_attn_fwd<DH, true>(a, st) : launch_attn_fwd<DH, false>(a, st);
}

extern "C" int rat_attn_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                            const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo,
                            int B, int T, int N, int D, int heads, int dim_head, float scale, float alpha, int mode,
                            void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && N > 0 && D > 0 && heads > 0, "rat_attn_fwd: bad shape");
    RAT_REQUIRE(mode == 0 || mode == 1, "rat_attn_fwd: mode must be 0 (intra) or 1 (cross)");
    RAT_REQUIRE(Wo != nullptr && bo != nullptr, "rat_attn_fwd: identity out-projection (heads==1 && dim_head==dim) is not supported");
    if (g_precision == 2) {
        if (rr_enabled() && !getenv("RAT_RR_FWD_OFF")) {
            const int rc4 = attn_fwd_rr_dispatch(x, res, out, ln_w, ln_b, Wq, Wk, Wv, Wo, bo, B, T, N, D, heads, dim_head, scale,
                                                 alpha, mode, (cudaStream_t)stream);
            if (rc4 <= 0) return rc4;
        }
        if (tc2_enabled()) {
            const int rc3 = attn_fwd_tc2_dispatch(x, res, out, ln_w, ln_b, Wq, Wk, Wv, Wo, bo, B, T, N, D, heads, dim_head,
                                                  scale, alpha, mode, (cudaStream_t)stream);
            if (rc3 <= 0) return rc3;
        }
        const int rc2 = attn_fwd_tc_dispatch(x, res, out, ln_w, ln_b, Wq, Wk, Wv, Wo, bo, B, T, N, D, heads, dim_head, scale,
                                             alpha, mode, (cudaStream_t)stream);
        if (rc2 <= 0) return rc2;
    }
    AttnArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo; a.bo = bo;
    a.g.S = mode == 0 ? N : T; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.D = D; a.H = heads; a.I = heads * dim_head; a.scale = scale; a.alpha = alpha;
    int rc = plan_attn(a.g.S, D, heads, dim_head, &a.p);
    if (rc != RAT_OK) { set_error("rat_attn_fwd: sequence length %d x dim %d does not fit in shared memory", a.g.S, D); return rc; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim_head) {
        case 4: return launch_attn_fwd_p<4>(a, st);
        case 8: return launch_attn_fwd_p<8>(a, st);
        case 10: return launch_attn_fwd_p<10>(a, st);
        case 16: return launch_attn_fwd_p<16>(a, st);
        case 20: return launch_attn_fwd_p<20>(a, st);
        case 32: return launch_attn_fwd_p<32>(a, st);
        default: set_error("rat_attn_fwd: dim_head=%d not instantiated (4,8,10,16,20,32)", dim_head); return RAT_EINVAL;
    }
}

template <bool MMA>
static int launch_ff_fwd(const FFArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_fwd<MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_fwd)");
        attr_set = true;
    }
    const long long ntiles = (a.rows + a.p.RPT - 1) / a.p.RPT;
    const int grid = (int)min(ntiles, (long long)num_sms());
    k_ff_fwd<MMA><<<grid, ENC_THREADS, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_ff_fwd");
    return RAT_OK;
}

extern "C" int rat_ff_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                          const float* W1, const float* b1, const float* W2, const float* b2, long long rows, int D,
                          int M, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && M > 0, "rat_ff_fwd: bad shape");
    FFArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.rows = rows; a.D = D; a.M = M;
    if (g_precision == 2 && rr_enabled() && !getenv("RAT_RR_FF_OFF")) {
        const int rc3 = ff_fwd_rr_dispatch(x, res, out, ln_w, ln_b, W1, b1, W2, b2, rows, D, M, (cudaStream_t)stream);
        if (rc3 <= 0) return rc3;
    }
    if (g_precision == 2) {
        const int rc2 = ff_fwd_tc_dispatch(x, res, out, ln_w, ln_b, W1, b1, W2, b2, rows, D, M, (cudaStream_t)stream);
        if (rc2 <= 0) return rc2;
    }
    int rc = plan_ff(D, M, &a.p);
    if (rc != RAT_OK) { set_error("rat_ff_fwd: D=%d M=%d does not fit in shared memory", D, M); return rc; }
    return g_precision ? launch_ff_fwd<true>(a, (cudaStream_t)stream) : launch_ff_fwd<false>(a, (cudaStream_t)stream);
}

extern "C" int rat_layernorm_fwd(const float* x, float* out, const float* w, const float* b, long long rows, int D,
                                 void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0, "rat_layernorm_fwd: bad shape");
    int lg = min(32, next_pow2(D));
    int groups = 256 / lg;
    long long blocks = (rows + groups - 1) / groups;
    int grid = (int)min(blocks, (long long)num_sms() * 16);
    k_ln_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, w, b, rows, D, lg);
    RAT_CHECK_LAUNCH("k_ln_fwd");
    return RAT_OK;
}
