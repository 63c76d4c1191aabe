// K2: fused RAT-block forward kernels.
//
//   k_attn_fwd : out = res + alpha * ( MHA(LayerNorm(x)) Wo^T + bo )        PreNorm+Attention+residual
//                (reference: PreNorm RAT_m2.py:155-161, Attention RAT_m2.py:176-202, residual :224/:231;
//                 RAT_m3.Attention RAT_m3.py:164-196 with separate Wq/Wk/Wv pointers)
//   k_ff_fwd   : out = res + W2 gelu(W1 [LN](x) + b1) + b2                  FeedForward (+optional PreNorm)
//                (reference: FeedForward RAT_m2.py:163-174 ; PreNorm'd in RAT_m0.py:193-208)
//   k_ln_fwd   : out = LayerNorm(x)                                         final norm of RAT_m0/m1 Transformer
//
// One CTA owns a tile of whole sequences (SPT sequences x S positions = R token rows).  LayerNorm output, the
// per-head-chunk q|k|v, the attention output and the out-projection accumulator all stay in shared memory; the
// weights of the current head chunk are staged (transposed, k-major) into shared memory.  "Intra" attention
// (sequence = one sample row, positions = fields) and "cross" attention (sequence = one field over the 1+K
// retrieved rows) differ only in the row-index map SeqGeom::grow, i.e. the reference's
// reshape/transpose/flatten copies (RAT_m2.py:221-235) are pure indexing here.
#include "tile.cuh"
#include "../../include/rat_b200.h"

namespace rat {

struct AttnPlan {
    int SPT;      // sequences per tile
    int hc;       // heads per chunk
    int Dp;       // padded D
    int Cq;       // hc*dh
    int C3p;      // padded 3*Cq
    int lg;       // lanes per LayerNorm row group
    int lpt;      // lanes per attention task
    size_t smem_bytes;
};

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

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// LayerNorm of R rows (global -> smem), eps=1e-5, biased variance (torch.nn.LayerNorm semantics).
// If stats != nullptr the per-row (mean, rstd) are stored for the backward pass.
__device__ __forceinline__ void ln_rows_to_smem(const float* __restrict__ x, const SeqGeom& g, long long s0, int R,
                                                int D, const float* __restrict__ w, const float* __restrict__ b,
                                                float* __restrict__ dst, int ld, int lg, float* __restrict__ stats,
                                                float* __restrict__ raw, int ldraw) {
    const int groups = blockDim.x / lg;
    const int gi = threadIdx.x / lg, li = threadIdx.x % lg;
    const float invD = 1.0f / (float)D;
    for (int r0 = 0; r0 < R; r0 += groups) {
        const int r = r0 + gi;
        const bool ok = r < R;
        const float* src = x;
        if (ok) src = x + g.grow(s0 + r / g.S, r % g.S) * D;
        float sum = 0.f;
        if (ok) for (int d = li; d < D; d += lg) sum += src[d];
        const float mean = group_sum(sum, lg) * invD;
        float sq = 0.f;
        if (ok) for (int d = li; d < D; d += lg) { float t = src[d] - mean; sq = fmaf(t, t, sq); }
        const float var = group_sum(sq, lg) * invD;
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        if (ok) {
            for (int d = li; d < D; d += lg) {
                const float xv = src[d];
                dst[(size_t)r * ld + d] = (xv - mean) * rstd * w[d] + b[d];
                if (raw) raw[(size_t)r * ldraw + d] = xv;
            }
            if (stats && li == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
        }
    }
}

// stage the head-chunk weights transposed: Wt[k][c], c in [0,C3p): q cols | k cols | v cols | zero pad
__device__ __forceinline__ void stage_qkv_weights(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                  const float* __restrict__ Wv, int D, int row0, int Cq, int C3p,
                                                  float* __restrict__ Wt) {
    const int total = C3p * D;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % D, c = i / D;
        float v = 0.f;
        if (c < Cq) v = __ldg(Wq + (size_t)(row0 + c) * D + k);
        else if (c < 2 * Cq) v = __ldg(Wk + (size_t)(row0 + c - Cq) * D + k);
        else if (c < 3 * Cq) v = __ldg(Wv + (size_t)(row0 + c - 2 * Cq) * D + k);
        Wt[(size_t)k * C3p + c] = v;
    }
}
// WoT[c][d] = Wo[d][col0 + c]   (c < Cq, d < Dp, zero pad)
__device__ __forceinline__ void stage_out_weights(const float* __restrict__ Wo, int D, int I, int col0, int Cq,
                                                  int Dp, float* __restrict__ WoT) {
    const int total = Cq * Dp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % Cq, d = i / Cq;
        WoT[(size_t)c * Dp + d] = (d < D) ? __ldg(Wo + (size_t)d * I + col0 + c) : 0.f;
    }
}

// softmax(q k^T) v for every (sequence, head) task of the tile; o overwrites q. Optionally stores the
// log-sum-exp of each row (for the backward pass).  q is pre-multiplied by `scale`.
template <int DH>
__device__ __forceinline__ void attn_core(float* __restrict__ qkv, int ld, int Cq, int nseq_tile, int S, int hc,
                                          int lpt, float scale, float* __restrict__ lse) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tpw = 32 / lpt;
    const int ntasks = nseq_tile * hc;
    const int sub = lane / lpt, li = lane % lpt;
    for (int task0 = warp * tpw; task0 < ntasks; task0 += nwarps * tpw) {
        const int task = task0 + sub;
        if (task >= ntasks) continue;
        const int ls = task / hc, hl = task % hc;
        float* base = qkv + (size_t)ls * S * ld + hl * DH;
        for (int i = li; i < S; i += lpt) {
            float q[DH], acc[DH];
            float* qrow = base + (size_t)i * ld;
#pragma unroll
            for (int d = 0; d < DH; ++d) { q[d] = qrow[d] * scale; acc[d] = 0.f; }
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < S; ++j) {
                const float* krow = base + (size_t)j * ld + Cq;
                const float* vrow = krow + Cq;
                float sc = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) sc = fmaf(q[d], krow[d], sc);
                const float mn = fmaxf(m, sc);
                const float corr = expf(m - mn);
                const float pj = expf(sc - mn);
                l = fmaf(l, corr, pj);
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] = fmaf(acc[d], corr, pj * vrow[d]);
                m = mn;
            }
            const float inv = 1.0f / l;
#pragma unroll
            for (int d = 0; d < DH; ++d) qrow[d] = acc[d] * inv;
            if (lse) lse[(ls * S + i) * hc + hl] = m + logf(l);
        }
    }
}

template <int DH>
__global__ void __launch_bounds__(256) k_attn_fwd(AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const AttnPlan& p = a.p;
    const int S = a.g.S, D = a.D;
    const int Rmax = p.SPT * S;
    float* as = smem;                               // [Rmax][Dp]  LayerNorm(x)
    float* ys = as + (size_t)Rmax * p.Dp;           // [Rmax][Dp]  out-projection accumulator
    float* qkv = ys + (size_t)Rmax * p.Dp;          // [Rmax][C3p]
    float* Wt = qkv + (size_t)Rmax * p.C3p;         // [D][C3p]
    float* WoT = Wt + (size_t)D * p.C3p;            // [Cq][Dp]
    const long long ntiles = (a.nseq + p.SPT - 1) / p.SPT;
    const int nchunks = a.H / p.hc;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * p.SPT;
        const int nseq_t = (int)min((long long)p.SPT, a.nseq - s0);
        const int R = nseq_t * S;
        __syncthreads();                            // previous tile fully consumed
        ln_rows_to_smem(a.x, a.g, s0, R, D, a.ln_w, a.ln_b, as, p.Dp, p.lg, nullptr, nullptr, 0);
        for (int i = threadIdx.x; i < R * p.Dp; i += blockDim.x) ys[i] = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int row0 = ch * p.Cq;             // first q/k/v feature of this head chunk
            __syncthreads();                        // (a) previous chunk's GEMM2 done with qkv/WoT; LN visible
            stage_qkv_weights(a.Wq, a.Wk, a.Wv, D, row0, p.Cq, p.C3p, Wt);
            stage_out_weights(a.Wo, D, a.I, row0, p.Cq, p.Dp, WoT);
            __syncthreads();
            tile_gemm<8>(as, p.Dp, Wt, p.C3p, qkv, p.C3p, R, p.C3p, D, false, EpiNone());
            __syncthreads();
            attn_core<DH>(qkv, p.C3p, p.Cq, nseq_t, S, p.hc, p.lpt, a.scale, nullptr);
            __syncthreads();
            tile_gemm<8>(qkv, p.C3p, WoT, p.Dp, ys, p.Dp, R, p.Dp, p.Cq, true, EpiNone());
        }
        __syncthreads();
        for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
            const int r = i / D, d = i % D;
            const long long gr = a.g.grow(s0 + r / S, r % S);
            float v = a.alpha * (ys[(size_t)r * p.Dp + d] + a.bo[d]);
            if (a.res) v += a.res[gr * D + d];
            a.out[gr * D + d] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
struct FFPlan { int RPT; int Dp; int Mp; int lg; size_t smem_bytes; };
struct FFArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;               // nullptr => no pre-norm (RAT_m2/m3)
    const float* W1; const float* b1; const float* W2; const float* b2;   // [M,D],[M],[D,M],[D]
    long long rows;
    int D, M;
    FFPlan p;
};

struct EpiBiasGelu {
    const float* b;
    __device__ __forceinline__ void operator()(int, int c0, float4& v) const {
        v.x = gelu_erf(v.x + b[c0]); v.y = gelu_erf(v.y + b[c0 + 1]);
        v.z = gelu_erf(v.z + b[c0 + 2]); v.w = gelu_erf(v.w + b[c0 + 3]);
    }
};

// stage Wt[k][c] = W[c][k] for a torch Linear weight W [C, K]; zero pad c >= C
__device__ __forceinline__ void stage_linear_T(const float* __restrict__ W, int C, int K, int Cp,
                                               float* __restrict__ Wt) {
    const int total = Cp * K;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % K, c = i / K;
        Wt[(size_t)k * Cp + c] = (c < C) ? __ldg(W + (size_t)c * K + k) : 0.f;
    }
}

__global__ void __launch_bounds__(256) k_ff_fwd(FFArgs a) {
    extern __shared__ __align__(16) float smem[];
    const FFPlan& p = a.p;
    const int D = a.D, M = a.M;
    float* xs = smem;                                   // [RPT][Dp]
    float* hs = xs + (size_t)p.RPT * p.Dp;              // [RPT][Mp]
    float* ys = hs + (size_t)p.RPT * p.Mp;              // [RPT][Dp]
    float* W1t = ys + (size_t)p.RPT * p.Dp;             // [D][Mp]
    float* W2t = W1t + (size_t)D * p.Mp;                // [M][Dp]
    float* b1s = W2t + (size_t)M * p.Dp;                // [Mp]
    stage_linear_T(a.W1, M, D, p.Mp, W1t);
    stage_linear_T(a.W2, D, M, p.Dp, W2t);
    for (int i = threadIdx.x; i < p.Mp; i += blockDim.x) b1s[i] = i < M ? a.b1[i] : 0.f;
    const long long ntiles = (a.rows + p.RPT - 1) / p.RPT;
    SeqGeom flat{1, 0, 1, 1};
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * p.RPT;
        const int R = (int)min((long long)p.RPT, a.rows - r0);
        __syncthreads();
        if (a.ln_w) {
            ln_rows_to_smem(a.x, flat, r0, R, D, a.ln_w, a.ln_b, xs, p.Dp, p.lg, nullptr, nullptr, 0);
        } else {
            for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
                const int r = i / D, d = i % D;
                xs[(size_t)r * p.Dp + d] = a.x[(r0 + r) * D + d];
            }
        }
        __syncthreads();
        tile_gemm<8>(xs, p.Dp, W1t, p.Mp, hs, p.Mp, R, p.Mp, D, false, EpiBiasGelu{b1s});
        __syncthreads();
        tile_gemm<8>(hs, p.Mp, W2t, p.Dp, ys, p.Dp, R, p.Dp, M, false, EpiNone());
        __syncthreads();
        for (int i = threadIdx.x; i < R * D; i += blockDim.x) {
            const int r = i / D, d = i % D;
            a.out[(r0 + r) * D + d] = a.res[(r0 + r) * D + d] + ys[(size_t)r * p.Dp + d] + a.b2[d];
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
// Choose (heads per chunk, sequences per tile): as many token rows per tile as fit (target 128..176), then the
// widest head chunk.  per_row_extra / fixed_extra (floats) let the backward kernel reserve its extra buffers:
// fixed_extra is counted per head-chunk column set by the caller through the callback-free formula below.
int plan_attn_ex(int S, int D, int H, int dh, int row_mul_D, int row_mul_C3, int row_mul_Cq, int row_extra,
                 int fix_mul_DC3, int fix_mul_CqD, int fix_extra, size_t budget, AttnPlan* out) {
    const int Dp = round_up(D, 4);
    const int cap_spt = max(1, 176 / S);
    AttnPlan best{};
    int bestR = 0;
    for (int pass = 0; pass < 2 && bestR == 0; ++pass) {
        const size_t bud = (pass == 0 ? budget : (size_t)max_smem_optin()) / 4;
        for (int hc = H; hc >= 1; --hc) {
            if (H % hc) continue;
            const int Cq = hc * dh, C3p = round_up(3 * Cq, 4), Cqp = round_up(Cq, 4);
            const size_t per_row = (size_t)row_mul_D * Dp + (size_t)row_mul_C3 * C3p + (size_t)row_mul_Cq * Cqp + row_extra * hc;
            const size_t fixed = (size_t)fix_mul_DC3 * D * C3p + (size_t)fix_mul_CqD * Cqp * Dp + fix_extra;
            if (fixed + per_row * S > bud) continue;
            int spt = (int)min((size_t)cap_spt, (bud - fixed) / (per_row * S));
            if (spt < 1) continue;
            const int R = spt * S;
            if (R > bestR) {
                bestR = R;
                best.SPT = spt; best.hc = hc; best.Dp = Dp; best.Cq = Cq; best.C3p = C3p;
                best.smem_bytes = (fixed + per_row * R) * 4;
            }
            if (R >= min(128, cap_spt * S)) break;
        }
    }
    if (bestR == 0) return RAT_ESMEM;
    best.lg = min(32, next_pow2(D));
    best.lpt = min(32, next_pow2(S));
    *out = best;
    return RAT_OK;
}

int plan_attn(int S, int D, int H, int dh, AttnPlan* out) {
    // forward: as + ys (2 x Dp per row), qkv (C3p per row); Wt [D][C3p], WoT [Cq][Dp]
    return plan_attn_ex(S, D, H, dh, 2, 1, 0, 0, 1, 1, 0, 100 * 1024, out);
}

}  // namespace rat

using namespace rat;

template <int DH>
static int launch_attn_fwd(const AttnArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd)");
        attr_set = true;
    }
    const long long ntiles = (a.nseq + a.p.SPT - 1) / a.p.SPT;
    int per_sm = max(1, (int)(220 * 1024 / (a.p.smem_bytes + 1024)));
    if (per_sm > 4) per_sm = 4;
    int grid = (int)min(ntiles, (long long)num_sms() * per_sm);
    k_attn_fwd<DH><<<grid, 256, a.p.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd");
    return RAT_OK;
}

extern "C" int rat_attn_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                            const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo,
                            int B, int T, int N, int D, int heads, int dim_head, float scale, float alpha, int mode,
                            void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && N > 0 && D > 0 && heads > 0, "rat_attn_fwd: bad shape");
    RAT_REQUIRE(mode == 0 || mode == 1, "rat_attn_fwd: mode must be 0 (intra) or 1 (cross)");
    RAT_REQUIRE(Wo != nullptr && bo != nullptr, "rat_attn_fwd: identity out-projection (heads==1 && dim_head==dim) is not supported");
    AttnArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo; a.bo = bo;
    a.g.S = mode == 0 ? N : T; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    a.D = D; a.H = heads; a.I = heads * dim_head; a.scale = scale; a.alpha = alpha;
    int rc = plan_attn(a.g.S, D, heads, dim_head, &a.p);
    if (rc != RAT_OK) { set_error("rat_attn_fwd: sequence length %d x dim %d does not fit in shared memory", a.g.S, D); return rc; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim_head) {
        case 4: return launch_attn_fwd<4>(a, st);
        case 8: return launch_attn_fwd<8>(a, st);
        case 10: return launch_attn_fwd<10>(a, st);
        case 16: return launch_attn_fwd<16>(a, st);
        case 20: return launch_attn_fwd<20>(a, st);
        case 32: return launch_attn_fwd<32>(a, st);
        default: set_error("rat_attn_fwd: dim_head=%d not instantiated (4,8,10,16,20,32)", dim_head); return RAT_EINVAL;
    }
}

namespace rat {
int plan_ff(int D, int M, FFPlan* out, int extra_row_floats) {
    FFPlan p{};
    p.Dp = round_up(D, 4); p.Mp = round_up(M, 4);
    p.lg = min(32, next_pow2(D));
    const size_t budget = 100 * 1024;
    const size_t wfl = (size_t)D * p.Mp + (size_t)M * p.Dp + p.Mp;
    int rpt = 192;
    for (; rpt >= 8; rpt -= 8) {
        size_t fl = (size_t)rpt * (2 * p.Dp + p.Mp + extra_row_floats) + wfl;
        if (fl * 4 <= budget) break;
    }
    if (rpt < 8) {
        rpt = 8;
        size_t fl = (size_t)rpt * (2 * p.Dp + p.Mp + extra_row_floats) + wfl;
        if (fl * 4 > (size_t)max_smem_optin()) return RAT_ESMEM;
    }
    p.RPT = rpt;
    p.smem_bytes = ((size_t)rpt * (2 * p.Dp + p.Mp + extra_row_floats) + wfl) * 4;
    *out = p;
    return RAT_OK;
}
}  // namespace rat

extern "C" int rat_ff_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                          const float* W1, const float* b1, const float* W2, const float* b2, long long rows, int D,
                          int M, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && M > 0, "rat_ff_fwd: bad shape");
    FFArgs a{};
    a.x = x; a.res = res; a.out = out; a.ln_w = ln_w; a.ln_b = ln_b; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.rows = rows; a.D = D; a.M = M;
    int rc = plan_ff(D, M, &a.p, 0);
    if (rc != RAT_OK) { set_error("rat_ff_fwd: D=%d M=%d does not fit in shared memory", D, M); return rc; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin());
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_fwd)");
        attr_set = true;
    }
    const long long ntiles = (rows + a.p.RPT - 1) / a.p.RPT;
    int per_sm = max(1, (int)(220 * 1024 / (a.p.smem_bytes + 1024)));
    if (per_sm > 4) per_sm = 4;
    int grid = (int)min(ntiles, (long long)num_sms() * per_sm);
    k_ff_fwd<<<grid, 256, a.p.smem_bytes, (cudaStream_t)stream>>>(a);
    RAT_CHECK_LAUNCH("k_ff_fwd");
    return RAT_OK;
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
