// FeedForward forward, register-resident generation (see encoder_rr.cuh):  out = res + W2 gelu_erf(W1 LNopt(x) + b1) + b2
// over `rows` tokens.  Replaces FeedForward (models/RAT_m2.py:163-174; with ln_w / ln_b the PreNorm of RAT_m0.py:197-201).
// One warp owns 16 consecutive token rows from the (bulk-copied) load to the store:
//   xa = LNopt(x) as mma.sync A fragments ; for every pair of 8-wide hidden n-tiles p:
//       h_p [16 x 16] = xa . W1_p^T (KS HMMA pairs) ; + b1 ; exact-erf GELU in the accumulator registers ;
//       cvt.f16x2 -> that IS k-step p of the second product:  acc [16 x D] += gelu(h_p) . W2[:, p]^T
//   so the hidden activation never exists as a whole, not even in registers (8 fp32 at a time).
// Weights live in shared memory in fragment order (15 KB at D = 40, M = 80).
#include "encoder_rr.cuh"

namespace rat {

struct FFRRArgs {
    const float* x; const float* res; float* out;
    const float* ln_w; const float* ln_b;
    const float* W1; const float* b1; const float* W2; const float* b2;
    long long rows;
    int D, M;
};

// KS = pad16(D) / 16, NTO = ceil(D / 8), KS2 = pad16(M) / 16 (pairs of hidden n-tiles = k-steps of the second product)
template <int KS, int NTO, int KS2, int WARPS, int CTAS, bool BULK>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_ff_fwd_rr(FFRRArgs a) {
    extern __shared__ __align__(16) uint4 ffr_smem[];
    constexpr int NP = (NTO + 1) / 2;
    const int D = a.D, M = a.M;
    uint4* W1_i = ffr_smem;                                // [KS2][KS][32]
    uint4* W2_i = W1_i + KS2 * KS * 32;                    // [NP][KS2][32]
    float* lnw_s = reinterpret_cast<float*>(W2_i + NP * KS2 * 32);   // [KS * 16]
    float* lnb_s = lnw_s + KS * 16;
    float* b1_s = lnb_s + KS * 16;                         // [KS2 * 16]
    float* b2_s = b1_s + KS2 * 16;                         // [NP * 16]
    float* stage = b2_s + NP * 16;                         // BULK: [WARPS][2][16 rows][D]
    __shared__ __align__(8) uint64_t row_bar[WARPS][2];
    const bool prenorm = a.ln_w != nullptr;
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < KS2 * KS * 32; i += blockDim.x) {
        const int p = i / (KS * 32), ks = (i >> 5) % KS, ln = i & 31;
        W1_i[i] = frag_pair_entry(ln, 16 * p, 16 * ks, [&](int m, int c) { return (m < M && c < D) ? __ldg(a.W1 + (size_t)m * D + c) : 0.f; });
    }
    for (int i = threadIdx.x; i < NP * KS2 * 32; i += blockDim.x) {
        const int p = i / (KS2 * 32), ks2 = (i >> 5) % KS2, ln = i & 31;
        W2_i[i] = frag_pair_entry(ln, 16 * p, 16 * ks2, [&](int c, int m) { return (c < D && m < M) ? __ldg(a.W2 + (size_t)c * M + m) : 0.f; });
    }
    for (int i = threadIdx.x; i < KS * 16; i += blockDim.x) {
        lnw_s[i] = (prenorm && i < D) ? a.ln_w[i] : 0.f;
        lnb_s[i] = (prenorm && i < D) ? a.ln_b[i] : 0.f;
    }
    for (int i = threadIdx.x; i < KS2 * 16; i += blockDim.x) b1_s[i] = i < M ? a.b1[i] : 0.f;
    for (int i = threadIdx.x; i < NP * 16; i += blockDim.x) b2_s[i] = i < D ? a.b2[i] : 0.f;
    if (threadIdx.x < WARPS * 2) tc5::mbar_init(&row_bar[0][0] + threadIdx.x, 1);
    tc5::fence_mbar_init();
    __syncthreads();
    pdl_wait();                                            // x / res are written by the previous kernel of the stream
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    const float invD = 1.0f / (float)D;
    const long long ntasks = (a.rows + 15) >> 4;
    const long long wstride = (long long)gridDim.x * WARPS;
    const bool res_is_x = a.res == a.x;
    float* my_stage = stage + (size_t)warp * 2 * 16 * D;
    auto issue_rows = [&](long long task, int buf) {           // 16 consecutive rows = one contiguous copy
        if (lane == 0) {
            const int nr = (int)min(16LL, a.rows - task * 16);
            mbar_arrive_expect_tx(&row_bar[warp][buf], (uint32_t)(nr * D * 4));
            bulk_g2s(tc5::smem_u32(my_stage + (size_t)buf * 16 * D), a.x + task * 16 * D, (uint32_t)(nr * D * 4), &row_bar[warp][buf]);
        }
    };
    long long task = (long long)blockIdx.x * WARPS + warp;
    uint32_t ph0 = 0, ph1 = 0;
    int buf = 0;
    if (BULK && task < ntasks) issue_rows(task, 0);
    for (; task < ntasks; task += wstride, buf ^= 1) {
        const long long rlo = task * 16 + g, rhi = rlo + 8;
        const bool vlo = rlo < a.rows, vhi = rhi < a.rows;
        const float* slo = my_stage + (size_t)(buf * 16 + g) * D;
        const float* shi = slo + 8 * D;
        uint32_t xa[KS][4];
        {
            float2 xl[NTO], xh[NTO];
            if (BULK) {
                if (task + wstride < ntasks) issue_rows(task + wstride, buf ^ 1);
                tc5::mbar_wait(&row_bar[warp][buf], buf ? ph1 : ph0);
                if (buf) ph1 ^= 1; else ph0 ^= 1;
                rr_load_rows<NTO>(slo, shi, vlo, vhi, D, t, xl, xh);
            } else {
                rr_load_rows<NTO>(a.x + rlo * D, a.x + rhi * D, vlo, vhi, D, t, xl, xh);
            }
            float ml = 0.f, rl = 1.f, mh = 0.f, rh = 1.f;
            if (prenorm) {
                rr_row_stats<NTO>(xl, D, invD, t, ml, rl);
                rr_row_stats<NTO>(xh, D, invD, t, mh, rh);
            }
#pragma unroll
            for (int nt = 0; nt < 2 * KS; ++nt) {
                uint32_t lo = 0u, hi = 0u;
                if (nt < NTO) {
                    if (prenorm) {
                        const float2 w = *reinterpret_cast<const float2*>(lnw_s + 8 * nt + 2 * t);
                        const float2 b = *reinterpret_cast<const float2*>(lnb_s + 8 * nt + 2 * t);
                        lo = vlo ? pack_h2(fmaf((xl[nt].x - ml) * rl, w.x, b.x), fmaf((xl[nt].y - ml) * rl, w.y, b.y)) : 0u;
                        hi = vhi ? pack_h2(fmaf((xh[nt].x - mh) * rh, w.x, b.x), fmaf((xh[nt].y - mh) * rh, w.y, b.y)) : 0u;
                    } else {
                        lo = pack_h2(xl[nt].x, xl[nt].y);
                        hi = pack_h2(xh[nt].x, xh[nt].y);
                    }
                }
                xa[nt >> 1][(nt & 1) * 2] = lo;
                xa[nt >> 1][(nt & 1) * 2 + 1] = hi;
            }
        }
        float acc[NTO][4];
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int p = 0; p < KS2; ++p) {
            float h[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 f = W1_i[(p * KS + ks) * 32 + lane];
                rr_mma(h[0], xa[ks], f.x, f.y);
                rr_mma(h[1], xa[ks], f.z, f.w);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 b = *reinterpret_cast<const float2*>(b1_s + 16 * p + 8 * j + 2 * t);
                float dg;
                gelu_fast(h[j][0] + b.x, h[j][0], dg); gelu_fast(h[j][1] + b.y, h[j][1], dg);
                gelu_fast(h[j][2] + b.x, h[j][2], dg); gelu_fast(h[j][3] + b.y, h[j][3], dg);
            }
            // hidden columns >= M: W1 image rows and b1 are zero -> gelu(0) = 0
            uint32_t ha[4];
            c_to_a(h, ha);
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                const uint4 f = W2_i[(pp * KS2 + p) * 32 + lane];
                rr_mma(acc[2 * pp], ha, f.x, f.y);
                if (2 * pp + 1 < NTO) rr_mma(acc[2 * pp + 1], ha, f.z, f.w);
            }
        }
        {
            const bool from_stage = BULK && res_is_x;
            const float* rl_p = from_stage ? slo : a.res + rlo * D;
            const float* rh_p = from_stage ? shi : a.res + rhi * D;
            float* ol = a.out + rlo * D;
            float* oh = a.out + rhi * D;
#pragma unroll
            for (int nt = 0; nt < NTO; ++nt) {
                const int c = 8 * nt + 2 * t;
                if (c < D) {
                    const float2 b = *reinterpret_cast<const float2*>(b2_s + c);
                    if (vlo) {
                        float2 r = *reinterpret_cast<const float2*>(rl_p + c);
                        r.x += acc[nt][0] + b.x; r.y += acc[nt][1] + b.y;
                        *reinterpret_cast<float2*>(ol + c) = r;
                    }
                    if (vhi) {
                        float2 r = *reinterpret_cast<const float2*>(rh_p + c);
                        r.x += acc[nt][2] + b.x; r.y += acc[nt][3] + b.y;
                        *reinterpret_cast<float2*>(oh + c) = r;
                    }
                }
            }
        }
        if (BULK) __syncwarp();
    }
}

template <int KS, int NTO, int KS2, bool BULK, int WARPS, int CTAS>
static int launch_ff_fwd_rr_w(const FFRRArgs& a, cudaStream_t st) {
    constexpr int NP = (NTO + 1) / 2;
    const size_t smem = ((size_t)KS2 * KS * 32 + (size_t)NP * KS2 * 32) * sizeof(uint4) + (size_t)(2 * KS * 16 + KS2 * 16 + NP * 16) * 4 +
                        (size_t)(BULK ? WARPS * 2 * 16 * a.D * 4 : 0);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_fwd_rr<KS, NTO, KS2, WARPS, CTAS, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_fwd_rr)");
        attr_smem = smem;
    }
    const long long ntasks = (a.rows + 15) / 16;
    const long long nblk = (ntasks + WARPS - 1) / WARPS;
    const int grid = (int)std::min<long long>(nblk, (long long)CTAS * num_sms());
    if (launch_pdl(k_ff_fwd_rr<KS, NTO, KS2, WARPS, CTAS, BULK>, dim3(grid), dim3(WARPS * 32), smem, st, a) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "k_ff_fwd_rr");
    RAT_CHECK_LAUNCH("k_ff_fwd_rr");
    return RAT_OK;
}
template <int KS, int NTO, int KS2, bool BULK>
static int launch_ff_fwd_rr_v(const FFRRArgs& a, cudaStream_t st) {
    static int variant = -1;     // RAT_RR_FF_VARIANT (tuning aid): 1 = 16 warps x 1 CTA per SM instead of 12 x 2
    if (variant < 0) { const char* e = getenv("RAT_RR_FF_VARIANT"); variant = e ? atoi(e) : 0; }
    return variant == 1 ? launch_ff_fwd_rr_w<KS, NTO, KS2, BULK, 16, 1>(a, st) : launch_ff_fwd_rr_w<KS, NTO, KS2, BULK, 12, 2>(a, st);
}
template <int KS, int NTO>
static int launch_ff_fwd_rr(const FFRRArgs& a, cudaStream_t st) {
    const bool bulk = (a.D % 4) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    const int ks2 = pad16(a.M) / 16;
#define RAT_FFR(K2_) (bulk ? launch_ff_fwd_rr_v<KS, NTO, K2_, true>(a, st) : launch_ff_fwd_rr_v<KS, NTO, K2_, false>(a, st))
    switch (ks2) {
        case 1: return RAT_FFR(1);
        case 2: return RAT_FFR(2);
        case 3: return RAT_FFR(3);
        case 5: return RAT_FFR(5);
        case 6: return RAT_FFR(6);
        default: return 1;
    }
#undef RAT_FFR
}


// ------------------------------------------------------------------------------------------------ FeedForward backward
// Backward of  out = res + W2 gelu(W1 x + b1) + b2  (no pre-norm: the RAT_m2 / RAT_m3 FeedForward):
//   z = x W1^T + b1 (recomputed) ; dh = dy W2 ; dz = dh * gelu'(z) ; dx = base + dz W1
//   gW1 = dz^T x ; gb1 = colsum(dz) (a column of ones in the x tile) ; gW2^T = h^T dy ; gb2 = colsum(dy)
// 8 warps, one CTA per SM.  Per warp task (16 rows), per pair p of hidden n-tiles, all on mma.sync fragments:
//   z_p, dh_p (2 KS HMMA pairs each) -> gelu, gelu' in registers -> dz_p packed = k-step p of  dx += dz_p . W1_p ;
//   h_p and dz_p go as fp16 rows into the CTA's 128-row token tile, next to x (+ ones column) and dy.
// The LAST warp to finish its rows issues the tile's two tcgen05 weight-gradient products (MN-major views of the token
// tile, K = 128 tokens, accumulators in TMEM for the whole kernel); dx never waits for the tensor core.
struct FFBwdRRArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* W1; const float* b1; const float* W2;
    float* partials; const float* dout_amax; float* dx_amax;
    long long rows;
    int D, M, psize, smem_bytes;
};
constexpr int FFB_WARPS = 8;

template <int KS, int NTO, int KS2>
__global__ void __launch_bounds__(FFB_WARPS * 32, 1) k_ff_bwd_rr(FFBwdRRArgs a) {
    extern __shared__ __align__(128) unsigned char ffb_smem[];
    constexpr int Kp = 16 * KS, KC1 = 2 * KS, Mp = 16 * KS2, NP = (NTO + 1) / 2;
    const int D = a.D, M = a.M;
    uint4* W1_i = reinterpret_cast<uint4*>(ffb_smem);      // z = x W1^T        (n = m, k = c)    [KS2][KS][32]
    uint4* W2T_i = W1_i + KS2 * KS * 32;                   // dh = dy W2        (n = m, k = c)    [KS2][KS][32]
    uint4* W1T_i = W2T_i + KS2 * KS * 32;                  // dx = dz W1        (n = c, k = m)    [NP][KS2][32]
    float* b1_s = reinterpret_cast<float*>(W1T_i + NP * KS2 * 32);        // [Mp]
    unsigned char* Xt = reinterpret_cast<unsigned char*>(b1_s + Mp);      // [128 x Kp]  x, column D = 1
    unsigned char* DYt = Xt + (size_t)KC1 * tc5::TILE_CHUNK;               // [128 x Kp]  gs * dout
    unsigned char* Zt = DYt + (size_t)KC1 * tc5::TILE_CHUNK;               // [128 x 128] dz   (columns >= Mp stay zero)
    unsigned char* Ht = Zt + (size_t)16 * tc5::TILE_CHUNK;                 // [128 x 128] h
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int arrive_cnt;
    __shared__ float red[FFB_WARPS][Kp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < KS2 * KS * 32; i += blockDim.x) {
        const int p = i / (KS * 32), ks = (i >> 5) % KS, ln = i & 31;
        W1_i[i] = frag_pair_entry(ln, 16 * p, 16 * ks, [&](int m, int c) { return (m < M && c < D) ? __ldg(a.W1 + (size_t)m * D + c) : 0.f; });
        W2T_i[i] = frag_pair_entry(ln, 16 * p, 16 * ks, [&](int m, int c) { return (m < M && c < D) ? __ldg(a.W2 + (size_t)c * M + m) : 0.f; });
    }
    for (int i = threadIdx.x; i < NP * KS2 * 32; i += blockDim.x) {
        const int p = i / (KS2 * 32), ks2 = (i >> 5) % KS2, ln = i & 31;
        W1T_i[i] = frag_pair_entry(ln, 16 * p, 16 * ks2, [&](int c, int m) { return (c < D && m < M) ? __ldg(a.W1 + (size_t)m * D + c) : 0.f; });
    }
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) b1_s[i] = i < M ? a.b1[i] : 0.f;
    {
        const int tile16 = (2 * KC1 + 32) * (int)tc5::TILE_CHUNK / 16;
        for (int i = threadIdx.x; i < tile16; i += blockDim.x) reinterpret_cast<uint4*>(Xt)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) { tc5::mbar_init(&done_bar, 1); tc5::fence_mbar_init(); arrive_cnt = 0u; }
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 128);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    pdl_wait();                                            // x / dout / base / the amax slot come from the previous kernels
    const float gs = tc_grad_scale(a.dout_amax), inv_gs = 1.0f / gs;
    const uint32_t tmem_W1 = tmem_base_s, tmem_W2 = tmem_base_s + Kp;      // gW1 [m][c] (c = D: gb1), gW2^T [m][c]
    const uint32_t Xt_s = tc5::smem_u32(Xt), DYt_s = tc5::smem_u32(DYt), Zt_s = tc5::smem_u32(Zt), Ht_s = tc5::smem_u32(Ht);
    const uint32_t idesc_w = tc5::instr_desc(TC_FMT, 128, Kp, 1, 1);
    const long long ntasks = (a.rows + 15) >> 4;
    const long long ntiles = (ntasks + FFB_WARPS - 1) / FFB_WARPS;
    const uint32_t row_lo = (uint32_t)(warp * 16 + g) * 16u, row_hi = row_lo + 128u;
    float dx_max = 0.f;
    float acc_b2[NTO][2];
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) acc_b2[nt][0] = acc_b2[nt][1] = 0.f;
    uint32_t dph = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long task = tile * FFB_WARPS + warp;
        const long long rlo = task * 16 + g, rhi = rlo + 8;
        const bool vlo = rlo < a.rows, vhi = rhi < a.rows;
        uint32_t xa[KS][4], da[KS][4];
        {
            float2 xl[NTO], xh[NTO], dl[NTO], dh2[NTO];
            rr_load_rows<NTO>(a.x + rlo * D, a.x + rhi * D, vlo, vhi, D, t, xl, xh);
            rr_load_rows<NTO>(a.dout + rlo * D, a.dout + rhi * D, vlo, vhi, D, t, dl, dh2);
            if (tile + gridDim.x < ntiles && lane < 16) {         // next tile's rows towards L2
                const long long r1 = (task + (long long)gridDim.x * FFB_WARPS) * 16 + lane;
                if (r1 < a.rows) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + r1 * D));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dout + r1 * D));
                    if (D > 32) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + r1 * D + 32)); asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dout + r1 * D + 32)); }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2 * KS; ++nt) {
                uint32_t lo = 0u, hi = 0u, dlo = 0u, dhi = 0u;
                if (nt < NTO) {
                    lo = pack_h2(xl[nt].x, xl[nt].y); hi = pack_h2(xh[nt].x, xh[nt].y);
                    const float a0 = dl[nt].x * gs, a1 = dl[nt].y * gs, b0 = dh2[nt].x * gs, b1 = dh2[nt].y * gs;
                    dlo = pack_h2(a0, a1); dhi = pack_h2(b0, b1);
                    acc_b2[nt][0] += a0 + b0; acc_b2[nt][1] += a1 + b1;
                }
                xa[nt >> 1][(nt & 1) * 2] = lo; xa[nt >> 1][(nt & 1) * 2 + 1] = hi;
                da[nt >> 1][(nt & 1) * 2] = dlo; da[nt >> 1][(nt & 1) * 2 + 1] = dhi;
            }
        }
        // the previous tile's weight-gradient products have read the token tile
        if (it > 0) { tc5::mbar_wait(&done_bar, dph); dph ^= 1; }
        // ---- x (+ ones column D, valid rows only) and dy rows of the token tile
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t c0 = (uint32_t)(2 * ks) * tc5::TILE_CHUNK + 4u * t, c1 = c0 + tc5::TILE_CHUNK;
            uint32_t x0 = xa[ks][0], x1 = xa[ks][1], x2 = xa[ks][2], x3 = xa[ks][3];
            if (16 * ks + 2 * t == D) { x0 = vlo ? (TC_ONES2 & 0xffffu) : 0u; x1 = vhi ? (TC_ONES2 & 0xffffu) : 0u; }
            if (16 * ks + 8 + 2 * t == D) { x2 = vlo ? (TC_ONES2 & 0xffffu) : 0u; x3 = vhi ? (TC_ONES2 & 0xffffu) : 0u; }
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c0 + row_lo), "r"(x0) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c0 + row_hi), "r"(x1) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c1 + row_lo), "r"(x2) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(Xt_s + c1 + row_hi), "r"(x3) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c0 + row_lo), "r"(da[ks][0]) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c0 + row_hi), "r"(da[ks][1]) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c1 + row_lo), "r"(da[ks][2]) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(DYt_s + c1 + row_hi), "r"(da[ks][3]) : "memory");
        }
        float acc[NTO][4];
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int p = 0; p < KS2; ++p) {
            float z[2][4] = {}, dh[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 f1 = W1_i[(p * KS + ks) * 32 + lane], f2 = W2T_i[(p * KS + ks) * 32 + lane];
                rr_mma(z[0], xa[ks], f1.x, f1.y);
                rr_mma(z[1], xa[ks], f1.z, f1.w);
                rr_mma(dh[0], da[ks], f2.x, f2.y);
                rr_mma(dh[1], da[ks], f2.z, f2.w);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 b = *reinterpret_cast<const float2*>(b1_s + 16 * p + 8 * j + 2 * t);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float gv, dg;
                    gelu_fast(z[j][e] + ((e & 1) ? b.y : b.x), gv, dg);
                    // rows that do not exist: x = 0 gives h = gelu(b1) != 0, which must not reach gW2 (dy = 0 there keeps gW1 clean)
                    z[j][e] = ((e < 2) ? vlo : vhi) ? gv : 0.f;          // h
                    dh[j][e] *= dg;                                        // dz
                }
            }
            uint32_t ha[4], za[4];
            c_to_a(z, ha);
            c_to_a(dh, za);
            // fp16 rows of the token tile: columns 16 p + {2t, 2t+1} (chunk 2p) and 16 p + 8 + {2t, 2t+1} (chunk 2p + 1)
            {
                const uint32_t c0 = (uint32_t)(2 * p) * tc5::TILE_CHUNK + 4u * t, c1 = c0 + tc5::TILE_CHUNK;
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Ht_s + c0 + row_lo), "r"(ha[0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Ht_s + c0 + row_hi), "r"(ha[1]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Ht_s + c1 + row_lo), "r"(ha[2]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Ht_s + c1 + row_hi), "r"(ha[3]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Zt_s + c0 + row_lo), "r"(za[0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Zt_s + c0 + row_hi), "r"(za[1]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Zt_s + c1 + row_lo), "r"(za[2]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(Zt_s + c1 + row_hi), "r"(za[3]) : "memory");
            }
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                const uint4 f = W1T_i[(pp * KS2 + p) * 32 + lane];
                rr_mma(acc[2 * pp], za, f.x, f.y);
                if (2 * pp + 1 < NTO) rr_mma(acc[2 * pp + 1], za, f.z, f.w);
            }
        }
        // ---- this warp's rows are complete; the last of the 8 warps issues the tile's weight-gradient products
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        __threadfence_block();
        __syncwarp();
        unsigned int cnt = 0u;
        if (lane == 0) cnt = atomicAdd(&arrive_cnt, 1u);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        if ((cnt & (FFB_WARPS - 1)) == FFB_WARPS - 1) {
            __threadfence_block();
            tc5::fence_after_sync();
            for (int j = 0; j < 8; ++j) {        // K = 128 token rows in steps of 16; A = dz / h (M = 128 hidden columns), B = x / dy
                tc5::mma_f16_w(tmem_W1, tc5::smem_desc(Zt_s + j * 256, 128u, tc5::TILE_CHUNK), tc5::smem_desc(Xt_s + j * 256, 128u, tc5::TILE_CHUNK),
                               idesc_w, (it > 0 || j > 0) ? 1u : 0u);
                tc5::mma_f16_w(tmem_W2, tc5::smem_desc(Ht_s + j * 256, 128u, tc5::TILE_CHUNK), tc5::smem_desc(DYt_s + j * 256, 128u, tc5::TILE_CHUNK),
                               idesc_w, (it > 0 || j > 0) ? 1u : 0u);
            }
            tc5::mma_commit_w(&done_bar);
        }
        // ---- dx = base + dz W1 / gs
        {
            const float* bl = a.base ? a.base + rlo * D : nullptr;
            const float* bh = a.base ? a.base + rhi * D : nullptr;
#pragma unroll
            for (int nt = 0; nt < NTO; ++nt) {
                const int c = 8 * nt + 2 * t;
                if (c < D) {
                    if (vlo) {
                        float2 r = bl ? *reinterpret_cast<const float2*>(bl + c) : make_float2(0.f, 0.f);
                        r.x = fmaf(acc[nt][0], inv_gs, r.x); r.y = fmaf(acc[nt][1], inv_gs, r.y);
                        dx_max = fmaxf(dx_max, fmaxf(fabsf(r.x), fabsf(r.y)));
                        *reinterpret_cast<float2*>(a.dx + rlo * D + c) = r;
                    }
                    if (vhi) {
                        float2 r = bh ? *reinterpret_cast<const float2*>(bh + c) : make_float2(0.f, 0.f);
                        r.x = fmaf(acc[nt][2], inv_gs, r.x); r.y = fmaf(acc[nt][3], inv_gs, r.y);
                        dx_max = fmaxf(dx_max, fmaxf(fabsf(r.x), fabsf(r.y)));
                        *reinterpret_cast<float2*>(a.dx + rhi * D + c) = r;
                    }
                }
            }
        }
    }
    if (it > 0) { tc5::mbar_wait(&done_bar, dph); dph ^= 1; tc5::fence_after_sync(); }
    // ---- per-CTA record: [Mp][Kp] gW1 (column D = gb1) | [Mp][Kp] gW2^T | [Kp] gb2
    {
        float* rec = a.partials + (size_t)blockIdx.x * a.psize;
        const int erow = (warp & 3) * 32 + lane, ehalf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll
        for (int which = 0; which < 2; ++which)
#pragma unroll
            for (int u = 0; u < KS; ++u) {
                const int gq = ehalf * KS + u;
                float v[8];
                tc5::tmem_ld8((which ? tmem_W2 : tmem_W1) + lane_base + gq * 8, v);
                tc5::tmem_ld_wait();
                if (erow < Mp && it > 0) {
                    float* dst = rec + (size_t)which * Mp * Kp + (size_t)erow * Kp + gq * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0] * inv_gs, v[1] * inv_gs, v[2] * inv_gs, v[3] * inv_gs);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * inv_gs, v[5] * inv_gs, v[6] * inv_gs, v[7] * inv_gs);
                }
            }
#pragma unroll
        for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float v = acc_b2[nt][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0) red[warp][8 * nt + 2 * t + e] = v;
            }
        __syncthreads();
        if (threadIdx.x < Kp) {
            float s = 0.f;
            if (threadIdx.x < 8 * NTO)
#pragma unroll
                for (int w = 0; w < FFB_WARPS; ++w) s += red[w][threadIdx.x];
            rec[(size_t)2 * Mp * Kp + threadIdx.x] = s * inv_gs;
        }
    }
    publish_amax_block(a.dx_amax, dx_max);
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem_base_s, 128);
}

static bool ff_bwd_rr_plan(int D, int M, FFBwdRRArgs* a) {
    if (D < 2 || (D & 1) || D > 48 || M < 1 || M > 96) return false;
    const int Kp = pad16(D), Mp = pad16(M), KS = Kp / 16, KS2 = Mp / 16, NP = ((D + 7) / 8 + 1) / 2;
    if (Kp == D) return false;                       // the ones column (gb1) needs a pad column of the x tile
    if (KS2 == 4) return false;                      // not instantiated
    a->D = D; a->M = M;
    a->psize = 2 * Mp * Kp + Kp;
    a->smem_bytes = (int)(((size_t)2 * KS2 * KS * 32 + (size_t)NP * KS2 * 32) * 16 + (size_t)Mp * 4 + (size_t)(4 * KS + 32) * tc5::TILE_CHUNK);
    return a->smem_bytes <= max_smem_optin() - 4096;
}
static int ff_bwd_rr_grid(long long rows) {
    const long long ntiles = ((rows + 15) / 16 + FFB_WARPS - 1) / FFB_WARPS;
    return (int)std::min<long long>(ntiles, (long long)num_sms());
}

template <int KS, int NTO, int KS2>
static int launch_ff_bwd_rr(const FFBwdRRArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_bwd_rr<KS, NTO, KS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin() - 2048);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_bwd_rr)");
        attr_set = true;
    }
    if (launch_pdl(k_ff_bwd_rr<KS, NTO, KS2>, dim3(grid), dim3(FFB_WARPS * 32), (size_t)a.smem_bytes, st, a) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "k_ff_bwd_rr");
    RAT_CHECK_LAUNCH("k_ff_bwd_rr");
    return RAT_OK;
}

}  // namespace rat

using namespace rat;

// returns 1 when the shape is outside this kernel's envelope (the caller falls back to the tile kernel)
int ff_fwd_rr_dispatch(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b, const float* W1,
                       const float* b1, const float* W2, const float* b2, long long rows, int D, int M, cudaStream_t st) {
    if (D < 2 || (D & 1) || D > 48 || M < 1 || M > 96 || res == nullptr) return 1;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(res)) & 7) != 0) return 1;
    FFRRArgs a{x, res, out, ln_w, ln_b, W1, b1, W2, b2, rows, D, M};
    const int NTO = (D + 7) / 8;
    switch (NTO) {
        case 2: return launch_ff_fwd_rr<1, 2>(a, st);
        case 3: return launch_ff_fwd_rr<2, 3>(a, st);
        case 4: return launch_ff_fwd_rr<2, 4>(a, st);
        case 5: return launch_ff_fwd_rr<3, 5>(a, st);
        case 6: return launch_ff_fwd_rr<3, 6>(a, st);
        default: return 1;
    }
}

int ff_reduce_records(const float* partials, int nparts, int psize, float* dW1, float* db1, float* dW2, float* db2, int D, int M,
                      int Kp, int Mp, cudaStream_t st);

size_t ff_bwd_rr_workspace_bytes(long long rows, int D, int M) {
    FFBwdRRArgs a{};
    if (!ff_bwd_rr_plan(D, M, &a)) return 0;
    return (size_t)ff_bwd_rr_grid(rows) * a.psize * sizeof(float);
}

// returns 1 when the shape is outside this kernel's envelope (the caller falls back to the tile kernel)
int ff_bwd_rr_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* W1, const float* b1,
                       const float* W2, float* dW1, float* db1, float* dW2, float* db2, long long rows, int D, int M,
                       const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, cudaStream_t st) {
    FFBwdRRArgs a{};
    if (!ff_bwd_rr_plan(D, M, &a)) return 1;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(base) |
          reinterpret_cast<uintptr_t>(dx)) & 7) != 0) return 1;
    const int grid = ff_bwd_rr_grid(rows);
    if (!workspace || workspace_bytes < (size_t)grid * a.psize * sizeof(float)) return 1;
    reduce_ws_acquire(st, workspace);       // a deferred reduction may still be reading the records of an earlier call
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.partials = workspace;
    a.dout_amax = dout_amax; a.dx_amax = dx_amax; a.rows = rows;
    const int NTO = (D + 7) / 8, KS2 = pad16(M) / 16;
    int rc = 1;
#define RAT_FFB(KS_, NTO_) (KS2 == 1 ? launch_ff_bwd_rr<KS_, NTO_, 1>(a, grid, st) : KS2 == 2 ? launch_ff_bwd_rr<KS_, NTO_, 2>(a, grid, st) : \
                            KS2 == 3 ? launch_ff_bwd_rr<KS_, NTO_, 3>(a, grid, st) : KS2 == 5 ? launch_ff_bwd_rr<KS_, NTO_, 5>(a, grid, st) : \
                            KS2 == 6 ? launch_ff_bwd_rr<KS_, NTO_, 6>(a, grid, st) : 1)
    switch (NTO) {
        case 2: rc = RAT_FFB(1, 2); break;
        case 3: rc = RAT_FFB(2, 3); break;
        case 5: rc = RAT_FFB(3, 5); break;
        default: return 1;
    }
#undef RAT_FFB
    if (rc != RAT_OK) return rc;
    return ff_reduce_records(workspace, grid, a.psize, dW1, db1, dW2, db2, D, M, pad16(D), pad16(M), st);
}
