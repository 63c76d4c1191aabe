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
    int Kp, Np;                 // pad16(D + 1): K of the q|k|v GEMM (the + 1 is the column of ones), pad16(D): N of the out-projection
    int SPT;                    // sequences per tile
    int smem_bytes, group_bytes;
    int off_wqkv, off_wo, off_f32, off_grp;       // shared images ; first group's tiles
    int off_q, off_k, off_v, off_o, off_p;        // inside a group's block (the x tile is at 0)
};

constexpr int T2_DHP = 16;                        // padded head width
constexpr int T2_HT = 128 * T2_DHP * 2;           // bytes of one head tile (q, k, v, o or compact P): 4 KB

// Wqkv image: rows n = [head][q|k|v][DHP], K-major, Kp columns:
//   [n][c] = mul * W[n][c] * gamma[c]  (c < D),   [n][D] = mul * sum_c W[n][c] beta[c]   (the LayerNorm affine, folded;
//   the x tile carries normalised rows and a column of ones at column D)
__device__ __forceinline__ void t2_stage_wqkv(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                              const float* __restrict__ Wv, const float* __restrict__ ln_w,
                                              const float* __restrict__ ln_b, float qscale, int D, int dh, int H, int Kp,
                                              unsigned char* __restrict__ img) {
    const int RI = 3 * T2_DHP * H, KC1 = Kp >> 3;
    for (int i = threadIdx.x; i < RI * KC1; i += blockDim.x) {
        const int n = i % RI, kc = i / RI;
        const int h = n / (3 * T2_DHP), rem = n - h * 3 * T2_DHP;
        const int w = rem / T2_DHP, dd = rem - w * T2_DHP;
        const bool live = dd < dh;
        const float* W = (w == 0 ? Wq : w == 1 ? Wk : Wv) + (size_t)(h * dh + dd) * D;
        const float mul = w == 0 ? qscale : 1.0f;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = kc * 8 + k;
            v[k] = 0.f;
            if (live && c < D) v[k] = mul * __ldg(W + c) * __ldg(ln_w + c);
            else if (live && c == D) {                 // D is even: two interleaved chains, loads issued in batches of 8
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
                for (int c2 = 0; c2 < D; c2 += 2) {
                    acc0 = fmaf(__ldg(W + c2), __ldg(ln_b + c2), acc0);
                    acc1 = fmaf(__ldg(W + c2 + 1), __ldg(ln_b + c2 + 1), acc1);
                }
                v[k] = mul * (acc0 + acc1);
            }
        }
        sts128(img + tc5::kmajor_off(n, kc, RI), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
}

// stage one whole tile (128 rows) with the 128 threads of a group: 4 passes of 32 rows, all global loads first
template <bool VEC4, int SLSH>
__device__ __forceinline__ void t2_stage_tile(const float* __restrict__ x, const SeqGeom& g, long long s0, long long nseq, int D,
                                              unsigned char* __restrict__ Xt, int gt, float* __restrict__ stats) {
#pragma unroll 1
    for (int p0 = 0; p0 < 4; p0 += 2) {               // two rounds of two rows per thread (register budget)
        XRegs<VEC4> xr[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) t2_rows_load<VEC4, SLSH>(x, g, s0, nseq, D, 32 * (p0 + p) + (gt >> 2), gt & 3, xr[p]);
#pragma unroll
        for (int p = 0; p < 2; ++p) t2_rows_finish<VEC4>(xr[p], D, Xt, 32 * (p0 + p) + (gt >> 2), gt & 3, stats);
    }
}

template <int SL, int ST, bool VEC4>
__global__ void __launch_bounds__(T2_THREADS, 1) k_attn_fwd_tc2(AttnTc2Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_a[4], bar_y[4];
    __shared__ uint32_t tmem_base_s;
    constexpr int SLSH = SL == 16 ? 4 : 3;
    constexpr int NV = ST > 0 ? ST : SL;              // keys visited by the softmax loops
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int grp = warp >> 2, q = warp & 3, gt = threadIdx.x & 127;
    const bool issuer = q == grp;                     // the four groups issue from four different SM sub-partitions
    const int D = a.D, Kp = a.Kp, Np = a.Np, H = a.H, S = a.g.S;
    const int RI = 3 * T2_DHP * H;                    // rows of the q|k|v weight image
    unsigned char* Wqkv_i = smem + a.off_wqkv;        // [RI x Kp]
    unsigned char* Wo_i = smem + a.off_wo;            // [Np x H*DHP]
    float* bos = reinterpret_cast<float*>(smem + a.off_f32);     // [Np]
    unsigned char* gb = smem + a.off_grp + (size_t)grp * a.group_bytes;
    unsigned char* Xt = gb;                           // [128 x Kp]   normalised x | 1
    unsigned char* Qh = gb + a.off_q;                 // [2 halves][2 chunks][64 rows][16 B]   q of the current head (scaled)
    unsigned char* Kh = gb + a.off_k;
    unsigned char* Vh = gb + a.off_v;
    unsigned char* Oh = gb + a.off_o;                 // [2 chunks][128 rows][16 B]            o of the current head
    unsigned char* Pc = gb + a.off_p;                 // [2 halves][2 chunks][64 rows][16 B]   compact P: row x 16 keys of its row group

    // ---- resident weight images, zero-initialised activation tiles
    t2_stage_wqkv(a.Wq, a.Wk, a.Wv, a.ln_w, a.ln_b, a.qscale, D, a.dh, H, Kp, Wqkv_i);
    {
        const int KCo = (H * T2_DHP) >> 3;
        for (int i = threadIdx.x; i < Np * KCo; i += blockDim.x) {
            const int n = i % Np, kc = i / Np;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = kc * 8 + k, h = c / T2_DHP, dd = c - h * T2_DHP;
                v[k] = (n < D && dd < a.dh) ? __ldg(a.Wo + (size_t)n * a.I + h * a.dh + dd) : 0.f;
            }
            sts128(Wo_i + tc5::kmajor_off(n, kc, Np), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
        for (int i = threadIdx.x; i < Np; i += blockDim.x) bos[i] = i < D ? a.bo[i] : 0.f;
        // x tiles (pad columns stay zero), compact P tiles (the other sub-slot's chunk stays zero when SL == 8)
        for (int i = threadIdx.x; i < (a.smem_bytes - a.off_grp) / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem + a.off_grp)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x < 4) { tc5::mbar_init(&bar_a[threadIdx.x], 1); tc5::mbar_init(&bar_y[threadIdx.x], 1); }
    if (threadIdx.x == 0) tc5::fence_mbar_init();
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t t_w = tmem_base_s + (uint32_t)(grp * 128);            // q|k|v accumulator, then scores, then o (64 columns)
    const uint32_t t_Y = t_w + 64;                                       // out-projection accumulator (Np columns)
    const uint32_t idesc_q = tc5::instr_desc(TC_FMT, 128, 3 * T2_DHP);
    const uint32_t idesc_s = tc5::instr_desc(TC_FMT, 64, 64);
    const uint32_t idesc_pv = tc5::instr_desc(TC_FMT, 64, 64, 0, 1);
    const uint32_t idesc_y = tc5::instr_desc(TC_FMT, 128, Np);
    const uint32_t Xs = tc5::smem_u32(Xt), Qs = tc5::smem_u32(Qh), Ks = tc5::smem_u32(Kh), Vs = tc5::smem_u32(Vh);
    const uint32_t Os = tc5::smem_u32(Oh), Ps = tc5::smem_u32(Pc), Wqs = tc5::smem_u32(Wqkv_i), Wos = tc5::smem_u32(Wo_i);
    uint64_t* bar = &bar_a[grp];
    uint64_t* bary = &bar_y[grp];
    // rows of this thread: M=128 accumulators (q|k|v, y): row_e = TMEM lane ; M=64 accumulators (scores, o): half hf, row 16q+li
    const int row_e = q * 32 + lane;
    const int hf = lane >> 4, li = lane & 15;
    const int row_s = 64 * hf + 16 * q + li;
    const int sb = (li >> 3) & 1;                                       // sub-slot inside the row group (SL == 8)
    unsigned char* const e_row = (unsigned char*)0 + ((q >> 1) * 2048 + ((q & 1) * 32 + lane) * 16);   // offset of row_e inside a half-tiled head tile
    unsigned char* const p_row = Pc + hf * 2048 + ((SL == 8 ? sb : 0) * 64 + 16 * q + li) * 16;
    unsigned char* const o_row = Oh + row_s * 16;
    const size_t e_off = (size_t)(e_row - (unsigned char*)0);
    uint32_t ph = 0, phy = 0;
    const int nck = (D + 7) >> 3;

#define T2_QKV_MMA(h_)                                                                                             \
    do {                                                                                                           \
        for (int k_ = 0; k_ < Kp / 16; ++k_)                                                                       \
            tc5::mma_f16_w(t_w, tc5::kdesc(Xs, 128, k_),                                                           \
                           tc5::smem_desc(Wqs + (uint32_t)((k_ * 2 * RI + 3 * T2_DHP * (h_)) * 16), (uint32_t)RI * 16, 128u), idesc_q, k_ > 0); \
    } while (0)

    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    const long long tstride = (long long)gridDim.x * 4;
    long long tile = (long long)blockIdx.x * 4 + grp;
    if (tile < ntiles) {
        t2_stage_tile<VEC4, SLSH>(a.x, a.g, tile * a.SPT, a.nseq, D, Xt, gt, nullptr);
        tc5::fence_proxy_async();
        group_sync(grp);
        if (issuer) { tc5::fence_after_sync(); T2_QKV_MMA(0); tc5::mma_commit_w(bar); }
    }
    for (; tile < ntiles; tile += tstride) {
        const long long s0 = tile * a.SPT;
        const long long seq_s = s0 + (row_s >> SLSH);
        const bool valid_s = (row_s & (SL - 1)) < S && seq_s < a.nseq;
        for (int h = 0; h < H; ++h) {
            // ---- q|k|v of head h: accumulator -> fp16 head tiles (thread = row row_e)
            tc5::mbar_wait(bar, ph); ph ^= 1;
            tc5::fence_after_sync();
            {
                float v[3][16];
#pragma unroll
                for (int w = 0; w < 3; ++w) tc5::tmem_ld16(t_w + lane_base + w * 16, v[w]);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    unsigned char* dst = (w == 0 ? Qh : w == 1 ? Kh : Vh) + e_off;
                    sts128(dst, pack_h2(v[w][0], v[w][1]), pack_h2(v[w][2], v[w][3]), pack_h2(v[w][4], v[w][5]), pack_h2(v[w][6], v[w][7]));
                    sts128(dst + 1024, pack_h2(v[w][8], v[w][9]), pack_h2(v[w][10], v[w][11]), pack_h2(v[w][12], v[w][13]), pack_h2(v[w][14], v[w][15]));
                }
            }
            tc5::fence_proxy_async();
            tc5::fence_before_sync();
            group_sync(grp);
            if (issuer) {                             // scores: one M=64, N=64, K=16 product per half
                tc5::fence_after_sync();
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                    tc5::mma_f16_w(t_w + ((uint32_t)(16 * h2) << 16), tc5::smem_desc(Qs + h2 * 2048, 1024u, 128u),
                                   tc5::smem_desc(Ks + h2 * 2048, 1024u, 128u), idesc_s, 0u);
                tc5::mma_commit_w(bar);
            }
            // ---- softmax of this thread's row
            tc5::mbar_wait(bar, ph); ph ^= 1;
            tc5::fence_after_sync();
            {
                float v[16];
                tc5::tmem_ld16(t_w + lane_base + 16 * q, v);
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
                if (SL == 16) sts128(p_row + 1024, pk[SL / 2 - 4], pk[SL / 2 - 3], pk[SL / 2 - 2], pk[SL / 2 - 1]);
            }
            tc5::fence_proxy_async();
            tc5::fence_before_sync();
            group_sync(grp);
            if (issuer) {                             // o = P v: compact P (K = the 16 keys of the row group) against v read MN-major,
                tc5::fence_after_sync();              // N = 64 = (dim chunk, row group, dim % 8); a row's own block is at columns 8g and 32 + 8g
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                    tc5::mma_f16_w(t_w + ((uint32_t)(16 * h2) << 16), tc5::smem_desc(Ps + h2 * 2048, 1024u, 128u),
                                   tc5::smem_desc(Vs + h2 * 2048, 128u, 256u), idesc_pv, 0u);
                tc5::mma_commit_w(bar);
            }
            // ---- o of this head -> fp16 o tile
            tc5::mbar_wait(bar, ph); ph ^= 1;
            tc5::fence_after_sync();
            {
                float o0[8], o1[8];
                tc5::tmem_ld8(t_w + lane_base + 8 * q, o0);
                tc5::tmem_ld8(t_w + lane_base + 32 + 8 * q, o1);
                tc5::tmem_ld_wait();
                sts128(o_row, pack_h2(o0[0], o0[1]), pack_h2(o0[2], o0[3]), pack_h2(o0[4], o0[5]), pack_h2(o0[6], o0[7]));
                sts128(o_row + 2048, pack_h2(o1[0], o1[1]), pack_h2(o1[2], o1[3]), pack_h2(o1[4], o1[5]), pack_h2(o1[6], o1[7]));
            }
            tc5::fence_proxy_async();
            tc5::fence_before_sync();
            group_sync(grp);
            if (issuer) {                             // y (+)= o_h Wo_h^T ; then the q|k|v product of the next head
                tc5::fence_after_sync();
                tc5::mma_f16_w(t_Y, tc5::kdesc(Os, 128, 0), tc5::kdesc(Wos, Np, h), idesc_y, h > 0);
                if (h + 1 < H) { T2_QKV_MMA(h + 1); tc5::mma_commit_w(bar); }
                else tc5::mma_commit_w(bary);
            }
        }
        // ---- next tile: stage its rows and start its first q|k|v product, then finish this tile under it
        const long long gr = [&]() -> long long {
            const int slot = row_e >> SLSH, pos = row_e & (SL - 1);
            const long long seq = s0 + slot;
            return (pos < S && seq < a.nseq) ? a.g.grow(seq, pos) : -1;
        }();
        if (tile + tstride < ntiles) {
            t2_stage_tile<VEC4, SLSH>(a.x, a.g, (tile + tstride) * a.SPT, a.nseq, D, Xt, gt, nullptr);
            tc5::fence_proxy_async();
            group_sync(grp);
            if (issuer) { tc5::fence_after_sync(); T2_QKV_MMA(0); tc5::mma_commit_w(bar); }
        }
        // ---- epilogue: out = res + alpha * (y + bo)
        tc5::mbar_wait(bary, phy); phy ^= 1;
        tc5::fence_after_sync();
        for (int c0 = 0; c0 < nck; c0 += 3) {          // three 8-column chunks per round: their residual loads are in flight together
            float v[3][8], rv[3][8];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (c0 + u < nck) {
                    tc5::tmem_ld8(t_Y + lane_base + (c0 + u) * 8, v[u]);
                    if (gr >= 0 && a.res) load8<VEC4>(a.res + gr * D, (c0 + u) * 8, D, rv[u]);
                }
            }
            tc5::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (c0 + u < nck && gr >= 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        v[u][k] = a.alpha * (v[u][k] + bos[(c0 + u) * 8 + k]);
                        if (a.res) v[u][k] += rv[u][k];
                    }
                    store8<VEC4>(a.out + gr * D, (c0 + u) * 8, D, v[u]);
                }
            }
        }
        tc5::fence_before_sync();
        group_sync(grp);                              // y is consumed before the next tile's first out-projection overwrites it
    }
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace rat

using namespace rat;

template <int SL, int ST, bool VEC4>
static int launch_attn_fwd_tc2(const AttnTc2Args& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_fwd_tc2<SL, ST, VEC4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 2048);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_fwd_tc2)");
        attr_set = true;
    }
    const long long ntiles = (a.nseq + a.SPT - 1) / a.SPT;
    const int grid = (int)std::min<long long>((ntiles + 3) / 4, (long long)num_sms());
    k_attn_fwd_tc2<SL, ST, VEC4><<<grid, T2_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_fwd_tc2");
    return RAT_OK;
}

template <bool VEC4>
static int launch_attn_fwd_tc2_s(const AttnTc2Args& a, cudaStream_t st) {
    const int S = a.g.S;
    if (S > 8) return S == 14 ? launch_attn_fwd_tc2<16, 14, VEC4>(a, st) : launch_attn_fwd_tc2<16, 0, VEC4>(a, st);
    return S == 6 ? launch_attn_fwd_tc2<8, 6, VEC4>(a, st) : launch_attn_fwd_tc2<8, 0, VEC4>(a, st);
}

// Planning of the second-generation attention forward: shared-memory map.  Returns false if the shape is not covered.
bool attn_tc2_plan(int S, int D, int heads, int dh, AttnTc2Args* a) {
    if (S > 16 || S < 1 || dh > 16 || dh < 2 || (dh & 1) || D < 2 || (D & 1) || D > 64) return false;
    if ((D % 4) != 0 && D > 32) return false;
    a->D = D; a->H = heads; a->dh = dh; a->I = heads * dh;
    a->Kp = pad16(D + 1); a->Np = pad16(D);          // + 1: the column of ones that carries the LayerNorm beta
    const int SL = S > 8 ? 16 : 8;
    a->SPT = 128 / SL;
    size_t off = 0;
    a->off_wqkv = (int)off; off += (size_t)3 * T2_DHP * heads * a->Kp * 2;
    a->off_wo = (int)off; off += (size_t)a->Np * heads * T2_DHP * 2;
    a->off_f32 = (int)off; off += (size_t)a->Np * 4;
    off = (off + 127) & ~(size_t)127;
    a->off_grp = (int)off;
    size_t g = (size_t)128 * a->Kp * 2;
    a->off_q = (int)g; g += T2_HT;
    a->off_k = (int)g; g += T2_HT;
    a->off_v = (int)g; g += T2_HT;
    a->off_o = (int)g; g += T2_HT;
    a->off_p = (int)g; g += T2_HT;
    a->group_bytes = (int)g;
    off += 4 * g;
    a->smem_bytes = (int)off;
    return off <= (size_t)max_smem_optin() - 2048;
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
    return (D % 4) == 0 ? launch_attn_fwd_tc2_s<true>(a, st) : launch_attn_fwd_tc2_s<false>(a, st);
}
