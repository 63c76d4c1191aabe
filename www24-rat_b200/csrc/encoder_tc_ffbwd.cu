// K5 (Blackwell path): FeedForward backward -- tcgen05 token-major GEMMs + register-resident mma.sync weight gradients.
#include "encoder_tc.cuh"

namespace rat {

// Backward of  y = x + W2 gelu(W1 x + b1) + b2  (no pre-norm):
//   pre = x W1^T + b1 (recomputed) ; dh = dy W2 ; dpre = dh * gelu'(pre) ; dx = base + dpre W1
//   gW1 = dpre^T x ; gb1 = colsum(dpre) (ones column of the x tile) ; gW2^T = h^T dy ; gb2 = colsum(dy)
// The three token-major products run on tcgen05 (accumulators in TMEM); the weight-gradient products reduce over
// tokens with mma.sync and stay in REGISTERS across all tiles of the CTA (one record per team is written at the end
// and k_reduce_ff_tc sums the records in fixed order: bitwise deterministic).
struct FFBwdTcArgs {
    const float* x; const float* dout; const float* base; float* dx;
    const float* W1; const float* b1; const float* W2;
    float* partials;         // [2 * gridDim.x][psize]
    const float* dout_amax;  // device: max|dout| (nullptr: no gradient scaling)
    float* dx_amax;          // device: receives max|dx| (nullptr: not wanted)
    long long rows;
    int D, M, Kp, Mp;
    int psize;               // 2 * Mp * Kp + Kp
    int smem_bytes;
};

template <int KCH, bool VEC4, int JW>
__global__ void __launch_bounds__(TC_THREADS, 1) k_ff_bwd_tc(FFBwdTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = a.D, M = a.M, Kp = a.Kp, Mp = a.Mp;
    const int KC1 = Kp >> 3, KC2 = Mp >> 3;
    unsigned char* W1i = smem_raw;                                   // [Mp x Kp]  (n = m, k = d) = W1[m][d]
    unsigned char* W2ti = W1i + (size_t)Mp * Kp * 2;                 // [Mp x Kp]  (n = m, k = d) = W2[d][m]
    unsigned char* W1ti = W2ti + (size_t)Mp * Kp * 2;                // [Kp x Mp]  (n = d, k = m) = W1[m][d]
    float* b1s = reinterpret_cast<float*>(W1ti + (size_t)Kp * Mp * 2);   // [Mp]
    unsigned char* team_base = reinterpret_cast<unsigned char*>(b1s + Mp);
    const size_t team_bytes = (size_t)TILE_M * (2 * Kp + 2 * Mp) * 2;
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int team = threadIdx.x / TEAM_THREADS, tid2 = threadIdx.x % TEAM_THREADS;
    const int warp2 = tid2 >> 5, lane = tid2 & 31;
    unsigned char* Xt = team_base + team * team_bytes;               // [128 x Kp]  x (+ ones column)
    unsigned char* DYt = Xt + (size_t)TILE_M * Kp * 2;               // [128 x Kp]  dy
    unsigned char* Ht = DYt + (size_t)TILE_M * Kp * 2;               // [128 x Mp]  h = gelu(pre)
    unsigned char* DPt = Ht + (size_t)TILE_M * Mp * 2;               // [128 x Mp]  dpre

    stage_weight_image(a.W1, M, D, Mp, Kp, W1i);
    {   // transposed images
        const int total = Mp * KC1;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int m = i % Mp, kc = i / Mp;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int d = kc * 8 + k; v[k] = (m < M && d < D) ? __ldg(a.W2 + (size_t)d * M + m) : 0.f; }
            sts128(W2ti + tc5::kmajor_off(m, kc, Mp), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                   pack_h2(v[6], v[7]));
        }
        const int total2 = Kp * KC2;
        for (int i = threadIdx.x; i < total2; i += blockDim.x) {
            const int d = i % Kp, kc = i / Kp;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int m = kc * 8 + k; v[k] = (m < M && d < D) ? __ldg(a.W1 + (size_t)m * D + d) : 0.f; }
            sts128(W1ti + tc5::kmajor_off(d, kc, Kp), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                   pack_h2(v[6], v[7]));
        }
    }
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) b1s[i] = i < M ? a.b1[i] : 0.f;
    if (threadIdx.x == 0) { tc5::mbar_init(&mbar[0], 1); tc5::mbar_init(&mbar[1], 1); tc5::fence_mbar_init(); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&tmem_base_s, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_P = tmem_base_s + team * 256;                // pre [0, Mp)  -> later dxa [0, Kp)
    const uint32_t tmem_H = tmem_P + Mp;                             // dh  [Mp, 2 Mp)
    const uint32_t idesc_m = tc5::instr_desc(TC_FMT, TILE_M, Mp);
    const uint32_t idesc_d = tc5::instr_desc(TC_FMT, TILE_M, Kp);
    const uint32_t lane_base = (uint32_t)((warp2 & 3) * 32) << 16;
    const int chalf = warp2 >> 2;
    const int row_e = (warp2 & 3) * 32 + lane;
    uint32_t phase = 0;
    uint64_t* bar = &mbar[team];

    // weight-gradient jobs of this warp: job id = warp2 + 8 j
    //   [0, MT*NP): gW1 (A = dpre, B = x) ; [MT*NP, 2 MT*NP): gW2^T (A = h, B = dy) ; then NP jobs gb2 (A = ones, B = dy)
    const int MT = Mp >> 4, NP = KC1 >> 1, njobs = 2 * MT * NP + NP;
    float acc[JW][2][4];
#pragma unroll
    for (int j = 0; j < JW; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) acc[j][q][0] = acc[j][q][1] = acc[j][q][2] = acc[j][q][3] = 0.f;

    const float gs = tc_grad_scale(a.dout_amax), inv_gs = 1.0f / gs;
    float dx_max = 0.f;
    const long long ntiles = (a.rows + TILE_M - 1) / TILE_M;
    for (long long tile = (long long)blockIdx.x * 2 + team; tile < ntiles; tile += (long long)gridDim.x * 2) {
        const long long r0 = tile * TILE_M;
        const int R = (int)min((long long)TILE_M, a.rows - r0);
        {
            const int row = tid2 >> 1, h = tid2 & 1;
            stage_row_h<KCH, VEC4>(a.x + (r0 + row) * D, row < R, D, KC1, row, h, nullptr, nullptr, Xt, D);
            stage_row_h<KCH, VEC4>(a.dout + (r0 + row) * D, row < R, D, KC1, row, h, nullptr, nullptr, DYt, -1, nullptr, gs);
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        if (tid2 == 0) {
            tc5::fence_after_sync();
            const uint32_t x0 = tc5::smem_u32(Xt), y0 = tc5::smem_u32(DYt), w1 = tc5::smem_u32(W1i), w2 = tc5::smem_u32(W2ti);
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_P, tc5::kdesc(x0, TILE_M, k), tc5::kdesc(w1, Mp, k),
                             idesc_m, k > 0);
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_H, tc5::kdesc(y0, TILE_M, k), tc5::kdesc(w2, Mp, k),
                             idesc_m, k > 0);
            tc5::mma_commit(bar);
        }
        tc5::mbar_wait(bar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 1: h = gelu(pre + b1), dpre = dh * gelu'(pre + b1)  -> bf16 tiles
        {
            const int ng = Mp >> 3, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
            for (int g = g0; g < g1; ++g) {
                float p[8], dh[8], hv[8], dp[8];
                tc5::tmem_ld8(tmem_P + lane_base + g * 8, p);
                tc5::tmem_ld8(tmem_H + lane_base + g * 8, dh);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float gd;
                    gelu_fast(p[k] + b1s[g * 8 + k], hv[k], gd);
                    dp[k] = dh[k] * gd;
                }
                sts128(Ht + tc5::toff(row_e, g), pack_h2(hv[0], hv[1]), pack_h2(hv[2], hv[3]),
                       pack_h2(hv[4], hv[5]), pack_h2(hv[6], hv[7]));
                sts128(DPt + tc5::toff(row_e, g), pack_h2(dp[0], dp[1]), pack_h2(dp[2], dp[3]),
                       pack_h2(dp[4], dp[5]), pack_h2(dp[6], dp[7]));
            }
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        // ---- dxa[128 x Kp] = dpre . W1   (tcgen05) overlapped with the weight-gradient jobs (mma.sync)
        if (tid2 == 0) {
            tc5::fence_after_sync();
            const uint32_t p0 = tc5::smem_u32(DPt), w = tc5::smem_u32(W1ti);
            for (int k = 0; k < Mp / 16; ++k)
                tc5::mma_f16(tmem_P, tc5::kdesc(p0, TILE_M, k), tc5::kdesc(w, Kp, k),
                             idesc_d, k > 0);
            tc5::mma_commit(bar);
        }
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            const int job = warp2 + 8 * j;
            if (job < njobs) {
                if (job < 2 * MT * NP) {
                    const int which = job / (MT * NP), rem = job - which * (MT * NP);
                    const int mi = rem / NP, np = rem - mi * NP;
                    wgrad_job(which ? Ht : DPt, KC2, 2 * mi, false, which ? DYt : Xt, KC1, 2 * np, lane, acc[j]);
                } else {
                    wgrad_job(Ht, KC2, 0, true, DYt, KC1, 2 * (job - 2 * MT * NP), lane, acc[j]);
                }
            }
        }
        tc5::mbar_wait(bar, phase);
        phase ^= 1;
        tc5::fence_after_sync();
        // ---- epilogue 2: dx = base + dxa
        {
            const int ng = Kp >> 3, g0 = chalf * (ng >> 1), g1 = chalf ? ng : (ng >> 1);
            for (int g = g0; g < g1; ++g) {
                if (g * 8 >= D) break;
                float v[8], bv[8];
                tc5::tmem_ld8(tmem_P + lane_base + g * 8, v);
                if (row_e < R && a.base) load8<VEC4>(a.base + (r0 + row_e) * D, g * 8, D, bv);
                tc5::tmem_ld_wait();
                if (row_e < R) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] *= inv_gs;
                    if (a.base) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] += bv[k];
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (g * 8 + k < D) dx_max = fmaxf(dx_max, fabsf(v[k]));
                    store8<VEC4>(a.dx + (r0 + row_e) * D, g * 8, D, v);
                }
            }
        }
        tc5::fence_before_sync();
        team_sync(team);                    // every warp is done reading this tile's operand tiles
    }
    // ---- per-team gradient record: gW1 [Mp][Kp] | gW2T [Mp][Kp] | gb2 [Kp]
    {
        float* rec = a.partials + (size_t)(blockIdx.x * 2 + team) * a.psize;
#pragma unroll
        for (int j = 0; j < JW; ++j) {
            const int job = warp2 + 8 * j;
            if (job < njobs) {
                if (job < 2 * MT * NP) {
                    const int which = job / (MT * NP), rem = job - which * (MT * NP);
                    const int mi = rem / NP, np = rem - mi * NP;
                    wgrad_store(rec + (size_t)which * Mp * Kp, Kp, 16 * mi, 16 * np, lane, acc[j], false, inv_gs);
                } else {
                    wgrad_store(rec + (size_t)2 * Mp * Kp, Kp, 0, 16 * (job - 2 * MT * NP), lane, acc[j], true, inv_gs);
                }
            }
        }
    }
    publish_amax_block(a.dx_amax, dx_max);
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tmem_base_s, 512);
}

struct FFReduceTcArgs {
    const float* partials; int nparts, psize;
    float* dW1; float* db1; float* dW2; float* db2;
    int D, M, Kp, Mp;
};
__global__ void k_reduce_ff_tc(FFReduceTcArgs a) {
    pdl_launch_dependents();            // the next backward kernel may run its prologue under this reduction (common.cuh)
    const int D = a.D, M = a.M, Kp = a.Kp, Mp = a.Mp;
    const int total = 2 * M * D + M + D;
    for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
        const int i = min(base + (int)threadIdx.x, total - 1);
        const bool active = base + (int)threadIdx.x < total;
        size_t src;
        float* dst;
        if (i < M * D) { const int m = i / D, d = i - m * D; src = (size_t)m * Kp + d; dst = a.dW1 ? a.dW1 + i : nullptr; }
        else if (i < 2 * M * D) { const int rem = i - M * D; const int d = rem / M, m = rem - d * M;
                                  src = (size_t)Mp * Kp + (size_t)m * Kp + d; dst = a.dW2 ? a.dW2 + rem : nullptr; }
        else if (i < 2 * M * D + M) { const int m = i - 2 * M * D; src = (size_t)m * Kp + D; dst = a.db1 ? a.db1 + m : nullptr; }
        else { const int d = i - 2 * M * D - M; src = (size_t)2 * Mp * Kp + d; dst = a.db2 ? a.db2 + d : nullptr; }
        const float s = record_sum_sliced(a.partials, a.psize, a.nparts, src, active && dst != nullptr);
        if (active && dst != nullptr && threadIdx.y == 0) *dst = s;
    }
}


}  // namespace rat

using namespace rat;

// ---- FF backward (tcgen05) host side -------------------------------------------------------------------------
static bool ff_bwd_tc_plan(int D, int M, FFBwdTcArgs* a) {
    if (!ff_tc_supported(D, M)) return false;
    const int Kp = pad16(D), Mp = pad16(M);
    if (Kp == D) return false;                       // needs a pad column for the ones trick (gb1)
    if (2 * Mp > 256) return false;
    const int njobs = 2 * (Mp / 16) * (Kp / 16) + Kp / 16;
    if ((njobs + 7) / 8 > 5) return false;
    a->D = D; a->M = M; a->Kp = Kp; a->Mp = Mp;
    a->psize = 2 * Mp * Kp + Kp;
    const size_t fixed = (size_t)3 * Mp * Kp * 2 + (size_t)Mp * 4;
    const size_t team = (size_t)TILE_M * (2 * Kp + 2 * Mp) * 2;
    a->smem_bytes = (int)(fixed + 2 * team);
    return a->smem_bytes <= max_smem_optin() - 1024;
}
static int ff_bwd_tc_grid(long long rows) {
    const long long ntiles = (rows + TILE_M - 1) / TILE_M;
    return (int)std::min<long long>((ntiles + 1) / 2, (long long)num_sms());
}
size_t ff_bwd_tc_workspace_bytes(long long rows, int D, int M) {
    FFBwdTcArgs a{};
    if (!ff_bwd_tc_plan(D, M, &a)) return 0;
    return (size_t)2 * ff_bwd_tc_grid(rows) * a.psize * sizeof(float);
}

template <int KCH, bool VEC4, int JW>
static int launch_ff_bwd_tc(const FFBwdTcArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ff_bwd_tc<KCH, VEC4, JW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ff_bwd_tc)");
        attr_set = true;
    }
    k_ff_bwd_tc<KCH, VEC4, JW><<<grid, TC_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_ff_bwd_tc");
    return RAT_OK;
}

int ff_bwd_tc_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                       const float* W1, const float* b1, const float* W2, float* dW1, float* db1, float* dW2, float* db2,
                       long long rows, int D, int M, const float* dout_amax, float* dx_amax, float* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
    FFBwdTcArgs a{};
    if (ln_w != nullptr || !ff_bwd_tc_plan(D, M, &a)) return 1;
    const int grid = ff_bwd_tc_grid(rows);
    if (!workspace || workspace_bytes < (size_t)2 * grid * a.psize * sizeof(float)) return 1;
    reduce_ws_acquire(st, workspace);
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.partials = workspace; a.rows = rows;
    a.dout_amax = dout_amax; a.dx_amax = dx_amax;
    const int kch = a.Kp / 16;
    const bool v4 = (D % 4) == 0;
    const int njobs = 2 * (a.Mp / 16) * (a.Kp / 16) + a.Kp / 16;
    const int jw = (njobs + 7) / 8;
    int rc = 1;
#define RAT_FFB2(K_, V_) (jw <= 1 ? launch_ff_bwd_tc<K_, V_, 1>(a, grid, st) : jw == 2 ? launch_ff_bwd_tc<K_, V_, 2>(a, grid, st) : \
                          jw == 3 ? launch_ff_bwd_tc<K_, V_, 3>(a, grid, st) : jw == 4 ? launch_ff_bwd_tc<K_, V_, 4>(a, grid, st) : \
                          launch_ff_bwd_tc<K_, V_, 5>(a, grid, st))
#define RAT_FFB(K_) (v4 ? RAT_FFB2(K_, true) : RAT_FFB2(K_, false))
    switch (kch) {
        case 1: rc = RAT_FFB(1); break;
        case 2: rc = RAT_FFB(2); break;
        case 3: rc = RAT_FFB(3); break;
        default: return 1;
    }
#undef RAT_FFB
#undef RAT_FFB2
    if (rc != RAT_OK) return rc;
    FFReduceTcArgs r{workspace, 2 * grid, a.psize, dW1, db1, dW2, db2, D, M, a.Kp, a.Mp};
    const int total = 2 * M * D + M + D;
    cudaStream_t rs = reduce_fork(st, workspace);
    k_reduce_ff_tc<<<std::max(1, std::min((total + 31) / 32, 1024)), dim3(32, 8), 0, rs>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_ff_tc");
    reduce_forked(rs, st, workspace);
    return RAT_OK;
}


// fixed-order reduction of per-CTA FF gradient records [Mp][Kp] gW1 (column D = gb1) | [Mp][Kp] gW2^T | [Kp] gb2 -- shared with
// the register-resident backward (encoder_rr_ff.cu)
int ff_reduce_records(const float* partials, int nparts, int psize, float* dW1, float* db1, float* dW2, float* db2, int D, int M,
                      int Kp, int Mp, cudaStream_t st) {
    FFReduceTcArgs r{partials, nparts, psize, dW1, db1, dW2, db2, D, M, Kp, Mp};
    const int total = 2 * M * D + M + D;
    cudaStream_t rs = reduce_fork(st, partials);
    k_reduce_ff_tc<<<std::max(1, std::min((total + 31) / 32, 1024)), dim3(32, 8), 0, rs>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_ff_tc");
    reduce_forked(rs, st, partials);
    return RAT_OK;
}
