// K3 (Blackwell path): dense GEMM of the DNN head on the 5th-generation tensor cores.
//
//   C[m][n] = sum_k opA(m,k) * opB(k,n) (+ bias[n])       fp32 in / fp32 out, fp16 operands (bf16 with RAT_TC_FP16=0), fp32 accumulate in TMEM
//
// Replaces the nn.Linear products of MLP_Layer (fuxictr/pytorch/layers/deep.py:126-137) and their autograd reverse:
//   forward   z  = h W^T + b        opA = h  [B x K]  (K contiguous)   opB = W  [N x K]  (K contiguous)
//   dgrad     dh = dz W             opA = dz [B x u]  (K contiguous)   opB = W  [u x Kin] (rows = reduction index)
//   wgrad     dW = dz^T h           opA = dz [B x u]  (rows = reduction index), opB = h [B x Kin] (rows = reduction index)
// A source tile [rows x cols] is always staged the same way -- converted to the 16-bit operand type and written chunk-major
// ([col/8][row][16 B], tc5.cuh) -- and only the UMMA descriptor changes: an operand whose rows are M/N indices is
// "K-major" (LBO = rows*16, SBO = 128); an operand whose rows are the reduction index is "MN-major" (the same
// 128-byte core matrices read as [k%8][mn%8]: LBO = 128, SBO = rows*16, major bit set).  No transposed copies.
// 128 x BN output tile per CTA (BN <= 256, accumulator = BN TMEM columns), K in stages of 64 with two shared-memory
// buffers: the threads convert/stage buffer s+1 while the tensor core consumes buffer s.  Optional split-K over
// gridDim.z writes fp32 partials that k_splitk_reduce (mlp.cu) adds in fixed order.
#include "encoder_tc.cuh"

namespace rat {

constexpr int GT_THREADS = 256;
constexpr int GT_KB = 64;            // reduction elements per stage

struct GemmTcArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K, lda, ldb, ldc;
    const float* a_amax;             // device max|A| (nullptr: no scaling): A is lifted by grad_scale_from_amax while it
                                     // is staged and the fp32 accumulator is unscaled in the epilogue
    int BN;                          // output-tile width (multiple of 16, <= 256)
    int kchunk;                      // reduction range per split (multiple of 64)
    int splits;
};

// Staging of a [ROWS x 8*CHUNKS] fp32 source tile (row stride ld) as fp16, chunk-major; rows/cols outside the matrix -> 0.
// Split in two halves so that the global loads of stage s+1 are in flight while stage s is converted, stored and multiplied:
//   gemm_tile_load  : this thread's <= U items (8 floats each) -> registers
//   gemm_tile_store : registers -> fp16 -> shared memory
// item -> (row, chunk) of a staged tile: a warp covers 8 rows x 4 chunks, lane = (chunk % 4) * 8 + row % 8.  Its global loads
// touch 8 cache lines (each row contributes 4 x 32 B = one 128-byte line) instead of 32, and every quarter-warp stores 8
// consecutive rows of ONE chunk = 128 contiguous bytes of shared memory (conflict-free).  ROWS % 8 == 0.
__device__ __forceinline__ bool gemm_item(int it, int ROWS, int CHUNKS, int& r, int& c) {
    const int lane = it & 31, tile = it >> 5, rbs = ROWS >> 3;
    r = (tile % rbs) * 8 + (lane & 7);
    c = (tile / rbs) * 4 + (lane >> 3);
    return c < CHUNKS;
}
// Per-thread description of its <= U items of one operand tile, computed ONCE before the stage loop (the first version
// recomputed the item -> (row, chunk) mapping with divisions every stage: 37 % of the kernel's instructions).
//   K-major tile  (rows = m/n index, cols = k):   stage s adds s*64 floats to the source pointer;
//   MN-major tile (rows = k index,  cols = m/n):  stage s adds s*64 rows.
template <int U>
struct TileItems {
    const float* src[U];         // source of the item in stage 0 (nullptr: outside the matrix in the fixed dimension)
    uint32_t soff[U];            // byte offset inside the staged tile
    int kpos[U];                 // position along the reduction dimension inside the stage (first of 8 for K-major, row for MN-major)
    int mnrem[U];                // MN-major: valid columns of the 8 (clipped at the matrix edge); K-major: unused
};
template <int U, bool MN>
__device__ __forceinline__ void gemm_items_init(TileItems<U>& t, const float* __restrict__ base, int ld, int mn0, int mn_total,
                                                int k0, int ROWS, int CHUNKS) {
    const int items = ROWS * ((CHUNKS + 3) & ~3);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int it = threadIdx.x + u * GT_THREADS;
        int r = 0, c = 0;
        t.src[u] = nullptr; t.soff[u] = 0; t.kpos[u] = 0; t.mnrem[u] = 0;
        if (it >= items || !gemm_item(it, ROWS, CHUNKS, r, c)) continue;
        t.soff[u] = tc5::kmajor_off(r, c, ROWS);
        if (!MN) {                                   // row r = m/n index, chunk c = 8 reduction positions
            t.kpos[u] = c * 8;
            if (mn0 + r < mn_total) t.src[u] = base + (size_t)(mn0 + r) * ld + k0 + c * 8;
            else t.kpos[u] = -1;                     // zero row: still stored (the tile is rewritten every stage)
        } else {                                     // row r = reduction position, chunk c = 8 m/n indices
            t.kpos[u] = r;
            t.mnrem[u] = min(8, mn_total - (mn0 + c * 8));
            if (t.mnrem[u] > 0) t.src[u] = base + (size_t)(k0 + r) * ld + mn0 + c * 8;
        }
    }
}
// klen = reduction positions left in this stage's range (kend - k0 of the stage)
template <int U, bool MN>
__device__ __forceinline__ void gemm_tile_load(const TileItems<U>& t, size_t adv, int klen, bool vec_ok, float (&v)[U][8]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = 0.f;
        if (t.src[u] == nullptr) continue;
        const float* p = t.src[u] + adv;
        const int n = MN ? (t.kpos[u] < klen ? t.mnrem[u] : 0) : min(8, klen - t.kpos[u]);
        if (n == 8 && vec_ok) {
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(p)), t1 = __ldg(reinterpret_cast<const float4*>(p + 4));
            v[u][0] = t0.x; v[u][1] = t0.y; v[u][2] = t0.z; v[u][3] = t0.w;
            v[u][4] = t1.x; v[u][5] = t1.y; v[u][6] = t1.z; v[u][7] = t1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < n) v[u][k] = __ldg(p + k);
        }
    }
}
template <int U>
__device__ __forceinline__ void gemm_tile_store(const TileItems<U>& t, int nitems_thread, const float (&v)[U][8],
                                                unsigned char* __restrict__ dst, float mul) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (u >= nitems_thread) continue;
        sts128(dst + t.soff[u], pack_h2(v[u][0] * mul, v[u][1] * mul), pack_h2(v[u][2] * mul, v[u][3] * mul),
               pack_h2(v[u][4] * mul, v[u][5] * mul), pack_h2(v[u][6] * mul, v[u][7] * mul));
    }
}

template <bool A_MN, bool B_MN, int UB>
__global__ void __launch_bounds__(GT_THREADS) k_gemm_tc(GemmTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int BN = a.BN;
    const int m0 = blockIdx.y * TILE_M, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * a.kchunk, kend = min(a.K, kbeg + a.kchunk);
    const int nst = (kend - kbeg + GT_KB - 1) / GT_KB;
    const size_t a_bytes = (size_t)TILE_M * GT_KB * 2, b_bytes = (size_t)BN * GT_KB * 2;
    unsigned char* Abuf[2] = {smem_raw, smem_raw + a_bytes + b_bytes};
    unsigned char* Bbuf[2] = {smem_raw + a_bytes, smem_raw + 2 * a_bytes + b_bytes};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc5::mbar_init(&mbar[0], 1); tc5::mbar_init(&mbar[1], 1); tc5::fence_mbar_init(); }
    const uint32_t tcols = BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;    // several CTAs share an SM's 512 columns
    if (warp == 0) tc5::tmem_alloc(&tmem_base_s, tcols);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_D = tmem_base_s;
    const float a_scale = tc_grad_scale(a.a_amax), c_scale = 1.0f / a_scale;
    const uint32_t idesc = tc5::instr_desc(TC_FMT, TILE_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const bool a_vec = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    const bool b_vec = (a.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0);

    // A tile: K-major -> rows = m (128), cols = k (64) ; MN-major -> rows = k (64), cols = m (128): 1024 items = 4 per thread;
    // B tile: BN x 64 -> <= 2048 items = 8 per thread (BN <= 256), 2 when BN <= 64
    constexpr int UA = TILE_M * (GT_KB / 8) / GT_THREADS;
    TileItems<UA> ia;
    TileItems<UB> ib;
    if (!A_MN) gemm_items_init<UA, false>(ia, a.A, a.lda, m0, a.M, kbeg, TILE_M, GT_KB / 8);
    else gemm_items_init<UA, true>(ia, a.A, a.lda, m0, a.M, kbeg, GT_KB, TILE_M / 8);
    if (!B_MN) gemm_items_init<UB, false>(ib, a.B, a.ldb, n0, a.N, kbeg, BN, GT_KB / 8);
    else gemm_items_init<UB, true>(ib, a.B, a.ldb, n0, a.N, kbeg, GT_KB, BN / 8);
    // items of this thread that exist in the tile (their slots are rewritten every stage, zeros included)
    int na = 0, nb = 0;
    {
        const int items_a = (A_MN ? GT_KB * (TILE_M / 8) : TILE_M * (GT_KB / 8));
        const int items_b = B_MN ? GT_KB * ((BN / 8 + 3) & ~3) : BN * (GT_KB / 8);
        for (int u = 0; u < UA; ++u) if ((int)threadIdx.x + u * GT_THREADS < items_a) na = u + 1;
        for (int u = 0; u < UB; ++u) {
            const int it = threadIdx.x + u * GT_THREADS;
            int r, c;
            if (it < items_b && gemm_item(it, B_MN ? GT_KB : BN, B_MN ? BN / 8 : GT_KB / 8, r, c)) nb = u + 1;
        }
    }
    const size_t adv_a = A_MN ? (size_t)GT_KB * a.lda : (size_t)GT_KB, adv_b = B_MN ? (size_t)GT_KB * a.ldb : (size_t)GT_KB;
    const bool a_v = a_vec && (A_MN || (kbeg % 4 == 0)), b_v = b_vec && (B_MN ? (n0 % 4 == 0) : (kbeg % 4 == 0));
    float ra[UA][8], rb[UB][8];
    auto load_stage = [&](int s) {
        const int klen = kend - (kbeg + s * GT_KB);
        if (!A_MN) gemm_tile_load<UA, false>(ia, adv_a * s, klen, a_v, ra); else gemm_tile_load<UA, true>(ia, adv_a * s, klen, a_v, ra);
        if (!B_MN) gemm_tile_load<UB, false>(ib, adv_b * s, klen, b_v, rb); else gemm_tile_load<UB, true>(ib, adv_b * s, klen, b_v, rb);
    };
    if (nst > 0) load_stage(0);
    for (int s = 0; s < nst; ++s) {
        const int buf = s & 1;
        if (s >= 2) tc5::mbar_wait(&mbar[buf], ((s >> 1) - 1) & 1);      // the MMAs that read this buffer are done
        gemm_tile_store<UA>(ia, na, ra, Abuf[buf], a_scale);
        gemm_tile_store<UB>(ib, nb, rb, Bbuf[buf], 1.0f);
        if (s + 1 < nst) load_stage(s + 1);        // in flight under this stage's barrier, MMA issue and the next buffer wait
        tc5::fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc5::fence_after_sync();
            const uint32_t as = tc5::smem_u32(Abuf[buf]), bs = tc5::smem_u32(Bbuf[buf]);
#pragma unroll
            for (int j = 0; j < GT_KB / 16; ++j) {
                const uint64_t da = A_MN ? tc5::smem_desc(as + j * 256, 128, GT_KB * 16) : tc5::kdesc(as, TILE_M, j);
                const uint64_t db = B_MN ? tc5::smem_desc(bs + j * 256, 128, GT_KB * 16) : tc5::kdesc(bs, BN, j);
                tc5::mma_f16(tmem_D, da, db, idesc, (s > 0 || j > 0) ? 1u : 0u);
            }
            tc5::mma_commit(&mbar[buf]);
        }
    }
    if (nst > 0) tc5::mbar_wait(&mbar[(nst - 1) & 1], ((nst - 1) >> 1) & 1);
    tc5::fence_after_sync();
    // ---- epilogue: thread = accumulator row; 16-column groups alternate between the two warps of a lane quadrant
    {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int row = m0 + (warp & 3) * 32 + lane;
        float* Cz = a.C + (a.splits > 1 ? (size_t)blockIdx.z * a.M * a.ldc : 0);
        const bool c_vec = (a.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0) && (n0 % 4 == 0);
        for (int gq = warp >> 2; gq < BN / 16; gq += 2) {
            float v[16];
            if (nst > 0) { tc5::tmem_ld16(tmem_D + lane_base + gq * 16, v); tc5::tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            const int c0 = n0 + gq * 16;
            if (c_scale != 1.0f) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= c_scale;
            }
            if (row < a.M && c0 < a.N) {
                if (a.bias && a.splits == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) if (c0 + i < a.N) v[i] += __ldg(a.bias + c0 + i);
                }
                float* dst = Cz + (size_t)row * a.ldc + c0;
                if (c_vec && c0 + 16 <= a.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) if (c0 + i < a.N) dst[i] = v[i];
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem_base_s, tcols);
}

struct GemmTcPlan { int BN, n_tiles, m_tiles, splits, kchunk; size_t smem; };

static bool gemm_tc_plan(int M, int N, int K, size_t workspace_bytes, GemmTcPlan* p) {
    if (M < 64 || N < 16 || K < 64) return false;
    p->m_tiles = ceil_div(M, TILE_M);
    // narrow output tiles until every SM has TWO CTAs (each is a short latency chain of K/64 load -> convert -> MMA stages;
    // a second resident CTA fills the first one's waits; TMEM is allocated per CTA as the next power of two >= BN); the
    // A tile re-reads of the extra column tiles hit L2
    p->n_tiles = (N + 255) / 256;
    if (K < 1024) p->n_tiles = std::max(p->n_tiles, std::min(ceil_div(N, 32), ceil_div(2 * num_sms(), p->m_tiles)));   // long K: split-K instead
    p->BN = round_up(ceil_div(N, p->n_tiles), 16);
    p->n_tiles = ceil_div(N, p->BN);
    const int tiles = p->n_tiles * p->m_tiles;
    int splits = 1;
    if (tiles * 2 <= num_sms() && K >= 1024) {
        splits = std::min(32, 2 * num_sms() / tiles);
        while (splits > 1 && K / splits < 256) --splits;
        const size_t per = (size_t)M * N * sizeof(float);
        while (splits > 1 && (size_t)splits * per > workspace_bytes) --splits;
    }
    p->kchunk = round_up(ceil_div(K, splits), GT_KB);
    p->splits = ceil_div(K, p->kchunk);
    p->smem = 2 * ((size_t)TILE_M * GT_KB * 2 + (size_t)p->BN * GT_KB * 2);
    return true;
}

}  // namespace rat

using namespace rat;

size_t gemm_tc_workspace_bytes(int M, int N, int K) {
    GemmTcPlan p{};
    if (!gemm_tc_plan(M, N, K, (size_t)1 << 40, &p)) return 0;
    return p.splits > 1 ? (size_t)p.splits * M * N * sizeof(float) : 0;
}

template <bool A_MN, bool B_MN, int UB>
static int launch_gemm_tc_u(const GemmTcArgs& a, const GemmTcPlan& p, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, UB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_gemm_tc)");
        attr_set = true;
    }
    dim3 grid(p.n_tiles, p.m_tiles, p.splits);
    k_gemm_tc<A_MN, B_MN, UB><<<grid, GT_THREADS, p.smem, st>>>(a);
    RAT_CHECK_LAUNCH("k_gemm_tc");
    return RAT_OK;
}
template <bool A_MN, bool B_MN>
static int launch_gemm_tc(const GemmTcArgs& a, const GemmTcPlan& p, cudaStream_t st) {
    // B tile items per thread: K-major BN x 8 chunks, MN-major 64 rows x round_up(BN / 8, 4) chunks; / 256 threads
    const int items = B_MN ? GT_KB * ((p.BN / 8 + 3) & ~3) : round_up(p.BN, 8) * (GT_KB / 8);
    if (items <= 2 * GT_THREADS) return launch_gemm_tc_u<A_MN, B_MN, 2>(a, p, st);
    if (items <= 4 * GT_THREADS) return launch_gemm_tc_u<A_MN, B_MN, 4>(a, p, st);
    return launch_gemm_tc_u<A_MN, B_MN, 8>(a, p, st);
}

// returns RAT_OK if launched (C, or `splits` partials in workspace with *splits_out > 1), 1 if not covered
int gemm_tc_dispatch(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                     int ldc, int trans_a, int trans_b, const float* a_amax, float* workspace, size_t workspace_bytes,
                     int* splits_out, cudaStream_t st) {
    GemmTcPlan p{};
    if (!gemm_tc_plan(M, N, K, workspace ? workspace_bytes : 0, &p)) return 1;
    GemmTcArgs a{A, B, p.splits > 1 ? workspace : C, bias, M, N, K, lda, ldb, p.splits > 1 ? N : ldc, a_amax,
                 p.BN, p.kchunk, p.splits};
    *splits_out = p.splits;
    if (!trans_a && !trans_b) return launch_gemm_tc<false, false>(a, p, st);
    if (!trans_a && trans_b) return launch_gemm_tc<false, true>(a, p, st);
    if (trans_a && !trans_b) return launch_gemm_tc<true, false>(a, p, st);
    return launch_gemm_tc<true, true>(a, p, st);
}
