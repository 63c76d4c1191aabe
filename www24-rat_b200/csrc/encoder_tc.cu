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
#include "tile.cuh"
#include "encoder_common.cuh"
#include "tc5.cuh"
#include "../../include/rat_b200.h"
#include <cuda_bf16.h>

namespace rat {

int precision_mode();

constexpr int TC_THREADS = 512;
constexpr int TEAM_THREADS = 256;
constexpr int TILE_M = 128;

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TEAM_THREADS) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void sts128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// W [rows x cols] fp32 row-major (torch Linear layout = N x K, K-major)  ->  bf16 canonical image [rows_p x cols_p]
// (rows_p % 8 == 0, cols_p % 16 == 0), zero padded.  One 16-byte chunk (8 bf16) per loop iteration.
__device__ __forceinline__ void stage_weight_image(const float* __restrict__ W, int rows, int cols, int rows_p,
                                                   int cols_p, unsigned char* __restrict__ dst) {
    const int KC = cols_p >> 3;
    const int total = rows_p * KC;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i % rows_p, kc = i / rows_p;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = kc * 8 + k;
            v[k] = (r < rows && c < cols) ? __ldg(W + (size_t)r * cols + c) : 0.f;
        }
        sts128(dst + tc5::kmajor_off(r, kc, KC), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
               pack_bf16(v[6], v[7]));
    }
}

// 8 consecutive floats of a token row, columns [c0, c0+8) clipped to D (zero fill)
template <bool VEC4>
__device__ __forceinline__ void load8(const float* __restrict__ row, int c0, int D, float (&v)[8]) {
    if (VEC4) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + 4 * q < D) t = *reinterpret_cast<const float4*>(row + c0 + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 t = make_float2(0.f, 0.f);
            if (c0 + 2 * q < D) t = *reinterpret_cast<const float2*>(row + c0 + 2 * q);
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    }
}
template <bool VEC4>
__device__ __forceinline__ void store8(float* __restrict__ row, int c0, int D, const float (&v)[8]) {
    if (VEC4) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (c0 + 4 * q < D)
                *reinterpret_cast<float4*>(row + c0 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c0 + 2 * q < D) *reinterpret_cast<float2*>(row + c0 + 2 * q) = make_float2(v[2 * q], v[2 * q + 1]);
    }
}

// Load one token row half (KCH chunks of 8 columns starting at chunk h*KCH), optionally LayerNorm it (the two lanes
// of a row pair exchange partial sums by shuffle), convert to bf16 and store the chunks into the canonical A tile.
//   tid2 = thread index inside the team (0..255): row = tid2 / 2, h = tid2 % 2.   valid=false -> zero row.
template <int KCH, bool VEC4>
__device__ __forceinline__ void stage_row_bf16(const float* __restrict__ src, bool valid, int D, int KC, int row, int h,
                                               const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                               unsigned char* __restrict__ At) {
    float v[KCH][8];
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
        if (valid) load8<VEC4>(src, (h * KCH + j) * 8, D, v[j]);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[j][k] = 0.f;
        }
    }
    if (ln_w != nullptr) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[j][k];                    // pad columns hold zeros
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        const float mean = s / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = (h * KCH + j) * 8 + k;
                const float t = c < D ? v[j][k] - mean : 0.f;
                sq = fmaf(t, t, sq);
            }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        const float rstd = 1.0f / sqrtf(sq / (float)D + 1e-5f);
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = (h * KCH + j) * 8 + k;
                v[j][k] = (c < D && valid) ? (v[j][k] - mean) * rstd * __ldg(ln_w + c) + __ldg(ln_b + c) : 0.f;
            }
    }
#pragma unroll
    for (int j = 0; j < KCH; ++j)
        sts128(At + tc5::kmajor_off(row, h * KCH + j, KC), pack_bf16(v[j][0], v[j][1]), pack_bf16(v[j][2], v[j][3]),
               pack_bf16(v[j][4], v[j][5]), pack_bf16(v[j][6], v[j][7]));
}

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
    const uint32_t idesc1 = tc5::instr_desc(tc5::FMT_BF16, TILE_M, Mp);
    const uint32_t idesc2 = tc5::instr_desc(tc5::FMT_BF16, TILE_M, Np);
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
            stage_row_bf16<KCH, VEC4>(a.x + (r0 + row) * D, row < R, D, KC1, row, h, a.ln_w, a.ln_b, At);
        }
        tc5::fence_proxy_async();
        tc5::fence_before_sync();
        team_sync(team);
        // ---- GEMM1: H[128 x Mp] = A[128 x Kp] . W1^T
        if (tid2 == 0) {
            tc5::fence_after_sync();
            const uint32_t a0 = tc5::smem_u32(At), b0 = tc5::smem_u32(W1i);
            for (int k = 0; k < Kp / 16; ++k)
                tc5::mma_f16(tmem_H, tc5::smem_desc(a0 + k * 256, 128, KC1 * 128), tc5::smem_desc(b0 + k * 256, 128, KC1 * 128),
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
                for (int k = 0; k < 8; ++k) v[k] = gelu_erf(v[k] + b1s[g * 8 + k]);
                // columns >= M: W1 image rows are zero and b1s is zero -> gelu(0) = 0
                sts128(Ht + tc5::kmajor_off(row_e, g, KC2), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                       pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
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
                tc5::mma_f16(tmem_Y, tc5::smem_desc(a0 + k * 256, 128, KC2 * 128), tc5::smem_desc(b0 + k * 256, 128, KC2 * 128),
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

static bool ff_tc_supported(int D, int M) {
    if (D < 2 || (D & 1) || D > 64 || M < 1) return false;
    const int Kp = pad16(D), Mp = pad16(M);
    if (Mp + Kp > 128 || Mp > 256) return false;                     // TMEM columns per team
    if ((Mp >> 3) & 1) return false;                                 // column groups split evenly over two warps
    return true;
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
