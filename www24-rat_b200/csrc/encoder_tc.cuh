// Shared device helpers of the tcgen05 RAT-block kernels (encoder_tc_*.cu): 16-bit operand staging (fp16 by default, bf16 with
// RAT_TC_FP16=0; "bf16" in older comments below means "the 16-bit operand type") in the UMMA canonical layout, ldmatrix /
// mma.sync fragments, the register-resident weight-gradient jobs, fast exact-erf GELU.
#pragma once
#include "tile.cuh"
#include "encoder_common.cuh"
#include "tc5.cuh"
#include "../../include/rat_b200.h"
#include <cuda_bf16.h>
#include <algorithm>

namespace rat {

int precision_mode();

constexpr int TC_THREADS = 512;
constexpr int TEAM_THREADS = 256;
constexpr int TILE_M = 128;

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TEAM_THREADS) : "memory");
}
// 16-bit operand format of every tcgen05 / mma.sync product of the tensor-core path.  fp16 (default) has the 10-bit
// mantissa of TF32, i.e. 8x less operand rounding than bf16; its narrow exponent range is handled by lifting every
// gradient-domain tile by a per-launch power of two derived from max|dout| (grad_scale_from_amax, common.cuh): the
// backward is linear in dout, so dout is scaled while it is staged and dx / the weight-gradient records are unscaled
// in fp32 on the way out.  RAT_TC_FP16=0 selects bf16 operands (no scaling needed).
#ifndef RAT_TC_FP16
#define RAT_TC_FP16 1
#endif
#if RAT_TC_FP16
constexpr uint32_t TC_FMT = tc5::FMT_F16;
constexpr uint32_t TC_ONES2 = 0x3C003C00u;           // {1.0, 1.0}
#else
constexpr uint32_t TC_FMT = TC_FMT;
constexpr uint32_t TC_ONES2 = 0x3F803F80u;
#endif
__device__ __forceinline__ float tc_grad_scale(const float* amax) {
#if RAT_TC_FP16
    return grad_scale_from_amax(amax);
#else
    return 1.0f;
#endif
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
#if RAT_TC_FP16
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#endif
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
        sts128(dst + tc5::kmajor_off(r, kc, rows_p), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
               pack_h2(v[6], v[7]));
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
// The load and the (LayerNorm, convert, store) halves are separate so that a kernel can issue the loads of its NEXT tile
// early (registers) and finish them when the tile starts.
template <int KCH, bool VEC4>
__device__ __forceinline__ void stage_row_load(const float* __restrict__ src, bool valid, int D, int h, float (&v)[KCH][8]) {
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
        if (valid) load8<VEC4>(src, (h * KCH + j) * 8, D, v[j]);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[j][k] = 0.f;
        }
    }
}
template <int KCH>
__device__ __forceinline__ void stage_row_finish(float (&v)[KCH][8], bool valid, int D, int row, int h,
                                                 const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                 unsigned char* __restrict__ At, int ones_col = -1,
                                                 float* __restrict__ stats = nullptr, float mul = 1.0f) {
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
        if (stats != nullptr && h == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = (h * KCH + j) * 8 + k;
                v[j][k] = (c < D && valid) ? (v[j][k] - mean) * rstd * __ldg(ln_w + c) + __ldg(ln_b + c) : 0.f;
            }
    } else if (mul != 1.0f) {
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) v[j][k] *= mul;
    }
    if (ones_col >= 0 && valid) {        // bias-gradient trick: a column of ones turns colsum(g) into a GEMM column
#pragma unroll
        for (int j = 0; j < KCH; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((h * KCH + j) * 8 + k == ones_col) v[j][k] = 1.0f;
    }
#pragma unroll
    for (int j = 0; j < KCH; ++j)
        sts128(At + tc5::toff(row, h * KCH + j), pack_h2(v[j][0], v[j][1]), pack_h2(v[j][2], v[j][3]),
               pack_h2(v[j][4], v[j][5]), pack_h2(v[j][6], v[j][7]));
}
template <int KCH, bool VEC4>
__device__ __forceinline__ void stage_row_h(const float* __restrict__ src, bool valid, int D, int KC, int row, int h,
                                               const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                               unsigned char* __restrict__ At, int ones_col = -1,
                                               float* __restrict__ stats = nullptr, float mul = 1.0f) {
    float v[KCH][8];
    stage_row_load<KCH, VEC4>(src, valid, D, h, v);
    stage_row_finish<KCH>(v, valid, D, row, h, ln_w, ln_b, At, ones_col, stats, mul);
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_h_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if RAT_TC_FP16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}
// fp16 ACCUMULATORS (two packed registers: c[0] = row g, columns 2t, 2t+1 ; c[1] = row g + 8): the result is directly the
// packed operand fragment of a following product -- no cvt.f16x2 (F2FP runs on the XU pipe at 16 cycles per warp instruction
// on B200 and was the busiest unit of the register-resident attention kernels).  Only the fp16 build has it.
__device__ __forceinline__ void mma_hh_16x8x16(uint32_t (&c)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                 : "+r"(c[0]), "+r"(c[1])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float qmax(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float qsum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Per-lane constants of the warp-level attention cores.  A warp task is one sequence (8 < S <= 16) or a PAIR of
// sequences (S <= 8: rows 0-7 of the m16 fragment = first sequence, rows 8-15 = second, off-diagonal blocks masked).
// Everything that depends only on (S, lane) is computed once per kernel; a task adds its base row.
struct CoreLane {
    uint32_t a_off;       // A-fragment / transposed-B ldmatrix pattern: (relative row)*16 + chunk select
    uint32_t b_off;       // non-transposed B ldmatrix pattern
    int lo_rel, hi_rel;   // relative rows of fragment rows g and g+8 (-1: do not exist for this S)
    float madd[2][4];     // additive score mask: 0 or -inf
    bool packed;
};
__device__ __forceinline__ CoreLane make_core_lane(int S, int lane) {
    CoreLane c;
    c.packed = S <= 8;
    const bool packed = c.packed;
    const int g = lane >> 2, t = lane & 3;
    auto rel = [&](int i) { const int pos = packed ? (i & 7) : i; const int sq = packed ? (i >> 3) : 0; return pos < S ? sq * S + pos : -1; };
    const int ra = rel((lane & 7) + ((lane >> 3) & 1) * 8), rb = rel((lane & 7) + (lane >> 4) * 8);
    c.a_off = (uint32_t)(ra < 0 ? 0 : ra) * 16u + (uint32_t)(lane >> 4) * tc5::TILE_CHUNK;
    c.b_off = (uint32_t)(rb < 0 ? 0 : rb) * 16u + (uint32_t)((lane >> 3) & 1) * tc5::TILE_CHUNK;
    c.lo_rel = rel(g); c.hi_rel = rel(g + 8);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = (e < 2) ? g : g + 8, j = 8 * nt + 2 * t + (e & 1);
            const bool ok = rel(j) >= 0 && (!packed || ((i >> 3) == (j >> 3)));
            c.madd[nt][e] = ok ? 0.f : -INFINITY;
        }
    // make the constants opaque: at the register cap the compiler otherwise REMATERIALISES this whole function inside
    // every warp task (5 % of the attention-backward kernel's instructions) instead of keeping / spilling 12 values
    asm volatile("" : "+r"(c.a_off), "+r"(c.b_off), "+r"(c.lo_rel), "+r"(c.hi_rel));
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) asm volatile("" : "+f"(c.madd[nt][e]));
    return c;
}

// ------------------------------------------------------------------------------------------------ FF backward
// Exact-erf GELU and its derivative from ONE exponential (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7):
//   e = exp(-z^2/2) ; Phi(|z|) = 1 - 0.5 poly(t) e, t = 1/(1 + p |z|/sqrt2) ; gelu = z Phi(z) ; gelu' = Phi(z) + z e/sqrt(2 pi)
__device__ __forceinline__ void gelu_fast(float z, float& g, float& dg) {
    const float az = fabsf(z);
    const float e = ex2f(-0.72134752044448170368f * z * z);          // exp(-z^2/2)
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, az, 1.0f)));   // p/sqrt2 = 0.3275911/1.41421356
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float q = 0.5f * poly * e;                                  // 1 - Phi(|z|)
    const float phi = z >= 0.f ? 1.0f - q : q;
    g = z * phi;
    dg = fmaf(z * 0.39894228040143267794f, e, phi);
}

// C[16 x 16] (+)= sum over the 128 tile rows r of A[r][m0 + 0..15] * B[r][n0 + 0..15]   (A, B: bf16 canonical tiles,
// rows = tokens).  Weight-gradient product: the token dimension is the reduction index, so both fragments are
// ldmatrix.trans loads straight from the K-major activation tiles.  acc[0] = columns n0..n0+7, acc[1] = n0+8..n0+15.
__device__ __forceinline__ void wgrad_job(const unsigned char* __restrict__ At, int KCa, int mchunk, bool a_ones,
                                          const unsigned char* __restrict__ Bt, int KCb, int nchunk, int lane,
                                          float (&acc)[2][4]) {
    const uint32_t a_s = tc5::smem_u32(At), b_s = tc5::smem_u32(Bt);
    const int ra = (lane & 7) + ((lane >> 4) & 1) * 8, ca = mchunk + ((lane >> 3) & 1);
    const int rb = (lane & 7) + ((lane >> 3) & 1) * 8, cb = nchunk + (lane >> 4);
#pragma unroll
    for (int ks = 0; ks < TILE_M / 16; ++ks) {
        uint32_t af[4], bf[4];
        if (a_ones) af[0] = af[1] = af[2] = af[3] = TC_ONES2;
        else ldsm_x4_t(af, a_s + tc5::toff(16 * ks + ra, ca));
        ldsm_x4_t(bf, b_s + tc5::toff(16 * ks + rb, cb));
        mma_h_16x8x16(acc[0], af, bf[0], bf[1]);
        mma_h_16x8x16(acc[1], af, bf[2], bf[3]);
    }
}
// store a job's accumulators into rec[(m0 + row) * ld + n0 + col] (fp32, row-major)
__device__ __forceinline__ void wgrad_store(float* __restrict__ rec, int ld, int m0, int n0, int lane,
                                            const float (&acc)[2][4], bool first_row_only, float mul = 1.0f) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const int col = n0 + 8 * nt + 2 * t;
        if (!first_row_only || g == 0) {
            rec[(size_t)(m0 + g) * ld + col] = acc[nt][0] * mul;
            rec[(size_t)(m0 + g) * ld + col + 1] = acc[nt][1] * mul;
        }
        if (!first_row_only) {
            rec[(size_t)(m0 + g + 8) * ld + col] = acc[nt][2] * mul;
            rec[(size_t)(m0 + g + 8) * ld + col + 1] = acc[nt][3] * mul;
        }
    }
}


// Fixed-order sum of the per-CTA gradient records: blockDim = (32 outputs, 8 record slices).  Thread (tx, ty) adds records
// ty, ty+8, ... of output tx into four interleaved accumulators (four loads in flight; consecutive tx read
// consecutive floats), the 8 slice sums are combined in slice order through shared memory.  The association depends
// only on nparts => bitwise deterministic.  Returns the total in slice 0 (other slices return 0); contains
// __syncthreads, so every thread of the block must call it (active = false contributes nothing).
__device__ __forceinline__ float record_sum_sliced(const float* __restrict__ partials, size_t psize, int nparts, size_t src,
                                                   bool active) {
    __shared__ float red[8][33];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        int c = threadIdx.y;
        for (; c + 24 < nparts; c += 32) {
            const float v0 = partials[(size_t)c * psize + src], v1 = partials[(size_t)(c + 8) * psize + src];
            const float v2 = partials[(size_t)(c + 16) * psize + src], v3 = partials[(size_t)(c + 24) * psize + src];
            acc[0] += v0; acc[1] += v1; acc[2] += v2; acc[3] += v3;
        }
        for (int k = 0; c < nparts; c += 8, ++k) acc[k] += partials[(size_t)c * psize + src];
    }
    __syncthreads();                                   // red of the previous call consumed
    red[threadIdx.y][threadIdx.x] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    float s = 0.f;
    if (threadIdx.y == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    }
    return s;
}

static inline bool ff_tc_supported(int D, int M) {
    if (D < 2 || (D & 1) || D > 64 || M < 1) return false;
    const int Kp = pad16(D), Mp = pad16(M);
    if (Mp + Kp > 128 || Mp > 256) return false;                     // TMEM columns per team
    return true;
}

}  // namespace rat
