// Shared-memory tile primitives used by the fused RAT-block kernels (encoder_fwd.cu / encoder_bwd.cu).
//
// All tiles are fp32, row-major, with leading dimensions that are multiples of 4 floats so that every row
// start is 16-byte aligned for LDS.128.  The per-thread register tile is TM rows x 4 columns; a thread's rows
// are STRIDED by nrt (= ceil(R/TM)) so that neighbouring threads touch neighbouring rows (no 8*lda bank
// aliasing, see DESIGN.md "tile_gemm").
#pragma once
#include "common.cuh"

namespace rat {

struct EpiNone {
    __device__ __forceinline__ void operator()(int, int, float4&) const {}
};

// C[r][c] (+)= sum_k A[r*lda + k] * Bt[k*ldb + c]      r < R, c < Cc (Cc % 4 == 0, padded), k < K
template <int TM, class Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bt,
                                          int ldb, float* __restrict__ C, int ldc, int R, int Cc, int K,
                                          bool accum, Epi epi) {
    const int nct = Cc >> 2;
    const int nrt = (R + TM - 1) / TM;
    const int ntiles = nct * nrt;
    const int K4 = K & ~3;
    for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
        const int ct = tile % nct, rt = tile / nct;
        const int c0 = ct << 2;
        float4 acc[TM];
        const float* ap[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            int r = rt + i * nrt;
            ap[i] = A + (size_t)(r < R ? r : R - 1) * lda;
        }
        const float* bp = Bt + c0;
#pragma unroll 2
        for (int k = 0; k < K4; k += 4) {
            const float4 b0 = *reinterpret_cast<const float4*>(bp + (size_t)(k + 0) * ldb);
            const float4 b1 = *reinterpret_cast<const float4*>(bp + (size_t)(k + 1) * ldb);
            const float4 b2 = *reinterpret_cast<const float4*>(bp + (size_t)(k + 2) * ldb);
            const float4 b3 = *reinterpret_cast<const float4*>(bp + (size_t)(k + 3) * ldb);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(ap[i] + k);
                acc[i].x = fmaf(a.x, b0.x, acc[i].x); acc[i].y = fmaf(a.x, b0.y, acc[i].y);
                acc[i].z = fmaf(a.x, b0.z, acc[i].z); acc[i].w = fmaf(a.x, b0.w, acc[i].w);
                acc[i].x = fmaf(a.y, b1.x, acc[i].x); acc[i].y = fmaf(a.y, b1.y, acc[i].y);
                acc[i].z = fmaf(a.y, b1.z, acc[i].z); acc[i].w = fmaf(a.y, b1.w, acc[i].w);
                acc[i].x = fmaf(a.z, b2.x, acc[i].x); acc[i].y = fmaf(a.z, b2.y, acc[i].y);
                acc[i].z = fmaf(a.z, b2.z, acc[i].z); acc[i].w = fmaf(a.z, b2.w, acc[i].w);
                acc[i].x = fmaf(a.w, b3.x, acc[i].x); acc[i].y = fmaf(a.w, b3.y, acc[i].y);
                acc[i].z = fmaf(a.w, b3.z, acc[i].z); acc[i].w = fmaf(a.w, b3.w, acc[i].w);
            }
        }
        for (int k = K4; k < K; ++k) {
            const float4 b = *reinterpret_cast<const float4*>(bp + (size_t)k * ldb);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float a = ap[i][k];
                acc[i].x = fmaf(a, b.x, acc[i].x); acc[i].y = fmaf(a, b.y, acc[i].y);
                acc[i].z = fmaf(a, b.z, acc[i].z); acc[i].w = fmaf(a, b.w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = rt + i * nrt;
            if (r < R) {
                float4* cp = reinterpret_cast<float4*>(C + (size_t)r * ldc + c0);
                float4 v = acc[i];
                if (accum) { const float4 o = *cp; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                epi(r, c0, v);
                *cp = v;
            }
        }
    }
}

// G[i][j] += sum_r A[r*lda + i] * B[r*ldb + j]     i < Ka (Ka%4==0 padded), j < Kb (Kb%4==0 padded), r < R
// "TN" product used for weight gradients; the 4x4 register tile is accumulated into the CTA-private
// shared-memory gradient accumulator G (each (i,j) is owned by exactly one thread: no atomics).
__device__ __forceinline__ void tile_gemm_tn_acc(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                 int ldb, float* __restrict__ G, int ldg, int R, int Ka, int Kb) {
    const int na = Ka >> 2, nb = Kb >> 2;
    const int ntiles = na * nb;
    for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
        const int jb = tile % nb, ia = tile / nb;
        const float* ap = A + (ia << 2);
        const float* bp = B + (jb << 2);
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(ap + (size_t)r * lda);
            const float4 b = *reinterpret_cast<const float4*>(bp + (size_t)r * ldb);
            acc0.x = fmaf(a.x, b.x, acc0.x); acc0.y = fmaf(a.x, b.y, acc0.y);
            acc0.z = fmaf(a.x, b.z, acc0.z); acc0.w = fmaf(a.x, b.w, acc0.w);
            acc1.x = fmaf(a.y, b.x, acc1.x); acc1.y = fmaf(a.y, b.y, acc1.y);
            acc1.z = fmaf(a.y, b.z, acc1.z); acc1.w = fmaf(a.y, b.w, acc1.w);
            acc2.x = fmaf(a.z, b.x, acc2.x); acc2.y = fmaf(a.z, b.y, acc2.y);
            acc2.z = fmaf(a.z, b.z, acc2.z); acc2.w = fmaf(a.z, b.w, acc2.w);
            acc3.x = fmaf(a.w, b.x, acc3.x); acc3.y = fmaf(a.w, b.y, acc3.y);
            acc3.z = fmaf(a.w, b.z, acc3.z); acc3.w = fmaf(a.w, b.w, acc3.w);
        }
        float4* g = reinterpret_cast<float4*>(G + (size_t)(ia << 2) * ldg + (jb << 2));
        const int s = ldg >> 2;
        float4 t;
        t = g[0];     t.x += acc0.x; t.y += acc0.y; t.z += acc0.z; t.w += acc0.w; g[0] = t;
        t = g[s];     t.x += acc1.x; t.y += acc1.y; t.z += acc1.z; t.w += acc1.w; g[s] = t;
        t = g[2 * s]; t.x += acc2.x; t.y += acc2.y; t.z += acc2.z; t.w += acc2.w; g[2 * s] = t;
        t = g[3 * s]; t.x += acc3.x; t.y += acc3.y; t.z += acc3.z; t.w += acc3.w; g[3 * s] = t;
    }
}

// column sums: gsum[c] += sum_r A[r*lda + c]   (bias / LayerNorm-beta gradients)
__device__ __forceinline__ void tile_colsum_acc(const float* __restrict__ A, int lda, float* __restrict__ gsum,
                                                int R, int Cc) {
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += A[(size_t)r * lda + c];
        gsum[c] += s;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Tensor-core tile GEMM on shared-memory operands: warp-level mma.sync.m16n8k8 TF32 (fp32 accumulate).
//
//   C[m][n] (+)= sum_k A(m,k) * B(k,n)      A(m,k) = A[m*sam + k*sak]   B(k,n) = B[k*sbk + n*sbn]
//
// Arbitrary element strides let ONE natural-layout copy of a weight matrix serve as the n-major operand of the
// forward product and as the k-major operand of the data-gradient product, and let activations be read
// transposed for the weight-gradient product (reduction over token rows) -- no transposed staging copies.
// M is processed in 16-row tiles, N in 8-column tiles (NT of them per warp tile), K in steps of 8:
//   * rows  [M, round_up(M,16)) and cols [N, round_up(N,8)) of the operands must be ALLOCATED and finite;
//   * K8 = round_up(K,8) and the padding k-range must be zero in at least one operand (and finite in both).
// Operands are rounded to TF32 with cvt.rna on fragment load (unbiased), accumulation is fp32.
__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct EpiNone2 {
    __device__ __forceinline__ void operator()(int, int, float&, float&) const {}
};

template <int NT, class Epi>
__device__ __forceinline__ void mma_tile_gemm(const float* __restrict__ A, int sam, int sak,
                                              const float* __restrict__ B, int sbk, int sbn, float* __restrict__ C,
                                              int ldc, int M, int N, int K8, bool accum, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mt = (M + 15) >> 4, n8 = (N + 7) >> 3, ng = (n8 + NT - 1) / NT;
    for (int wt = warp; wt < mt * ng; wt += nwarps) {
        const int m0 = (wt / ng) << 4, nb = (wt % ng) * NT;
        const int nn = min(NT, n8 - nb);
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        // rows >= M are clamped to the last valid row (their results are never stored): no out-of-tile reads
        const float* a_lo = A + (size_t)min(m0 + g, M - 1) * sam + (size_t)t * sak;
        const float* a_hi = A + (size_t)min(m0 + g + 8, M - 1) * sam + (size_t)t * sak;
        const float* b_p = B + (size_t)t * sbk + (size_t)((nb << 3) + g) * sbn;
        for (int k0 = 0; k0 < K8; k0 += 8) {
            unsigned a[4];
            a[0] = f2tf32(a_lo[(size_t)k0 * sak]);
            a[1] = f2tf32(a_hi[(size_t)k0 * sak]);
            a[2] = f2tf32(a_lo[(size_t)(k0 + 4) * sak]);
            a[3] = f2tf32(a_hi[(size_t)(k0 + 4) * sak]);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (j < nn) {
                    const float* bp = b_p + (size_t)k0 * sbk + (size_t)(j << 3) * sbn;
                    const unsigned b0 = f2tf32(bp[0]);
                    const unsigned b1 = f2tf32(bp[(size_t)4 * sbk]);
                    mma_tf32_16x8x8(acc[j], a, b0, b1);
                }
            }
        }
        const bool lo_ok = (m0 + g) < M, hi_ok = (m0 + g + 8) < M;     // rows >= M are computed but never stored
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (j < nn) {
                const int col = ((nb + j) << 3) + 2 * t;
                float2* plo = reinterpret_cast<float2*>(C + (size_t)(m0 + g) * ldc + col);
                float2* phi = reinterpret_cast<float2*>(C + (size_t)(m0 + g + 8) * ldc + col);
                if (lo_ok) {
                    if (accum) { const float2 o = *plo; acc[j][0] += o.x; acc[j][1] += o.y; }
                    epi(m0 + g, col, acc[j][0], acc[j][1]);
                    *plo = make_float2(acc[j][0], acc[j][1]);
                }
                if (hi_ok) {
                    if (accum) { const float2 o = *phi; acc[j][2] += o.x; acc[j][3] += o.y; }
                    epi(m0 + g + 8, col, acc[j][2], acc[j][3]);
                    *phi = make_float2(acc[j][2], acc[j][3]);
                }
            }
        }
    }
}

// exact-fp32 SIMT twin of mma_tile_gemm with the same operand contract (the parity anchor; `precision=fp32`)
template <class Epi>
__device__ __forceinline__ void simt_tile_gemm(const float* __restrict__ A, int sam, int sak,
                                               const float* __restrict__ B, int sbk, int sbn, float* __restrict__ C,
                                               int ldc, int M, int N, int K8, bool accum, Epi epi) {
    const int n2 = (N + 1) >> 1;
    for (int i = threadIdx.x; i < M * n2; i += blockDim.x) {
        const int m = i / n2, n = (i % n2) << 1;
        float c0 = 0.f, c1 = 0.f;
        const float* ap = A + (size_t)m * sam;
        const float* bp = B + (size_t)n * sbn;
        for (int k = 0; k < K8; ++k) {
            const float a = ap[(size_t)k * sak];
            c0 = fmaf(a, bp[(size_t)k * sbk], c0);
            c1 = fmaf(a, bp[(size_t)k * sbk + sbn], c1);
        }
        float* cp = C + (size_t)m * ldc + n;
        if (accum) { c0 += cp[0]; c1 += cp[1]; }
        epi(m, n, c0, c1);
        cp[0] = c0; cp[1] = c1;
    }
}

template <bool MMA, int NT, class Epi>
__device__ __forceinline__ void tc_gemm(const float* A, int sam, int sak, const float* B, int sbk, int sbn, float* C,
                                        int ldc, int M, int N, int K8, bool accum, Epi epi) {
    if (MMA) mma_tile_gemm<NT>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K8, accum, epi);
    else simt_tile_gemm(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K8, accum, epi);
}

__host__ __device__ inline int pad_ld(int cols) { return ((cols + 7) & ~7) + 4; }   // conflict-free fragment loads
__host__ __device__ inline int pad8(int v) { return (v + 7) & ~7; }
__host__ __device__ inline int pad16(int v) { return (v + 15) & ~15; }

// gsum[c] += sum_r A[r][c] with the rows split over thread groups and a fixed-order second stage (deterministic).
// scratch: blockDim.x floats.  Contains two __syncthreads (call from uniform control flow).
__device__ __forceinline__ void tile_colsum_acc2(const float* __restrict__ A, int lda, float* __restrict__ gsum,
                                                 int R, int Cc, float* __restrict__ scratch) {
    const int G = max(1, (int)blockDim.x / Cc);
    const int c = threadIdx.x % Cc, rg = threadIdx.x / Cc;
    if (rg < G) {
        float s = 0.f;
        for (int r = rg; r < R; r += G) s += A[(size_t)r * lda + c];
        scratch[rg * Cc + c] = s;
    }
    __syncthreads();
    if (threadIdx.x < Cc) {
        float s = 0.f;
        for (int q = 0; q < G; ++q) s += scratch[q * Cc + threadIdx.x];
        gsum[threadIdx.x] += s;
    }
    __syncthreads();
}

// sequence geometry: local row (ls, p) of a tile that starts at sequence s0
struct SeqGeom {
    int S;          // positions per sequence
    int mode;       // 0: sequence = (b,t), rows contiguous ; 1: sequence = (b,n), rows strided by N
    int T, N;
    __device__ __forceinline__ long long grow(long long s, int p) const {
        if (mode == 0) return s * S + p;
        const long long b = (s >> 31) == 0 ? (long long)((unsigned int)s / (unsigned int)N) : s / N;   // 32-bit divide when possible
        int n = (int)(s - b * N);
        return b * (long long)T * N + (long long)p * N + n;
    }
};

__device__ __forceinline__ float group_sum(float v, int lg) {
    for (int o = lg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

}  // namespace rat
