// K3/K4: DNN head (MLP_Layer fuxictr/pytorch/layers/deep.py:108-141), BatchNorm1d, ReLU, dropout,
// fc + LR + sigmoid + BCE (RAT_m2.py:144-150, base_model.py:74-77) and their backward.
//
// k_sgemm is the fp32 SIMT shared-memory-tiled GEMM (exact fp32 parity anchor and tf32-mode path); in the default fp16
// tensor-core mode rat_sgemm dispatches to the tcgen05 / TMEM kernel of gemm_tc.cu.  The single-launch BatchNorm /
// head-gradient cluster kernels live in mlp_fused.cu; the split BatchNorm kernels here remain the eval path and the fallback.
#include <algorithm>
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace rat {

// C[m][n] = sum_k opA(m,k) * opB(k,n) (+ bias[n]) ; opA = TA ? A[k*lda+m] : A[m*lda+k] ; opB = TB ? B[k*ldb+n] : B[n*ldb+k]
// grid.z = split-K slices (each slice writes its own [M][N] partial at C + z*M*N when gridDim.z > 1).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, const float* __restrict__ B,
                                               float* __restrict__ C, const float* __restrict__ bias, int M, int N,
                                               int K, int lda, int ldb, int ldc, int kchunk) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    const int tid = threadIdx.x;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- load tiles (bounds-checked scalar loads; coalesced along the contiguous axis of each operand)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;                       // 0..1023
            if (!TA) { int k = e % BK, m = e / BK; int gm = m0 + m, gk = k0 + k;
                       As[k][m] = (gm < M && gk < kend) ? A[(size_t)gm * lda + gk] : 0.f; }
            else     { int m = e % BM, k = e / BM; int gm = m0 + m, gk = k0 + k;
                       As[k][m] = (gm < M && gk < kend) ? A[(size_t)gk * lda + gm] : 0.f; }
            if (!TB) { int k = e % BK, n = e / BK; int gn = n0 + n, gk = k0 + k;
                       Bs[k][n] = (gn < N && gk < kend) ? B[(size_t)gn * ldb + gk] : 0.f; }
            else     { int n = e % BN, k = e / BN; int gn = n0 + n, gk = k0 + k;
                       Bs[k][n] = (gn < N && gk < kend) ? B[(size_t)gk * ldb + gn] : 0.f; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cz = C + (gridDim.z > 1 ? (size_t)blockIdx.z * M * ldc : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + tm + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tn + j;
            if (gn < N) Cz[(size_t)gm * ldc + gn] = acc[i][j] + ((bias && gridDim.z == 1) ? bias[gn] : 0.f);
        }
    }
}

__global__ void k_splitk_reduce(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias,
                                int M, int N, int ldc, int splits) {
    const long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[((size_t)z * M + m) * N + n];   // fixed order: deterministic
        C[(size_t)m * ldc + n] = s + (bias ? bias[n] : 0.f);
    }
}

// ---- degenerate GEMM shapes of the head (N = 1 logit column, M = 1 gradient row, K = 1 outer product) --------------
// C[m] = sum_k opA(m,k) * b_k + bias   : one warp per output row (TA = 0: A row contiguous)
__global__ void __launch_bounds__(256) k_gemv_rows(const float* __restrict__ A, const float* __restrict__ B,
                                                   float* __restrict__ C, const float* __restrict__ bias, int M, int K,
                                                   int lda, long long b_stride, int ldc) {
    const int lane = threadIdx.x & 31;
    for (int m = blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += gridDim.x * 8) {
        const float* a = A + (size_t)m * lda;
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s = fmaf(a[k], __ldg(B + (size_t)k * b_stride), s);
        s = warp_sum(s);
        if (lane == 0) C[(size_t)m * ldc] = s + (bias ? bias[0] : 0.f);
    }
}
// part[slice][n] = sum_{k in slice} a_k * B[k*ldb + n]   (M = 1, B rows indexed by the reduction index)
__global__ void __launch_bounds__(256) k_wcolsum(const float* __restrict__ a, long long a_stride, const float* __restrict__ B,
                                                 int K, int N, int ldb, double* __restrict__ part) {
    __shared__ double s[8][32];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;
    const int per = (K + gridDim.y - 1) / gridDim.y, k0 = blockIdx.y * per, k1 = min(K, k0 + per);
    double acc = 0.0;
    if (n < N) for (int k = k0 + rg; k < k1; k += 8) acc += (double)(a[(size_t)k * a_stride] * B[(size_t)k * ldb + n]);
    s[rg][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rg == 0 && n < N) {
        for (int i = 1; i < 8; ++i) acc += s[i][threadIdx.x];
        part[(size_t)blockIdx.y * N + n] = acc;
    }
}
// C[m][n] = a_m * b_n (+ bias[n])   (K = 1)
__global__ void k_outer(const float* __restrict__ a, long long a_stride, const float* __restrict__ b, long long b_stride,
                        float* __restrict__ C, const float* __restrict__ bias, int M, int N, int ldc) {
    const long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        C[(size_t)m * ldc + n] = a[(size_t)m * a_stride] * __ldg(b + (size_t)n * b_stride) + (bias ? bias[n] : 0.f);
    }
}

// ---- BatchNorm1d -----------------------------------------------------------------------------------------
// per-column sums in double: sums[c] = sum_b z[b][c], sums[C + c] = sum_b z[b][c]^2 (raw sums so that the
// data-parallel path can all-reduce them = SyncBN-equivalent to the single-device reference at global batch)
// Column reductions over the batch: grid (column tiles of 32, row slices).  Every block reduces its row slice in a
// fixed order and writes a partial; k_colred_finalize adds the slices in slice order => bitwise deterministic,
// and the whole machine participates (a single block per column tile would leave 135 SMs idle).
constexpr int COLRED_SLICES = 32;
__global__ void __launch_bounds__(256) k_bn_sums(const float* __restrict__ z, int Bn, int C, double* __restrict__ part) {
    __shared__ double s1[8][32], s2[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;
    const int per = (Bn + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(Bn, r0 + per);
    double a = 0.0, b = 0.0;
    if (c < C)
        for (int r = r0 + rg; r < r1; r += 8) { const double v = z[(size_t)r * C + c]; a += v; b += v * v; }
    s1[rg][threadIdx.x & 31] = a; s2[rg][threadIdx.x & 31] = b;
    __syncthreads();
    if (rg == 0 && c < C) {
        for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
        part[(size_t)blockIdx.y * 2 * C + c] = a; part[(size_t)blockIdx.y * 2 * C + C + c] = b;
    }
}
// out[i] = sum_s part[s][i], i < n  (double or float output)
template <class TO>
__global__ void k_colred_finalize(const double* __restrict__ part, int nslices, int n, TO* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int q = 0; q < nslices; ++q) s += part[(size_t)q * n + i];
    out[i] = (TO)s;
}
// mean / rstd from (all-reduced) sums; running stats update (momentum 0.1, unbiased var), deep.py:128-129
__global__ void k_bn_finalize(const double* __restrict__ sums, double count, int C, float* __restrict__ mean,
                              float* __restrict__ rstd, float* __restrict__ running_mean,
                              float* __restrict__ running_var, float momentum, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
__global__ void k_bn_eval_stats(const float* __restrict__ running_mean, const float* __restrict__ running_var, int C,
                                float* __restrict__ mean, float* __restrict__ rstd, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = running_mean[c];
    rstd[c] = 1.0f / sqrtf(running_var[c] + eps);
}
// out = dropout(relu(bn(z)))  (bn optional: mean == nullptr -> out = dropout(relu(z)))
__global__ void k_bn_act_fwd(const float* __restrict__ z, const float* __restrict__ mean,
                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ out, long long total, int C,
                             float drop_p, unsigned long long seed, unsigned int stream0,
                             const unsigned int* __restrict__ step) {
    const unsigned int stream = rng_stream_of_step(stream0, step);
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = z[i];
        if (mean) v = (v - mean[c]) * rstd[c] * gamma[c] + beta[c];
        v = fmaxf(v, 0.f);
        if (drop_p > 0.f) v *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
        out[i] = v;
    }
}
// backward pass 1: dy = dout * 1[out>0] * dropscale ; sums[c] = sum dy ; sums[C+c] = sum dy * xhat
__global__ void __launch_bounds__(256) k_bn_act_bwd_sums(const float* __restrict__ dout, const float* __restrict__ out,
                                                         const float* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, int Bn, int C, float drop_p,
                                                         unsigned long long seed, unsigned int stream0,
                                                         const unsigned int* __restrict__ step,
                                                         double* __restrict__ sums) {
    const unsigned int stream = rng_stream_of_step(stream0, step);
    __shared__ double s1[8][32], s2[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const int per = (Bn + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(Bn, r0 + per);
    double a = 0.0, b = 0.0;
    if (c < C) {
        const float mu = mean[c], rs = rstd[c];
        for (int r = r0 + rg; r < r1; r += 8) {
            const size_t i = (size_t)r * C + c;
            float dy = out[i] > 0.f ? dout[i] : 0.f;
            if (drop_p > 0.f) dy *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
            a += dy;
            b += (double)dy * (double)((z[i] - mu) * rs);
        }
    }
    s1[rg][threadIdx.x & 31] = a; s2[rg][threadIdx.x & 31] = b;
    __syncthreads();
    if (rg == 0 && c < C) {
        for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
        sums[(size_t)blockIdx.y * 2 * C + c] = a; sums[(size_t)blockIdx.y * 2 * C + C + c] = b;
    }
}
// backward pass 2: dz = gamma*rstd*(dy - dbeta/count - xhat*dgamma/count)   (or dz = dy without bn);
// also emits dgamma/dbeta as floats.
__global__ void k_bn_act_bwd_apply(const float* __restrict__ dout, const float* __restrict__ out,
                                   const float* __restrict__ z, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                   const double* __restrict__ sums, double count, float* __restrict__ dz,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta, long long total, int C,
                                   float drop_p, unsigned long long seed, unsigned int stream0,
                                   const unsigned int* __restrict__ step,
                                   float* __restrict__ dz_amax, float param_grad_scale) {
    const unsigned int stream = rng_stream_of_step(stream0, step);
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float dy = out[i] > 0.f ? dout[i] : 0.f;
        if (drop_p > 0.f) dy *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
        float r = dy;
        if (mean) {
            const float xh = (z[i] - mean[c]) * rstd[c];
            const float db = (float)(sums[c] / count), dg = (float)(sums[C + c] / count);
            r = gamma[c] * rstd[c] * (dy - db - xh * dg);
            if (i < C) { dgamma[c] = (float)sums[C + c] * param_grad_scale; dbeta[c] = (float)sums[c] * param_grad_scale; }
        }
        dz[i] = r;
        amax = fmaxf(amax, fabsf(r));
    }
    if (dz_amax != nullptr) publish_amax_block(dz_amax, amax);     // max|dz| for the fp16 GEMMs that consume dz
}
// out[c] = sum_r A[r][c]   (bias gradients), deterministic
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ A, int R, int C, int lda,
                                                double* __restrict__ part) {
    __shared__ double s[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;
    const int per = (R + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(R, r0 + per);
    double a = 0.0;
    if (c < C) for (int r = r0 + rg; r < r1; r += 8) a += A[(size_t)r * lda + c];
    s[rg][threadIdx.x & 31] = a;
    __syncthreads();
    if (rg == 0 && c < C) {
        for (int i = 1; i < 8; ++i) a += s[i][threadIdx.x];
        part[(size_t)blockIdx.y * C + c] = a;
    }
}

// ---- head: logit = fc(enc[b,0,0,:]) + dnn_out + lr ; sigmoid ; BCE ; dlogit ---------------------------------
// F.binary_cross_entropy clamps log() at -100 and its backward divides by max(y(1-y),1e-12); sigmoid' = y(1-y).
// One WARP per sample (lane = model column, D <= 128): the pooled-token rows are B strided 160-byte reads, which a
// thread-per-sample kernel (16 blocks at B = 4096) turns into a 14 us latency chain between the forward and the backward.
// loss_part [2 * gridDim.x] doubles: per-block BCE sums, then per-block max |dlogit| (for max|denc| = max|dlogit| max|fc_w|).
__global__ void __launch_bounds__(256) k_head(const float* __restrict__ enc, long long enc_stride,
                                              const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                              const float* __restrict__ dnn_out, const float* __restrict__ lr_out,
                                              const float* __restrict__ y_true, int Bn, int D,
                                              float* __restrict__ y_pred, float* __restrict__ dlogit,
                                              float* __restrict__ denc, float inv_count,
                                              double* __restrict__ loss_part) {
    __shared__ double red[8];
    __shared__ float gred[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = lane + 32 * q < D ? fc_w[lane + 32 * q] : 0.f;
    const float bias = fc_b[0];
    double lsum = 0.0;
    float gmax = 0.f;
    for (int b = blockIdx.x * 8 + warp; b < Bn; b += gridDim.x * 8) {
        const float* e = enc + (size_t)b * enc_stride;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < D) s = fmaf(e[lane + 32 * q], w[q], s);
        float logit = warp_sum(s) + bias;                   // xor butterfly: the same bits in every lane
        if (dnn_out) logit += dnn_out[b];
        if (lr_out) logit += lr_out[b];
        const float y = 1.0f / (1.0f + expf(-logit));
        if (lane == 0) y_pred[b] = y;
        if (y_true) {
            const float t = y_true[b];
            const float l1 = fmaxf(logf(y), -100.f), l0 = fmaxf(logf(1.0f - y), -100.f);
            if (lane == 0) lsum += -(double)(t * l1 + (1.0f - t) * l0);
            if (dlogit) {
                const float yy = y * (1.0f - y);
                const float g = (y - t) / fmaxf(yy, 1e-12f) * yy * inv_count;
                if (lane == 0) dlogit[b] = g;
                gmax = fmaxf(gmax, fabsf(g));
                if (denc) {
                    float* de = denc + (size_t)b * enc_stride;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (lane + 32 * q < D) de[lane + 32 * q] = g * w[q];
                }
            }
        }
    }
    if (lane == 0) { red[warp] = lsum; gred[warp] = gmax; }
    __syncthreads();
    if (threadIdx.x == 0 && loss_part) {
        double s = 0.0;
        float m = 0.f;
        for (int i = 0; i < 8; ++i) { s += red[i]; m = fmaxf(m, gred[i]); }
        loss_part[blockIdx.x] = s;
        loss_part[gridDim.x + blockIdx.x] = (double)m;
    }
}
// one warp: block partials in a fixed order (lane-strided, then the xor butterfly) => deterministic
__global__ void k_loss_finalize(const double* __restrict__ part, int n, float inv_count, float* __restrict__ loss_sum,
                                float* __restrict__ loss_mean, const float* __restrict__ fc_w, int D,
                                float* __restrict__ denc_amax) {
    const int lane = threadIdx.x;
    double s = 0.0, m = 0.0;
    for (int i = lane; i < n; i += 32) { s += part[i]; m = fmax(m, part[n + i]); }
    s = warp_sum_d(s);
    float wm = 0.f;
    for (int d = lane; d < D; d += 32) wm = fmaxf(wm, fabsf(fc_w[d]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    }
    if (lane == 0) {
        if (loss_sum) loss_sum[0] = (float)s;
        if (loss_mean) loss_mean[0] = (float)(s * (double)inv_count);
        // denc[b][d] = dlogit[b] * fc_w[d] and float rounding is monotonic: max |denc| == max|dlogit| * max|fc_w| exactly
        if (denc_amax) denc_amax[0] = (float)m * wm;
    }
}

// library-internal scratch for the row-sliced column reductions (COLRED_SLICES x 2C doubles), allocated once per
// process on first use; it never holds user-visible state.
static double* colred_scratch(size_t doubles) {
    static double* buf = nullptr;
    static size_t cap = 0;
    if (doubles > cap) {
        if (buf) cudaFree(buf);
        cap = doubles < (size_t)COLRED_SLICES * 2 * 4096 ? (size_t)COLRED_SLICES * 2 * 4096 : doubles;
        if (cudaMalloc(&buf, cap * sizeof(double)) != cudaSuccess) { buf = nullptr; cap = 0; }
    }
    return buf;
}
static int colred_slices(int rows) { return rows >= 64 * COLRED_SLICES ? COLRED_SLICES : (rows >= 512 ? 8 : 1); }

static int ew_grid(long long total) {
    long long g = (total + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace rat

namespace rat {
// out = max(out, max |x[r*stride + c]|), r < rows, c < cols  (out zeroed by the caller; integer atomicMax on float bits)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, long long rows, int cols, long long stride,
                                                float* __restrict__ out) {
    __shared__ float wm[8];
    const long long total = rows * cols;
    float m = 0.f;
    if (stride == cols) {                                   // contiguous: no index arithmetic, 4 loads in flight
        const long long step = (long long)gridDim.x * blockDim.x;
        long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * step < total; i += 4 * step)
            m = fmaxf(fmaxf(m, fmaxf(fabsf(x[i]), fabsf(x[i + step]))), fmaxf(fabsf(x[i + 2 * step]), fabsf(x[i + 3 * step])));
        for (; i < total; i += step) m = fmaxf(m, fabsf(x[i]));
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / cols;
            m = fmaxf(m, fabsf(x[r * stride + (i - r * cols)]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {                                 // ONE atomic per block: same-address atomics serialise in L2
        for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
        if (m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
    }
}

}  // namespace rat

using namespace rat;

extern "C" int rat_absmax(const float* x, long long rows, int cols, long long row_stride, float* out, void* stream) {
    RAT_REQUIRE(rows > 0 && cols > 0 && row_stride >= cols, "rat_absmax: bad shape");
    const int grid = (int)std::min<long long>((rows * cols + 1023) / 1024, (long long)num_sms() * 4);
    k_absmax<<<std::max(grid, 1), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, row_stride, out);
    RAT_CHECK_LAUNCH("k_absmax");
    return RAT_OK;
}

namespace rat { int precision_mode(); }
size_t gemm_tc_workspace_bytes(int M, int N, int K);
int gemm_tc_dispatch(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                     int ldc, int trans_a, int trans_b, const float* a_amax, float* workspace, size_t workspace_bytes,
                     int* splits_out, cudaStream_t st);

static size_t sgemm_simt_workspace_bytes(int M, int N, int K);
extern "C" size_t rat_sgemm_workspace_bytes(int M, int N, int K) {
    const size_t a = sgemm_simt_workspace_bytes(M, N, K), b = gemm_tc_workspace_bytes(M, N, K);
    return a > b ? a : b;
}
static size_t sgemm_simt_workspace_bytes(int M, int N, int K) {
    // split-K only pays when the output grid under-fills the machine and K is long
    int tiles = ceil_div(M, 64) * ceil_div(N, 64);
    int splits = 1;
    if (tiles < num_sms() && K >= 512) splits = min(16, max(1, (2 * num_sms()) / tiles));
    while (splits > 1 && K / splits < 128) --splits;
    return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}

extern "C" int rat_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                         int ldb, int ldc, int trans_a, int trans_b, float* workspace, size_t workspace_bytes,
                         void* stream) {
    return rat_sgemm_scaled(A, B, C, bias, M, N, K, lda, ldb, ldc, trans_a, trans_b, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int rat_sgemm_scaled(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                                int ldb, int ldc, int trans_a, int trans_b, const float* a_amax, float* workspace,
                                size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(M > 0 && N > 0 && K > 0, "rat_sgemm: bad shape M=%d N=%d K=%d", M, N, K);
    cudaStream_t st0 = (cudaStream_t)stream;
    if (N == 1 && !trans_a && M >= 64) {                    // logit column: C[m] = A[m,:] . b
        k_gemv_rows<<<std::min(ceil_div(M, 8), num_sms() * 8), 256, 0, st0>>>(A, B, C, bias, M, K, lda, trans_b ? ldb : 1, ldc);
        RAT_CHECK_LAUNCH("k_gemv_rows");
        return RAT_OK;
    }
    if (M == 1 && trans_b && K >= 64 && !bias) {                     // gradient row: C[n] = sum_k a_k B[k,n]
        const int ns = colred_slices(K);
        double* part = colred_scratch((size_t)ns * N);
        RAT_REQUIRE(part != nullptr, "rat_sgemm: scratch allocation failed");
        k_wcolsum<<<dim3(ceil_div(N, 32), ns), 256, 0, st0>>>(A, trans_a ? lda : 1, B, K, N, ldb, part);
        RAT_CHECK_LAUNCH("k_wcolsum");
        k_colred_finalize<float><<<ceil_div(N, 128), 128, 0, st0>>>(part, ns, N, C);
        RAT_CHECK_LAUNCH("k_colred_finalize");
        return RAT_OK;
    }
    if (K == 1) {                                           // outer product
        k_outer<<<ew_grid((long long)M * N), 256, 0, st0>>>(A, trans_a ? 1 : lda, B, trans_b ? 1 : ldb, C, bias, M, N, ldc);
        RAT_CHECK_LAUNCH("k_outer");
        return RAT_OK;
    }
    if (precision_mode() == 2) {        // tcgen05 path (fp16 operands, fp32 accumulate); tiny shapes stay on the SIMT kernel
        int tc_splits = 1;
        const int rc = gemm_tc_dispatch(A, B, C, bias, M, N, K, lda, ldb, ldc, trans_a, trans_b, a_amax, workspace, workspace_bytes,
                                        &tc_splits, (cudaStream_t)stream);
        if (rc < 0) return rc;
        if (rc == 0) {
            if (tc_splits > 1) {
                k_splitk_reduce<<<ew_grid((long long)M * N), 256, 0, (cudaStream_t)stream>>>(workspace, C, bias, M, N, ldc, tc_splits);
                RAT_CHECK_LAUNCH("k_splitk_reduce");
            }
            return RAT_OK;
        }
    }
    size_t need = sgemm_simt_workspace_bytes(M, N, K);
    int splits = 1;
    if (need > 0 && workspace && workspace_bytes >= need) splits = (int)(need / ((size_t)M * N * sizeof(float)));
    const int kchunk = round_up(ceil_div(K, splits), 16);
    splits = ceil_div(K, kchunk);
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64), splits);
    cudaStream_t st = (cudaStream_t)stream;
    float* out = splits > 1 ? workspace : C;
    const int ldo = splits > 1 ? N : ldc;
    if (!trans_a && !trans_b) k_sgemm<false, false><<<grid, 256, 0, st>>>(A, B, out, bias, M, N, K, lda, ldb, ldo, kchunk);
    else if (!trans_a && trans_b) k_sgemm<false, true><<<grid, 256, 0, st>>>(A, B, out, bias, M, N, K, lda, ldb, ldo, kchunk);
    else if (trans_a && !trans_b) k_sgemm<true, false><<<grid, 256, 0, st>>>(A, B, out, bias, M, N, K, lda, ldb, ldo, kchunk);
    else k_sgemm<true, true><<<grid, 256, 0, st>>>(A, B, out, bias, M, N, K, lda, ldb, ldo, kchunk);
    RAT_CHECK_LAUNCH("k_sgemm");
    if (splits > 1) {
        k_splitk_reduce<<<ew_grid((long long)M * N), 256, 0, st>>>(workspace, C, bias, M, N, ldc, splits);
        RAT_CHECK_LAUNCH("k_splitk_reduce");
    }
    return RAT_OK;
}

extern "C" int rat_bn_sums(const float* z, int rows, int C, double* sums, void* stream) {
    RAT_REQUIRE(rows > 0 && C > 0, "rat_bn_sums: bad shape");
    const int ns = colred_slices(rows);
    double* part = colred_scratch((size_t)ns * 2 * C);
    RAT_REQUIRE(part != nullptr, "rat_bn_sums: scratch allocation failed");
    k_bn_sums<<<dim3(ceil_div(C, 32), ns), 256, 0, (cudaStream_t)stream>>>(z, rows, C, part);
    RAT_CHECK_LAUNCH("k_bn_sums");
    k_colred_finalize<double><<<ceil_div(2 * C, 128), 128, 0, (cudaStream_t)stream>>>(part, ns, 2 * C, sums);
    RAT_CHECK_LAUNCH("k_colred_finalize");
    return RAT_OK;
}

extern "C" int rat_bn_finalize(const double* sums, double count, int C, float* mean, float* rstd,
                               float* running_mean, float* running_var, float momentum, float eps, void* stream) {
    k_bn_finalize<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, C, mean, rstd, running_mean,
                                                                      running_var, momentum, eps);
    RAT_CHECK_LAUNCH("k_bn_finalize");
    return RAT_OK;
}

extern "C" int rat_bn_eval_stats(const float* running_mean, const float* running_var, int C, float* mean, float* rstd,
                                 float eps, void* stream) {
    k_bn_eval_stats<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, C, mean, rstd, eps);
    RAT_CHECK_LAUNCH("k_bn_eval_stats");
    return RAT_OK;
}

extern "C" int rat_bn_act_fwd(const float* z, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* out, int rows, int C, float drop_p, unsigned long long seed,
                              unsigned int rng_stream, void* stream) {
    const long long total = (long long)rows * C;
    k_bn_act_fwd<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(z, mean, rstd, gamma, beta, out, total, C, drop_p,
                                                                   seed, rng_stream, rng_step_ptr());
    RAT_CHECK_LAUNCH("k_bn_act_fwd");
    return RAT_OK;
}

extern "C" int rat_bn_act_bwd_sums(const float* dout, const float* out, const float* z, const float* mean,
                                   const float* rstd, int rows, int C, float drop_p, unsigned long long seed,
                                   unsigned int rng_stream, double* sums, void* stream) {
    const int ns = colred_slices(rows);
    double* part = colred_scratch((size_t)ns * 2 * C);
    RAT_REQUIRE(part != nullptr, "rat_bn_act_bwd_sums: scratch allocation failed");
    k_bn_act_bwd_sums<<<dim3(ceil_div(C, 32), ns), 256, 0, (cudaStream_t)stream>>>(dout, out, z, mean, rstd, rows, C,
                                                                                   drop_p, seed, rng_stream, rng_step_ptr(), part);
    RAT_CHECK_LAUNCH("k_bn_act_bwd_sums");
    k_colred_finalize<double><<<ceil_div(2 * C, 128), 128, 0, (cudaStream_t)stream>>>(part, ns, 2 * C, sums);
    RAT_CHECK_LAUNCH("k_colred_finalize");
    return RAT_OK;
}

extern "C" int rat_bn_act_bwd_apply(const float* dout, const float* out, const float* z, const float* mean,
                                    const float* rstd, const float* gamma, const double* sums, double count,
                                    float* dz, float* dgamma, float* dbeta, int rows, int C, float drop_p,
                                    unsigned long long seed, unsigned int rng_stream, float* dz_amax,
                                    float param_grad_scale, void* stream) {
    const long long total = (long long)rows * C;
    k_bn_act_bwd_apply<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(dout, out, z, mean, rstd, gamma, sums, count,
                                                                         dz, dgamma, dbeta, total, C, drop_p, seed,
                                                                         rng_stream, rng_step_ptr(), dz_amax, param_grad_scale);
    RAT_CHECK_LAUNCH("k_bn_act_bwd_apply");
    return RAT_OK;
}

extern "C" int rat_colsum(const float* A, int rows, int C, int lda, float* out, void* stream) {
    const int ns = colred_slices(rows);
    double* part = colred_scratch((size_t)ns * C);
    RAT_REQUIRE(part != nullptr, "rat_colsum: scratch allocation failed");
    k_colsum<<<dim3(ceil_div(C, 32), ns), 256, 0, (cudaStream_t)stream>>>(A, rows, C, lda, part);
    RAT_CHECK_LAUNCH("k_colsum");
    k_colred_finalize<float><<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(part, ns, C, out);
    RAT_CHECK_LAUNCH("k_colred_finalize");
    return RAT_OK;
}

extern "C" int rat_head_blocks(int B) { return min(ceil_div(B, 8), 2 * num_sms()); }

extern "C" int rat_head(const float* enc, long long enc_stride, const float* fc_w, const float* fc_b,
                        const float* dnn_out, const float* lr_out, const float* y_true, int B, int D, float* y_pred,
                        float* dlogit, float* denc, float inv_count, double* loss_part, float* loss_sum,
                        float* loss_mean, float* denc_amax, void* stream) {
    RAT_REQUIRE(B > 0 && D > 0 && D <= 128, "rat_head: bad shape B=%d D=%d (D <= 128)", B, D);
    RAT_REQUIRE(denc_amax == nullptr || (loss_part && denc), "rat_head: denc_amax needs loss_part and denc");
    const int grid = rat_head_blocks(B);
    cudaStream_t st = (cudaStream_t)stream;
    k_head<<<grid, 256, 0, st>>>(enc, enc_stride, fc_w, fc_b, dnn_out, lr_out, y_true, B, D, y_pred, dlogit, denc,
                                 inv_count, loss_part);
    RAT_CHECK_LAUNCH("k_head");
    if (loss_part && (loss_sum || loss_mean || denc_amax)) {
        k_loss_finalize<<<1, 32, 0, st>>>(loss_part, grid, 1.0f / (float)B, loss_sum, loss_mean, fc_w, D, denc_amax);
        RAT_CHECK_LAUNCH("k_loss_finalize");
    }
    return RAT_OK;
}
