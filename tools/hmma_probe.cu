// Micro-timing (run on the B200 box): issue rate of the legacy warp-level tensor path (mma.sync.m16n8k16 f16 -> HMMA) per SM
// sub-partition, alone and mixed with the shared-memory fragment loads / conversions a register-resident attention core
// needs.  Decides whether per-warp mma.sync projections (K <= 48, N = 16 per head) are tensor- or issue-bound.
//   hmma_probe            prints cycles per HMMA per sub-partition for 1..4 warps per sub-partition
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// mode 0: NCH independent accumulator chains, operands in registers.  mode 1: every pair of HMMAs is fed by one LDS.128
// (B fragments from a fragment-ordered shared image).  mode 2: mode 1 + 2 cvt.f16x2 packs per HMMA (accumulator -> operand).
template <int NCH>
__global__ void __launch_bounds__(512) k_hmma(int mode, int iters, long long* out, float* sink) {
    __shared__ uint4 img[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) img[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    __syncthreads();
    float acc[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) mma16816(acc[j], a, b0, b1);
        } else {
#pragma unroll
            for (int j = 0; j < NCH; j += 2) {
                const uint4 f = img[((it + j) & 31) * 32 + lane];
                if (mode == 2) {
                    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[0]) : "f"(acc[j][1]), "f"(acc[j][0]));
                    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[1]) : "f"(acc[j][3]), "f"(acc[j][2]));
                    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[2]) : "f"(acc[j + 1][1]), "f"(acc[j + 1][0]));
                    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[3]) : "f"(acc[j + 1][3]), "f"(acc[j + 1][2]));
                }
                mma16816(acc[j], a, f.x, f.y);
                mma16816(acc[j + 1], a, f.z, f.w);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    long long* d_out; float* d_sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    printf("cycles per HMMA (m16n8k16 f16, fp32 acc) per SM sub-partition; all 148 SMs busy\n");
    for (int mode = 0; mode < 3; ++mode)
        for (int warps_per_sp = 1; warps_per_sp <= 4; ++warps_per_sp) {
            const int threads = warps_per_sp * 4 * 32;
            for (int rep = 0; rep < 2; ++rep) k_hmma<8><<<148, threads>>>(mode, iters, d_out, d_sink);
            cudaDeviceSynchronize();
            long long cyc = 0;
            cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
            const double per = (double)cyc / ((double)iters * 8 * warps_per_sp);
            printf("mode %d (%s)  warps/sub-partition %d : %.2f cycles/HMMA  (%.0f dense fp16 TFLOP/s at 1.965 GHz over 148 SMs)\n", mode,
                   mode == 0 ? "registers only" : mode == 1 ? "LDS.128 per 2 HMMA" : "LDS.128 + 4 cvt per 2 HMMA", warps_per_sp, per,
                   16.0 * 8 * 16 * 2 / per * 4 * 148 * 1.965e9 / 1e12);
        }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
