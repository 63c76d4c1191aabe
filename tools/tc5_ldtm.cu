// TMEM -> register bandwidth (tcgen05.ld.32x32b) as a function of the number of warps and the vector width.
#include "../www24-rat_b200/csrc/tc5.cuh"
#include <cstdio>
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int W>
__global__ void __launch_bounds__(512) k_ld(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc5::tmem_alloc(&tbase, 512);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t base = tbase + (((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (uint32_t)((i * W) & 255) + (warp >> 2) * 0;
            if (W == 8) { float v[8]; tc5::tmem_ld8(base + col, v); tc5::tmem_ld_wait(); for (int k = 0; k < 8; ++k) acc += v[k]; }
            if (W == 16) { float v[16]; tc5::tmem_ld16(base + col, v); tc5::tmem_ld_wait(); for (int k = 0; k < 16; ++k) acc += v[k]; }
            if (W == 32) { float v[32]; ld32(base + col, v); tc5::tmem_ld_wait(); for (int k = 0; k < 32; ++k) acc += v[k]; }
            if (W == 64) { float v[32], w[32]; ld32(base + col, v); ld32(base + ((col + 32) & 255), w); tc5::tmem_ld_wait(); for (int k = 0; k < 32; ++k) acc += v[k] + w[k]; }
        }
    }
    const long long t1 = clock64();
    if (acc == 12345.f) sink[0] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tbase, 512);
}
int main() {
    long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    for (int W : {8, 16, 32, 64})
        for (int nw : {1, 4, 8, 16}) {
            if (W == 8) k_ld<8><<<1, 512>>>(nw, iters, d, s);
            if (W == 16) k_ld<16><<<1, 512>>>(nw, iters, d, s);
            if (W == 32) k_ld<32><<<1, 512>>>(nw, iters, d, s);
            if (W == 64) k_ld<64><<<1, 512>>>(nw, iters, d, s);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * iters * 32 * W * 4;
            printf("ld x%-2d (+wait, +%d adds) %2d warps: %6.1f cycles per load per warp, %7.1f B/clk per SM\n", W, W, nw, (double)h / iters, bytes / h);
        }
    return 0;
}
