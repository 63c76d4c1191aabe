// Micro-timings that shape the attention kernels (run on the B200 box): cost of issuing small tcgen05 MMAs from one warp and
// from four warps of different SM sub-partitions, commit -> mbarrier wake-up latency, named barrier / fence costs.
#include "../www24-rat_b200/csrc/tc5.cuh"
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void bar128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// mode 0: one warp issues NM MMAs (M, N given) then commits, waits.  mode 1: warps 0,1,2,3 (4 sub-partitions) issue NM/4 each.
// mode 2: warps 0,4,8,12 (same sub-partition) issue NM/4 each.
__global__ void __launch_bounds__(512) k_time(int mode, int M, int N, int NM, int reps, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[16];
    __shared__ uint32_t tbase;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) tc5::mbar_init(&bars[i], 1); tc5::fence_mbar_init(); }
    if (warp == 0) tc5::tmem_alloc(&tbase, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t s0 = tc5::smem_u32(smem);
    const uint32_t idesc = tc5::instr_desc(tc5::FMT_F16, M, N);
    int issuer = -1;
    if (mode == 0 && warp == 0) issuer = 0;
    if (mode == 1 && warp < 4) issuer = warp;
    if (mode == 2 && (warp & 3) == 0) issuer = warp >> 2;
    const int per = mode == 0 ? NM : NM / 4;
    long long t_issue = 0, t_total = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        if (issuer >= 0) {
            const long long t0 = clock64();
            for (int i = 0; i < per; ++i)
                tc5::mma_f16_w(tbase + issuer * 128, tc5::smem_desc(s0 + (i & 3) * 2048, 1024, 128), tc5::smem_desc(s0 + 16384, 1024, 128), idesc, i > 0);
            tc5::mma_commit_w(&bars[issuer]);
            const long long t1 = clock64();
            tc5::mbar_wait(&bars[issuer], ph);
            const long long t2 = clock64();
            t_issue += t1 - t0; t_total += t2 - t0;
        }
        ph ^= 1;
    }
    if (issuer >= 0 && (threadIdx.x & 31) == 0) { out[issuer * 2] = t_issue / reps; out[issuer * 2 + 1] = t_total / reps; }
    // sync primitive costs (warp 0..3 = one 128-thread group)
    if (warp < 4) {
        long long t0 = clock64();
        for (int i = 0; i < 64; ++i) bar128(1);
        long long t1 = clock64();
        for (int i = 0; i < 64; ++i) tc5::fence_proxy_async();
        long long t2 = clock64();
        for (int i = 0; i < 64; ++i) { tc5::fence_before_sync(); tc5::fence_after_sync(); }
        long long t3 = clock64();
        // wait on an already-completed barrier phase
        for (int i = 0; i < 64; ++i) tc5::mbar_wait(&bars[0], ph ^ 1);
        long long t4 = clock64();
        for (int i = 0; i < 64; ++i) tc5::mbar_wait_sleep(&bars[0], ph ^ 1);
        long long t5 = clock64();
        float v[16];
        for (int i = 0; i < 64; ++i) { tc5::tmem_ld16(tbase + ((warp * 32) << 16), v); tc5::tmem_ld_wait(); }
        long long t6 = clock64();
        if (threadIdx.x == 0) { out[8] = (t1 - t0) / 64; out[9] = (t2 - t1) / 64; out[10] = (t3 - t2) / 64; out[11] = (t4 - t3) / 64; out[12] = (t5 - t4) / 64; out[13] = (t6 - t5) / 64; out[14] = (long long)v[3]; }
    }
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tbase, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16 * 8);
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int shapes[][2] = {{64, 16}, {64, 64}, {128, 48}, {128, 96}, {128, 128}, {128, 256}};
    for (auto& sh : shapes)
        for (int mode = 0; mode < 3; ++mode)
            for (int NM : {8, 32}) {
                cudaMemset(d, 0, 16 * 8);
                k_time<<<1, 512, 40 * 1024>>>(mode, sh[0], sh[1], NM, 50, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                printf("M=%3d N=%3d K=16  %s  %2d MMAs: issue %5lld cyc (%4.1f / MMA of this warp), issue+complete %5lld cyc", sh[0], sh[1],
                       mode == 0 ? "1 warp           " : mode == 1 ? "4 warps, 4 subpart" : "4 warps, 1 subpart", NM, h[0], (double)h[0] / (mode ? NM / 4 : NM), h[1]);
                if (mode) printf("  [others: %lld/%lld %lld/%lld %lld/%lld]", h[2], h[3], h[4], h[5], h[6], h[7]);
                printf("\n");
                if (mode == 0 && NM == 8 && sh[1] == 16)
                    printf("  bar.sync(128) %lld cyc, fence.proxy.async %lld, tcgen05 fences %lld, mbar wait(done) %lld, mbar wait_sleep(done) %lld, tmem ld16+wait %lld\n",
                           h[8], h[9], h[10], h[11], h[12], h[13]);
            }
    return 0;
}
