// Shared device/host helpers for the RAT hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "rat_b200 kernels are written for sm_100a (B200) only"
#endif

#define RAT_OK 0
#define RAT_EINVAL -1
#define RAT_ECUDA -2
#define RAT_ESMEM -3

namespace rat {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

extern long long g_launches;     // kernels launched by this library (one RAT_CHECK_LAUNCH per launch)

#define RAT_CHECK_LAUNCH(what)                                  \
    do {                                                        \
        ++rat::g_launches;                                      \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return rat::cuda_fail(_e, what); \
    } while (0)

#define RAT_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            rat::set_error(__VA_ARGS__);  \
            return RAT_EINVAL;            \
        }                                 \
    } while (0)

int num_sms();
int max_smem_optin();

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 counter RNG (dropout masks are a pure function of (seed, stream, element)) ----
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// keep-mask scale for flattened element index e: 0 (dropped) or 1/(1-p)
__device__ __forceinline__ float dropout_scale(unsigned long long seed, uint32_t stream, unsigned long long e,
                                               float p, float inv_keep) {
    uint4 r = philox4x32(make_uint4((uint32_t)(e >> 2), (uint32_t)(e >> 34), stream, 0u),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t w = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
    float u = (float)(w >> 8) * (1.0f / 16777216.0f);
    return u < p ? 0.0f : inv_keep;
}
#endif

}  // namespace rat
