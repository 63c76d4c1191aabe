// Error state, device queries and ABI versioning for librat_b200.so.
#include "common.cuh"
#include "../../include/rat_b200.h"
#include <stdarg.h>
#include <string.h>

namespace rat {

static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return RAT_ECUDA;
}

static int g_sms = 0, g_smem = 0;
static void query() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_sms = 148; g_smem = 227 * 1024; return; }
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_sms <= 0) g_sms = 148;
    if (g_smem <= 0) g_smem = 227 * 1024;
}
static unsigned int* g_rng_step = nullptr;
const unsigned int* rng_step_ptr() {
    if (!g_rng_step) {
        if (cudaMalloc(&g_rng_step, sizeof(unsigned int)) != cudaSuccess) return nullptr;
        cudaMemset(g_rng_step, 0, sizeof(unsigned int));
    }
    return g_rng_step;
}
__global__ void k_rng_step(unsigned int* p, unsigned int set, int advance) { *p = advance ? *p + 1u : set; }
int num_sms() { if (!g_sms) query(); return g_sms; }
int max_smem_optin() { if (!g_smem) query(); return g_smem; }

}  // namespace rat

extern "C" const char* rat_last_error(void) { return rat::g_err; }
extern "C" int rat_abi_version(void) { return RAT_ABI_VERSION; }
extern "C" long long rat_launch_count(void) { return rat::g_launches; }

extern "C" int rat_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rat::cuda_fail(e, "rat_device_check: no CUDA device");
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        rat::set_error("rat_device_check: device %d is sm_%d%d; librat_b200 contains sm_100a code only", dev, major, minor);
        return RAT_ECUDA;
    }
    return RAT_OK;
}

extern "C" int rat_rng_step_set(unsigned int value, void* stream) {
    unsigned int* p = const_cast<unsigned int*>(rat::rng_step_ptr());
    RAT_REQUIRE(p != nullptr, "rat_rng_step_set: allocation failed");
    rat::k_rng_step<<<1, 1, 0, (cudaStream_t)stream>>>(p, value, 0);
    RAT_CHECK_LAUNCH("k_rng_step");
    return RAT_OK;
}
extern "C" int rat_rng_step_advance(void* stream) {
    unsigned int* p = const_cast<unsigned int*>(rat::rng_step_ptr());
    RAT_REQUIRE(p != nullptr, "rat_rng_step_advance: allocation failed");
    rat::k_rng_step<<<1, 1, 0, (cudaStream_t)stream>>>(p, 0u, 1);
    RAT_CHECK_LAUNCH("k_rng_step");
    return RAT_OK;
}
