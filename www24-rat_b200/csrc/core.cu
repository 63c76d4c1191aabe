// Error state, device queries and ABI versioning for librat_b200.so.
#include "common.cuh"
#include "../../include/rat_b200.h"
#include <stdarg.h>
#include <stdlib.h>
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
bool pdl_enabled() {
    static const bool on = !(getenv("RAT_PDL") && getenv("RAT_PDL")[0] == '0');
    return on;
}
int num_sms() { if (!g_sms) query(); return g_sms; }
int max_smem_optin() { if (!g_smem) query(); return g_smem; }

// ---- deferred record reductions (see common.cuh)
namespace {
constexpr int RF_EVENTS = 32, RF_BUSY = 8;
struct ReduceForkState {
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev[RF_EVENTS] = {};
    int next = 0;
    struct Busy { const void* ws; cudaEvent_t ev; bool valid; } busy[RF_BUSY] = {};
    cudaEvent_t join_ev = nullptr;
} g_rf;
cudaEvent_t rf_make(cudaEvent_t& e) {
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}
}  // namespace
cudaStream_t reduce_fork(cudaStream_t main, const void*) {
    if (!g_rf.side || g_rf.side == main) return main;
    cudaEvent_t e = rf_make(g_rf.fork_ev[g_rf.next]);
    g_rf.next = (g_rf.next + 1) % RF_EVENTS;
    if (!e || cudaEventRecord(e, main) != cudaSuccess || cudaStreamWaitEvent(g_rf.side, e, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return main;                                  // could not fork: stay in order on the caller's stream
    }
    return g_rf.side;
}
void reduce_forked(cudaStream_t launched_on, cudaStream_t main, const void* ws) {
    if (launched_on == main) return;
    int slot = -1;
    for (int i = 0; i < RF_BUSY; ++i) if (g_rf.busy[i].valid && g_rf.busy[i].ws == ws) slot = i;
    for (int i = 0; i < RF_BUSY && slot < 0; ++i) if (!g_rf.busy[i].valid) slot = i;
    if (slot < 0) {                                   // table full: join everything, then reuse slot 0
        cudaEvent_t j = rf_make(g_rf.join_ev);
        cudaEventRecord(j, launched_on);
        cudaStreamWaitEvent(main, j, 0);
        for (int i = 0; i < RF_BUSY; ++i) g_rf.busy[i].valid = false;
        return;
    }
    cudaEvent_t e = rf_make(g_rf.busy[slot].ev);
    cudaEventRecord(e, launched_on);
    g_rf.busy[slot].ws = ws;
    g_rf.busy[slot].valid = true;
}
void reduce_ws_acquire(cudaStream_t main, const void* ws) {
    for (int i = 0; i < RF_BUSY; ++i)
        if (g_rf.busy[i].valid && g_rf.busy[i].ws == ws) {
            cudaStreamWaitEvent(main, g_rf.busy[i].ev, 0);
            g_rf.busy[i].valid = false;
        }
}

}  // namespace rat

extern "C" int rat_set_reduce_stream(void* stream) {
    rat::g_rf.side = (cudaStream_t)stream;
    return RAT_OK;
}
extern "C" int rat_reduce_stream_join(void* stream) {
    using namespace rat;
    bool any = false;
    for (int i = 0; i < RF_BUSY; ++i) any = any || g_rf.busy[i].valid;
    if (g_rf.side && any && g_rf.side != (cudaStream_t)stream) {
        cudaEvent_t j = rf_make(g_rf.join_ev);
        RAT_REQUIRE(j != nullptr, "rat_reduce_stream_join: event creation failed");
        cudaError_t e = cudaEventRecord(j, g_rf.side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)stream, j, 0);
        if (e != cudaSuccess) return cuda_fail(e, "rat_reduce_stream_join");
    }
    for (int i = 0; i < RF_BUSY; ++i) g_rf.busy[i].valid = false;
    return RAT_OK;
}

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
