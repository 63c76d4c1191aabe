// Small-message collectives over NVLink peer memory (symmetric buffers), written for the latency-bound exchanges of the
// data-parallel step: the raw BatchNorm sums (2 x 400 doubles, 3 forward + 3 backward per step) and the shard-norm
// partials.  An NCCL all-reduce of such a message costs ~25 us of launch + protocol latency on the compute stream; a
// one-shot exchange (every rank stores its vector into every peer's slot, one flag per peer, local fixed-order sum) is a
// single ~5 us kernel and captures into the CUDA graph of the training step (the call counter lives on the device).
//
// Symmetric buffer of a rank (allocated by the caller, zero-initialised once, rat_oneshot_workspace_bytes):
//   [0]            uint32 seq            this rank's call counter
//   [64 ...]       uint32 flags[2][64]   flags[parity][src] = seq of the last vector src has delivered into slots[parity][src]
//   [1024 ...]     double slots[2][world][RAT_ONESHOT_MAX]
// Two slot sets alternate by call parity: a rank can only be one call ahead of the slowest rank (it needs everyone's flag
// of call k to finish call k), so set k%2 is never overwritten while a peer still reads call k.
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace rat {

constexpr int ONESHOT_MAX = 2048;          // doubles per message
constexpr int ONESHOT_HDR = 1024;          // bytes before the slots

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_oneshot_allreduce_f64(unsigned char* const* __restrict__ peers, int rank, int world,
                                                               const double* __restrict__ in, int n, double* __restrict__ out) {
    __shared__ unsigned int seq_s;
    unsigned char* mine = peers[rank];
    if (threadIdx.x == 0) {
        unsigned int* seqp = reinterpret_cast<unsigned int*>(mine);
        seq_s = *seqp + 1u;
        *seqp = seq_s;
    }
    __syncthreads();
    const unsigned int seq = seq_s, par = seq & 1u;
    const size_t slot_off = ONESHOT_HDR + ((size_t)par * world + rank) * ONESHOT_MAX * sizeof(double);
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(peers[p] + slot_off);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world)
        st_release_sys(reinterpret_cast<unsigned int*>(peers[threadIdx.x] + 64) + par * 64 + rank, seq);
    if ((int)threadIdx.x < world) {
        const unsigned int* f = reinterpret_cast<const unsigned int*>(mine + 64) + par * 64 + threadIdx.x;
        unsigned int it = 0;
        while (ld_acquire_sys(f) != seq) {
            if (++it > (1u << 28)) __trap();          // a peer never arrived: surface an error instead of hanging the GPU
        }
    }
    __syncthreads();
    const double* slots = reinterpret_cast<const double*>(mine + ONESHOT_HDR + (size_t)par * world * ONESHOT_MAX * sizeof(double));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += slots[(size_t)r * ONESHOT_MAX + i];      // rank order: identical on every rank
        out[i] = s;
    }
}

}  // namespace rat

using namespace rat;

extern "C" size_t rat_oneshot_workspace_bytes(int world) {
    return (size_t)ONESHOT_HDR + (size_t)2 * world * ONESHOT_MAX * sizeof(double);
}

extern "C" int rat_oneshot_allreduce_f64(const void* const* peer_bufs, int rank, int world, const double* in, int n, double* out,
                                         void* stream) {
    RAT_REQUIRE(peer_bufs != nullptr && world >= 1 && world <= 64 && rank >= 0 && rank < world, "rat_oneshot_allreduce_f64: bad ranks");
    RAT_REQUIRE(n >= 1 && n <= ONESHOT_MAX, "rat_oneshot_allreduce_f64: n=%d exceeds %d", n, ONESHOT_MAX);
    k_oneshot_allreduce_f64<<<1, 256, 0, (cudaStream_t)stream>>>((unsigned char* const*)peer_bufs, rank, world, in, n, out);
    RAT_CHECK_LAUNCH("k_oneshot_allreduce_f64");
    return RAT_OK;
}
