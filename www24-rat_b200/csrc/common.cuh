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
// Device-resident training-step counter of the dropout streams: kernels with dropout add 64 * (*ptr) to the rng stream
// they are given, so a CUDA graph of the training step replays with a fresh mask every time (the counter is advanced by
// a kernel INSIDE the graph: rat_rng_step_advance).
const unsigned int* rng_step_ptr();
// Deferred record reductions (rat_set_reduce_stream): the k_reduce_* launch that follows a backward kernel only produces
// weight gradients, which nothing reads before the optimizer, so it need not sit on the critical path between two backward
// kernels.  reduce_fork returns the stream the reduction has to be launched on (the caller's stream when no reduce stream
// is set), reduce_forked marks the reduction as the last reader of the record workspace `ws`, and reduce_ws_acquire makes
// `main` wait for that reader before a kernel overwrites `ws`.  All three are stream-capture safe (events only).
cudaStream_t reduce_fork(cudaStream_t main, const void* ws);
void reduce_forked(cudaStream_t launched_on, cudaStream_t main, const void* ws);
void reduce_ws_acquire(cudaStream_t main, const void* ws);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- programmatic dependent launch (PDL) of the RAT-block kernels.  Every register-resident kernel starts with a prologue
// that does not depend on the previous kernel of the stream (weight images in fragment order, LayerNorm / bias vectors,
// TMEM allocation, barrier init: 4-9 us per CTA).  Launched with the programmatic-stream-serialization attribute its CTAs
// may become resident as soon as every CTA of the previous kernel has passed pdl_launch_dependents() (first statement) and an
// SM has room, i.e. under the previous kernel's tail instead of after its last CTA has drained; pdl_wait() -- before the
// first read of an activation / gradient-scale slot and before any global write -- blocks until the previous grid has
// completed and its memory is visible.  Both are no-ops for a kernel launched the ordinary way.  RAT_PDL=0 turns the
// attribute off.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <class Kernel, class Args>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Args& args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}
#endif

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
// Dropout masks: ONE Philox4x32-10 call yields eight 16-bit uniforms = the keep decisions of the eight consecutive
// elements [8c, 8c+8) (counter c = e >> 3).  Element e is dropped iff its 16-bit lane < round(p * 65536).
// RAT_DROPOUT_PHILOX=1 selects Philox4x32-10 (about 100 instructions per call); the default is a keyed 32-bit integer
// hash (lowbias32, 2 multiplies + 3 xor-shifts per word) of the counter -- dropout only needs independent
// Bernoulli(p) decisions that the forward and backward kernels can both regenerate from (seed, stream, element).
#ifndef RAT_DROPOUT_PHILOX
#define RAT_DROPOUT_PHILOX 0
#endif
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t dropout_key(unsigned long long seed, uint32_t stream) {
    return lowbias32((uint32_t)seed ^ lowbias32((uint32_t)(seed >> 32) + 0x9E3779B9u * (stream + 1u)));
}
__device__ __forceinline__ uint4 dropout_bits8(unsigned long long seed, uint32_t stream, unsigned long long c) {
#if RAT_DROPOUT_PHILOX
    return philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), stream, 0u),
                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
#else
    const uint32_t key = dropout_key(seed, stream);
    const uint32_t hi = lowbias32((uint32_t)(c >> 30) ^ key);          // c*4 spans 34+ bits: fold the high part
    const uint32_t b = (uint32_t)c << 2;
    return make_uint4(lowbias32((b + 0u) ^ hi), lowbias32((b + 1u) ^ hi), lowbias32((b + 2u) ^ hi), lowbias32((b + 3u) ^ hi));
#endif
}
__device__ __forceinline__ uint32_t rng_stream_of_step(uint32_t stream, const unsigned int* __restrict__ step) {
    return step ? stream + 64u * __ldg(step) : stream;
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
// lane j (0..7) of the 128-bit Philox output
__device__ __forceinline__ uint32_t dropout_lane16(const uint4& r, int j) {
    const uint32_t w = (j >> 1) == 0 ? r.x : (j >> 1) == 1 ? r.y : (j >> 1) == 2 ? r.z : r.w;
    return (j & 1) ? (w >> 16) : (w & 0xffffu);
}
// keep-mask scale for flattened element index e: 0 (dropped) or 1/(1-p)
__device__ __forceinline__ float dropout_scale(unsigned long long seed, uint32_t stream, unsigned long long e,
                                               float p, float inv_keep) {
    const uint4 r = dropout_bits8(seed, stream, e >> 3);
    return dropout_lane16(r, (int)(e & 7)) < dropout_threshold(p) ? 0.0f : inv_keep;
}

// ---- dynamic gradient scaling of the fp16 tensor-core path.  Every kernel that writes a gradient tensor publishes
// max|dx| into a device slot (order-independent integer atomicMax on the bits of a non-negative float); the kernel
// that consumes the tensor lifts it by the power of two that puts that maximum into [4, 8) -- 2^13 of fp16 headroom for
// growth inside the kernel, 2^-17 of the maximum still a normal number -- and unscales its fp32 results.
__device__ __forceinline__ float grad_scale_from_amax(const float* amax) {
    if (amax == nullptr) return 1.0f;
    const float m = *amax;
    if (!(m > 0.f) || m > 3.0e38f) return 1.0f;
    int e;
    (void)frexpf(m, &e);                                  // m = f * 2^e, f in [0.5, 1)
    e = min(30, max(-24, 3 - e));
    return ldexpf(1.0f, e);
}
__device__ __forceinline__ void publish_amax(float* slot, float v) {      // v >= 0; called by whole warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (slot != nullptr && (threadIdx.x & 31) == 0 && v > 0.f) atomicMax(reinterpret_cast<int*>(slot), __float_as_int(v));
}

// block-wide variant (ONE atomic per block: same-address atomics serialise in L2); all threads of the block call it
__device__ __forceinline__ void publish_amax_block(float* slot, float v) {
    __shared__ float amax_w[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) amax_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0 && slot != nullptr) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) v = fmaxf(v, amax_w[w]);
        if (v > 0.f) atomicMax(reinterpret_cast<int*>(slot), __float_as_int(v));
    }
}

// ---- VW-wide (1, 2 or 4 floats) vector loads / stores
template <int VW> struct Vec;
template <> struct Vec<4> { typedef float4 T; };
template <> struct Vec<2> { typedef float2 T; };
template <> struct Vec<1> { typedef float T; };

template <int VW>
__device__ __forceinline__ void vload(const float* p, float (&v)[VW]) {
    typename Vec<VW>::T t = __ldg(reinterpret_cast<const typename Vec<VW>::T*>(p));
    const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int i = 0; i < VW; ++i) v[i] = f[i];
}
template <int VW>
__device__ __forceinline__ void vstore(float* p, const float (&v)[VW]) {
    typename Vec<VW>::T t;
    float* f = reinterpret_cast<float*>(&t);
#pragma unroll
    for (int i = 0; i < VW; ++i) f[i] = v[i];
    *reinterpret_cast<typename Vec<VW>::T*>(p) = t;
}

// ---- exact n / d for 0 <= n < 2^31 by multiply-high (round-up magic, Granlund-Montgomery); built on the host
struct FastDiv {
    uint32_t m, l;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(m, n) + n) >> l; }
};
static inline FastDiv make_fastdiv(uint32_t d) {
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    const unsigned long long m = (((1ull << l) - d) << 32) / d + 1;
    return FastDiv{(uint32_t)m, l};
}

template <int VW>
__device__ __forceinline__ void dropout_chunk(float (&val)[VW], unsigned long long idx, uint32_t key, uint32_t hk0,
                                              uint32_t thr, float inv_keep) {
    // element e = idx*VW + k uses 16-bit half (e & 1) of hash word e >> 1 (see dropout_bits8: word gw of counter
    // c = gw >> 2 is lowbias32(low32(gw) ^ lowbias32(high32(gw) ^ key)))
    if (VW >= 2) {
        const unsigned long long gw0 = idx * (VW / 2);
        const uint32_t hi32 = (uint32_t)(gw0 >> 32);
        const uint32_t hk = hi32 == 0u ? hk0 : lowbias32(hi32 ^ key);
#pragma unroll
        for (int q = 0; q < VW / 2; ++q) {
            const uint32_t w = lowbias32(((uint32_t)gw0 + (uint32_t)q) ^ hk);
            val[2 * q] *= (w & 0xffffu) < thr ? 0.0f : inv_keep;
            val[2 * q + 1] *= (w >> 16) < thr ? 0.0f : inv_keep;
        }
    } else {
        const unsigned long long gw = idx >> 1;
        const uint32_t hi32 = (uint32_t)(gw >> 32);
        const uint32_t hk = hi32 == 0u ? hk0 : lowbias32(hi32 ^ key);
        const uint32_t w = lowbias32((uint32_t)gw ^ hk);
        val[0] *= ((idx & 1) ? (w >> 16) : (w & 0xffffu)) < thr ? 0.0f : inv_keep;
    }
}

#endif

}  // namespace rat
