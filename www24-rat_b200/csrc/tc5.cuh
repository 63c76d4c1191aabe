// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; no CUTLASS dependency).
//
// Programming model used by the kernels of this library:
//   * operands are written to shared memory by ordinary threads (st.shared) in the UMMA *canonical no-swizzle*
//     layout: 8-row x 16-byte "core matrices" stored as 128 contiguous bytes; a [rows x K] K-major operand is
//       byte(r, k) = (r / 8) * SBO + (kbytes(k) / 16) * LBO + (r % 8) * 16 + kbytes(k) % 16
//     (LBO = distance between core matrices adjacent in K, SBO = distance between 8-row groups);
//   * writers execute fence.proxy.async (generic -> async proxy) and a CTA barrier; ONE thread issues
//     tcgen05.mma (D in TMEM, fp32) and tcgen05.commit -> mbarrier; consumers wait on the mbarrier, then read the
//     accumulator with tcgen05.ld (lane = accumulator row, column = accumulator column).
// Every mbarrier wait is bounded: a wait that never completes traps instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: ~seconds of polling, then trap (an error surfaces on the host instead of a hung GPU)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; ncols power of two in [32, 512]; base address (lane 0, first column) is written to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 bit x N columns: thread i of the warp receives columns [col, col+N) of lane (32*(warp%4) + i)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM (same lane / column mapping as tmem_ld8): lets a kernel park live registers in spare TMEM columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- descriptors
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B of the given format, M x N, majors
enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
__host__ __device__ constexpr uint32_t instr_desc(uint32_t fmt, int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// all previously issued MMAs of this thread complete -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- warp-uniform issue: called by ALL lanes of a converged warp with warp-uniform operands; one elected lane issues.
// Keeping the C++ control flow uniform lets ptxas hold descriptors in uniform registers (a divergent `if (tid == 0)`
// costs an ELECT / R2UR.BROADCAST waterfall of ~100 cycles per MMA).
__device__ __forceinline__ void mma_f16_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_w(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        ".reg .b64 st;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e mbarrier.arrive.shared::cta.b64 st, [%0];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
// wait with a hardware suspend-time hint: the thread sleeps inside try_wait instead of spinning through issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 16); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// descriptor of K-step `kstep` (16 bf16 = two 16-byte chunks) of a chunk-major K-major operand with `rows` rows
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, int rows, int kstep) {
    return smem_desc(saddr + (uint32_t)(kstep * 2 * rows * 16), (uint32_t)(rows * 16), 128u);
}

// ---------------------------------------------------------------------------------------------- layout helpers
// byte offset of the 16-byte chunk (row r, K-chunk kc) of a K-major operand with ROWS rows, stored CHUNK-major:
//   [kc][row][16 B]   =>   8 consecutive rows of one chunk = one 128-byte core matrix,
//                          SBO (next 8-row group) = 128, LBO (next K chunk) = ROWS * 16.
// The offset is linear in the row: a warp task adds its base row once and keeps per-lane constants.
__host__ __device__ constexpr uint32_t kmajor_off(int r, int kc, int ROWS) { return (uint32_t)((kc * ROWS + r) * 16); }
// activation tiles always have 128 rows
__host__ __device__ constexpr uint32_t toff(int r, int kc) { return (uint32_t)((kc * 128 + r) * 16); }
constexpr uint32_t TILE_CHUNK = 128 * 16;      // byte stride between K chunks of a 128-row tile

}  // namespace tc5
