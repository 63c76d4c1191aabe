// Register-resident attention (third generation of the RAT sub-block kernels, fp16 tensor-core mode).
//
// Why: at RAT's shapes (sequence <= 16 tokens, head width 10, D = 40) every product of the sub-block is a 16 x 16 x (16..48)
// GEMM.  The tile-based tcgen05 kernels spend their time moving those tiny matrices between TMEM, registers and shared memory
// in CTA-wide lock step (profiles/r01_final_backward_kernels_ncu_full.txt: 45 % of the instructions are integer / address
// work, issue slots 39 % busy, tensor pipe 12 %).  Here ONE WARP owns one 16-row fragment tile (a sequence, or two sequences
// of <= 8 tokens) from the global load to the global store:
//   * LayerNorm in registers, written straight into mma.sync A fragments (the thread that owns row g / g+8 and columns
//     2t, 2t+1, 2t+8, 2t+9 of a k-step loads exactly those floats);
//   * every accumulator fragment (C layout) IS the operand fragment of the next product after a cvt.f16x2 pack:
//     q -> A of S = q k^T,  k -> B of S,  P -> A of O = P v,  O -> A of the out-projection;  v is produced TRANSPOSED
//     (v^T = Wv LN(x)^T, the same registers with the operand roles swapped) so that it is the B operand of P v;
//   * weights live in shared memory in FRAGMENT ORDER (one conflict-free LDS.128 feeds two HMMAs).
// No activation ever touches shared memory in the forward, there is no CTA barrier after the weight images are built, and
// warps drift apart freely, so the HMMA pipe (8.1 cycles per m16n8k16 per sub-partition, tools/hmma_probe.cu) stays fed.
#pragma once
#include "encoder_tc.cuh"

namespace rat {

// mma.sync / movmatrix WITHOUT `volatile`: they are pure functions of their operands, and the compiler keeps volatile asm
// statements in program order -- which would pin the HMMAs of two unrolled heads one after the other instead of letting
// the scheduler interleave the independent chains (HMMA latency ~100 cycles on B200).
__device__ __forceinline__ void rr_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void rr_mma_hh(uint32_t (&c)[2], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
        : "+r"(c[0]), "+r"(c[1])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t rr_movm(uint32_t a) {
    uint32_t d;
    asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

// One fragment-ordered image entry = uint4 per lane for a PAIR of 8-wide n-tiles (16 "n" rows n0..n0+15) and one 16-wide
// k-step (columns k0..k0+15) of a matrix elem(n, k):
//   .x = {e(n0+g,   k0+2t), e(n0+g,   k0+2t+1)}   .y = {e(n0+g,   k0+2t+8), e(n0+g,   k0+2t+9)}
//   .z = {e(n0+8+g, k0+2t), e(n0+8+g, k0+2t+1)}   .w = {e(n0+8+g, k0+2t+8), e(n0+8+g, k0+2t+9)}
// As the B operand of D = A . E^T :  n-tile 0 = (.x, .y), n-tile 1 = (.z, .w).
// As the A operand (rows n0..n0+15) :  a = {.x, .z, .y, .w}.
template <class F>
__device__ __forceinline__ uint4 frag_pair_entry(int lane, int n0, int k0, F elem) {
    const int g = lane >> 2, t = lane & 3;
    uint4 r;
    r.x = pack_h2(elem(n0 + g, k0 + 2 * t), elem(n0 + g, k0 + 2 * t + 1));
    r.y = pack_h2(elem(n0 + g, k0 + 2 * t + 8), elem(n0 + g, k0 + 2 * t + 9));
    r.z = pack_h2(elem(n0 + 8 + g, k0 + 2 * t), elem(n0 + 8 + g, k0 + 2 * t + 1));
    r.w = pack_h2(elem(n0 + 8 + g, k0 + 2 * t + 8), elem(n0 + 8 + g, k0 + 2 * t + 9));
    return r;
}

// Per-lane constants of a warp task: fragment row i (0..15) -> (sequence of the task, position), -1 = no such token.
struct RRLane {
    int lo_sq, lo_pos, hi_sq, hi_pos;   // rows g and g + 8 (pos = -1: absent)
    float madd[2][4];                   // additive score mask of this thread's 8 score elements (0 / -inf)
    bool packed;                        // two sequences of <= 8 tokens per task
};
__device__ __forceinline__ RRLane make_rr_lane(int S, int lane) {
    RRLane c;
    c.packed = S <= 8;
    const int g = lane >> 2, t = lane & 3;
    const bool packed = c.packed;
    auto exists = [&](int i) { return (packed ? (i & 7) : i) < S; };
    c.lo_sq = 0; c.lo_pos = exists(g) ? g : -1;
    c.hi_sq = packed ? 1 : 0; c.hi_pos = exists(g + 8) ? (packed ? g : g + 8) : -1;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = (e < 2) ? g : g + 8, j = 8 * nt + 2 * t + (e & 1);
            const bool ok = exists(j) && (!packed || ((i >> 3) == (j >> 3)));
            c.madd[nt][e] = ok ? 0.f : -INFINITY;
        }
    asm volatile("" : "+r"(c.lo_pos), "+r"(c.hi_pos), "+r"(c.hi_sq));
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) asm volatile("" : "+f"(c.madd[nt][e]));
    return c;
}

// Row softmax of a 16 x 16 score fragment (log2 domain: the score scale * log2(e) is folded into Wq), two rows per thread
// (e = 0,1: row g ; e = 2,3: row g + 8), quad shuffles for the row reductions.  vlo / vhi: the row exists; rows that do not
// exist get P = 0.  On return sc holds P.
__device__ __forceinline__ void rr_softmax(float (&sc)[2][4], const RRLane& cl, bool vlo, bool vhi) {
    float mlo = -INFINITY, mhi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[nt][e] += cl.madd[nt][e];
            if (e < 2) mlo = fmaxf(mlo, sc[nt][e]); else mhi = fmaxf(mhi, sc[nt][e]);
        }
    mlo = qmax(mlo); mhi = qmax(mhi);
    if (mlo == -INFINITY) mlo = 0.f;
    if (mhi == -INFINITY) mhi = 0.f;
    float llo = 0.f, lhi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[nt][e] = ex2f(sc[nt][e] - ((e < 2) ? mlo : mhi));
            if (e < 2) llo += sc[nt][e]; else lhi += sc[nt][e];
        }
    llo = qsum(llo); lhi = qsum(lhi);
    const float ilo = vlo ? rcp_fast(llo) : 0.f, ihi = vhi ? rcp_fast(lhi) : 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[nt][e] *= (e < 2) ? ilo : ihi;
}

// accumulator fragment [16 x 16] (two n-tiles) -> A operand fragment of the next product
__device__ __forceinline__ void c_to_a(const float (&c)[2][4], uint32_t (&a)[4]) {
    a[0] = pack_h2(c[0][0], c[0][1]); a[1] = pack_h2(c[0][2], c[0][3]);
    a[2] = pack_h2(c[1][0], c[1][1]); a[3] = pack_h2(c[1][2], c[1][3]);
}

// One projection pair-of-n-tiles step  c[16 x 16] += a[16 x 16] . (f.x, f.y | f.z, f.w)  with fp16 (F16P) or fp32 accumulators.
// ProjAcc<true> keeps packed fp16 accumulators, ProjAcc<false> fp32 ones; frag() returns the 8x8 blocks
// {rows g cols 0-7, rows g+8 cols 0-7, rows g cols 8-15, rows g+8 cols 8-15} as packed operand registers.
template <bool F16P> struct ProjAcc;
template <> struct ProjAcc<true> {
    uint32_t c[2][2] = {};
    __device__ __forceinline__ void mma(int nt, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) { rr_mma_hh(c[nt], a, b0, b1); }
    __device__ __forceinline__ void frag(uint32_t (&f)[4]) const { f[0] = c[0][0]; f[1] = c[0][1]; f[2] = c[1][0]; f[3] = c[1][1]; }
};
template <> struct ProjAcc<false> {
    float c[2][4] = {};
    __device__ __forceinline__ void mma(int nt, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) { rr_mma(c[nt], a, b0, b1); }
    __device__ __forceinline__ void frag(uint32_t (&f)[4]) const { c_to_a(c, f); }
};

// Loads the two token rows of this thread (columns 8 nt + 2t, 8 nt + 2t + 1 for nt < NTO, zero beyond D or for absent rows)
template <int NTO>
__device__ __forceinline__ void rr_load_rows(const float* __restrict__ plo, const float* __restrict__ phi, bool vlo, bool vhi,
                                             int D, int t, float2 (&xl)[NTO], float2 (&xh)[NTO]) {
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) {
        const int c = 8 * nt + 2 * t;
        xl[nt] = make_float2(0.f, 0.f); xh[nt] = make_float2(0.f, 0.f);
        if (vlo && c < D) xl[nt] = *reinterpret_cast<const float2*>(plo + c);
        if (vhi && c < D) xh[nt] = *reinterpret_cast<const float2*>(phi + c);
    }
}

// LayerNorm statistics of one row spread over the 4 threads of a quad (values beyond D are zero on entry).
// invD = 1 / D.  rsqrt.approx (2 ulp) is far below the fp16 rounding of the normalised row that follows.
template <int NTO>
__device__ __forceinline__ void rr_row_stats(const float2 (&x)[NTO], int D, float invD, int t, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) s += x[nt].x + x[nt].y;
    s = qsum(s);
    mean = s * invD;
    float sq = 0.f;
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt) {
        const bool ok = 8 * nt + 2 * t < D;
        const float a = ok ? x[nt].x - mean : 0.f, b = ok ? x[nt].y - mean : 0.f;
        sq = fmaf(a, a, sq); sq = fmaf(b, b, sq);
    }
    sq = qsum(sq);
    rstd = rsqrtf(fmaf(sq, invD, 1e-5f));
}

// ---- bulk asynchronous copies (TMA unit, 1-D): global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc5::smem_u32(bar)), "r"(bytes) : "memory");
}
// size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(tc5::smem_u32(bar))
                 : "memory");
}

}  // namespace rat
