// Second-generation tcgen05 attention kernels (encoder_tc2_*.cu): EVERY product of the attention sub-block runs on
// the 5th-generation tensor cores, including the per-sequence score / probability products that the first generation
// ran as warp-level mma.sync tasks (180 warp-instructions per token forward, 510 backward).
//
// Geometry.  A tile is 128 token rows = 8 row groups of 16; a sequence of S <= 16 tokens occupies a slot of SL = 16
// (S > 8) or SL = 8 rows (two sequences per row group), pad rows are zero.  Scores of one head are TWO M=64, N=64, K=DHP
// products (one per 64-row half): an M=64 cta_group::1 accumulator puts row m into TMEM lane (m%16) + 32*(m/16), a
// lane offset of 16 puts the second half beside it (tools/tc5_probe.cu), so warp q of a 4-warp group finds row
// group q of half 0 in its lanes 0-15 and row group q of half 1 in its lanes 16-31, and ONE tcgen05.ld.32x32b.x16 at
// column 16q hands every thread the 16 scores of its own row against its own row group: thread = token row, the softmax is
// thread-local (no shuffles, no fragment layouts).  Probabilities go back to shared memory as a block-diagonal
// [64 x 64] fp16 tile per half (off-diagonal blocks are zero and never rewritten) and P.V is a tcgen05 product again
// (A = that tile, K-major; B = the v tile read MN-major).  In the backward the same block-diagonal tiles are read
// MN-major to get P^T and dS^T for free.
//
// Threads.  512 = 4 head GROUPS of 4 warps (one warp per TMEM lane quadrant).  A group owns heads {g, g+4, ..} of a
// head chunk and runs them as an independent pipeline (own TMEM column region, own named barrier, own mbarriers, its
// MMAs issued by its own elected thread), so while one group waits for the tensor pipe the other three compute.
#pragma once
#include "encoder_tc.cuh"

namespace rat {

constexpr int T2_THREADS = 512;
constexpr int T2_HALF_BYTES = 64 * 64 * 2;          // one block-diagonal [64 x 64] fp16 tile

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, float (&v)[16]) { tc5::tmem_ld16(taddr, v); }

// byte offset of row r, 16-byte chunk kc of a block-diagonal half tile (64 rows)
__device__ __forceinline__ uint32_t poff(int r, int kc) { return (uint32_t)((kc * 64 + r) * 16); }

// x[j] = score of own-slot key j, taken from the 16 scores of the row group (sb = sub-slot of this row when SL == 8)
template <int SL>
__device__ __forceinline__ float slot_pick(const float (&v)[16], int sb, int j) {
    if (SL == 16) return v[j];
    return sb ? v[8 + j] : v[j];
}

// Stage the rows [32*grp, 32*grp + 32) of a tile: 4 threads per row, LayerNorm (optional affine) in fp32, fp16 store
// into the chunk-major K-major tile.  gt = thread index inside the group (0..127).  Pad columns [D, Kp) are never
// written (zero-initialised once); invalid rows are written as zeros.  stats (nullable): [128][2] mean, rstd.
template <bool VEC4, int SLSH>
__device__ __forceinline__ void t2_stage_rows(const float* __restrict__ x, const SeqGeom& g, long long s0, long long nseq,
                                              int D, const float* __restrict__ lnw_s, const float* __restrict__ lnb_s,
                                              unsigned char* __restrict__ Xt, int grp, int gt, float* __restrict__ stats) {
    constexpr int U = VEC4 ? 4 : 2;
    const int row = 32 * grp + (gt >> 2), part = gt & 3;
    const int slot = row >> SLSH, pos = row & ((1 << SLSH) - 1);
    const long long seq = s0 + slot;
    const bool valid = pos < g.S && seq < nseq;
    const int nun = D / U;
    const float* src = x + (valid ? g.grow(seq, pos) : 0) * D;
    float v[4][U];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = part + 4 * k;
#pragma unroll
        for (int e = 0; e < U; ++e) v[k][e] = 0.f;
        if (valid && u < nun) {
            if constexpr (VEC4) { const float4 t = *reinterpret_cast<const float4*>(src + u * 4); v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w; }
            else { const float2 t = *reinterpret_cast<const float2*>(src + u * 2); v[k][0] = t.x; v[k][1] = t.y; }
        }
#pragma unroll
        for (int e = 0; e < U; ++e) s += v[k][e];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float mean = s / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = part + 4 * k;
        if (u < nun) {
#pragma unroll
            for (int e = 0; e < U; ++e) { const float t = v[k][e] - mean; sq = fmaf(t, t, sq); }
        }
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    const float rstd = 1.0f / sqrtf(sq / (float)D + 1e-5f);
    if (stats != nullptr && part == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = part + 4 * k;
        if (u < nun) {
            const int c = u * U;
            float y[U];
#pragma unroll
            for (int e = 0; e < U; ++e) y[e] = valid ? (v[k][e] - mean) * rstd * lnw_s[c + e] + lnb_s[c + e] : 0.f;
            unsigned char* dst = Xt + tc5::toff(row, c >> 3) + (c & 7) * 2;
            if constexpr (VEC4) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_h2(y[0], y[1]), pack_h2(y[2], y[3]));
            else *reinterpret_cast<uint32_t*>(dst) = pack_h2(y[0], y[1]);
        }
    }
}

}  // namespace rat
