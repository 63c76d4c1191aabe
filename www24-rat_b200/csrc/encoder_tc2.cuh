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
// Threads.  512 = 4 GROUPS of 4 warps (one warp per TMEM lane quadrant).  Every group is an independent pipeline working
// on its OWN tile (own activation tiles in shared memory, own 128 TMEM columns, own named barrier and mbarriers, its
// MMAs issued by one of its own warps), heads streamed one at a time; only the weight images are shared.  Each step of a
// tile is a short latency chain (MMA -> mbarrier -> tcgen05.ld -> thread-per-row math -> st.shared -> fence -> MMA):
// four chains in flight per SM keep the issue slots and the tensor pipe busy while any one of them waits
// (tools/tc5_timing.cu: ~28 cycles per small MMA, ~200 cycles commit -> wake-up).
#pragma once
#include "encoder_tc.cuh"

namespace rat {

constexpr int T2_THREADS = 512;
constexpr int T2_HALF_BYTES = 64 * 64 * 2;          // one block-diagonal [64 x 64] fp16 tile

// warp index as a value ptxas can prove warp-uniform (descriptor math then stays in uniform registers)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

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

// Staging of a tile's token rows, 4 threads per row (thread = (row, part)), split in two so that several rows' global loads can
// be in flight before the first is consumed:
//   t2_rows_load   : the row's 16-/8-byte units part, part+4, .. -> registers
//   t2_rows_finish : LayerNorm WITHOUT the affine part (gamma is folded into the weight image, beta rides on a column of
//                    ones at column D), fp16, chunk-major K-major tile.  Pad columns (> D) are never written
//                    (zero-initialised once); invalid rows are written as zeros.  stats (nullable): [128][2] mean, rstd.
template <bool VEC4>
struct XRegs { float v[4][VEC4 ? 4 : 2]; bool valid; };

template <bool VEC4, int SLSH>
__device__ __forceinline__ void t2_rows_load(const float* __restrict__ x, const SeqGeom& g, long long s0, long long nseq,
                                             int D, int row, int part, XRegs<VEC4>& r) {
    constexpr int U = VEC4 ? 4 : 2;
    const int slot = row >> SLSH, pos = row & ((1 << SLSH) - 1);
    const long long seq = s0 + slot;
    r.valid = pos < g.S && seq < nseq;
    const int nun = D / U;
    const float* src = x + (r.valid ? g.grow(seq, pos) : 0) * D;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = part + 4 * k;
#pragma unroll
        for (int e = 0; e < U; ++e) r.v[k][e] = 0.f;
        if (r.valid && u < nun) {
            if constexpr (VEC4) { const float4 t = *reinterpret_cast<const float4*>(src + u * 4); r.v[k][0] = t.x; r.v[k][1] = t.y; r.v[k][2] = t.z; r.v[k][3] = t.w; }
            else { const float2 t = *reinterpret_cast<const float2*>(src + u * 2); r.v[k][0] = t.x; r.v[k][1] = t.y; }
        }
    }
}

template <bool VEC4>
__device__ __forceinline__ void t2_rows_finish(const XRegs<VEC4>& r, int D, unsigned char* __restrict__ Xt, int row, int part,
                                               float* __restrict__ stats) {
    constexpr int U = VEC4 ? 4 : 2;
    const int nun = D / U;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < U; ++e) s += r.v[k][e];                  // units beyond the row hold zeros
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float invD = 1.0f / (float)D;
    const float mean = s * invD;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (part + 4 * k < nun) {
#pragma unroll
            for (int e = 0; e < U; ++e) { const float t = r.v[k][e] - mean; sq = fmaf(t, t, sq); }
        }
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    const float rstd = r.valid ? rsqrtf(fmaf(sq, invD, 1e-5f)) : 0.f;
    const float nm = -mean * rstd;
    if (stats != nullptr && part == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
    unsigned char* dst = Xt + (size_t)row * 16 + (VEC4 ? (part & 1) * 8 : (part & 3) * 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int u = part + 4 * k;
        if (u < nun) {
            // VEC4: unit u = columns [4u, 4u+4) = half (u & 1) of chunk u >> 1 ; else unit u = quarter (u & 3) of chunk u >> 2
            unsigned char* d2 = dst + (VEC4 ? (u >> 1) : (u >> 2)) * tc5::TILE_CHUNK;
            if constexpr (VEC4)
                *reinterpret_cast<uint2*>(d2) = make_uint2(pack_h2(fmaf(r.v[k][0], rstd, nm), fmaf(r.v[k][1], rstd, nm)),
                                                           pack_h2(fmaf(r.v[k][2], rstd, nm), fmaf(r.v[k][3], rstd, nm)));
            else *reinterpret_cast<uint32_t*>(d2) = pack_h2(fmaf(r.v[k][0], rstd, nm), fmaf(r.v[k][1], rstd, nm));
        }
    }
    if (part == (nun & 3))                                           // the column of ones (beta / bias row of the weight image)
        *reinterpret_cast<unsigned short*>(Xt + tc5::toff(row, D >> 3) + (D & 7) * 2) = r.valid ? (unsigned short)(TC_ONES2 & 0xffffu) : (unsigned short)0;
}

}  // namespace rat
