// K6: deterministic sorted segment-reduce of the embedding gradient.
//
// Replaces ATen embedding_dense_backward reached from loss.backward() (base_model.py:223) for every per-field
// nn.Embedding of EmbeddingDictLayer (embedding.py:79-100) and LR_Layer (shallow.py:31), plus the 3-row label
// table (RAT_m2.py:64) and the backward of nn.Dropout(emb_dropout) (RAT_m2.py:135).  Two phases:
//   PLAN (depends on the ids only -> the engine runs it on a side stream under the forward/backward kernels)
//     k_build_keys        key[i] = table row of occurrence i=(b,t,l) (padding ids -> sentinel); the label token of
//                         row (b,t) is pseudo column L with key V + label;  val[i] = packed (bt, is-target, column)
//     k_radix_{hist,scan,scatter} x passes   stable LSD radix sort (8-bit digits) of (key, val)
//     k_plan_runs         runs of equal keys that cross 32-position chunk boundaries: for every chunk holding the
//                         HEAD of such a run, find the run's last chunk and emit fix-up work items (<= 32 partial
//                         records each): short runs get one FINAL item, hot ids get PARTIAL items (second-level
//                         records) plus one LONG item that combines them
//   REDUCE (critical path, 2 launches)
//     k_segment_scan      one warp per 32 sorted positions, LANE = POSITION: every lane loads its occurrence's
//                         gradient row (all vectors independent -> deep memory-level parallelism), applies the
//                         dropout mask, and a shuffle segmented scan sums each run of equal keys in a fixed tree
//                         order; the run's last lane stores the row once.  Runs crossing chunk boundaries leave
//                         per-chunk partial records (carryL: run head, carryF: continuation).
//     k_fixup_items       one warp per work item: <= 32 records (lane = record, fixed butterfly) -> final row or
//                         2nd-level record; the warp finishing a hot id's last item combines its 2nd-level records
// For a given input every sum has one fixed association => bitwise run-to-run deterministic.
// The gradient of occurrence (b,t,l) is mask*dBlock[b,t,1+field(l),:] (+ dXemb[b,field(l),:] for the target row
// t=0, the DNN path) and, for the LR table, dlogit[b] for t=0.
#include <algorithm>
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace rat {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block

// val = (bt << 6) | (t == 0 ? 32 : 0) | l : the segment reduce decodes an occurrence with shifts only (L <= 31,
// column code 31 = label token); because (bt, l) is lexicographic in the occurrence index the stable sort still
// yields the canonical order.  n = B*T*(L+1) occurrences.
constexpr unsigned int LABEL_COL = 31u;
__global__ void k_build_keys(const int* __restrict__ ids, const int* __restrict__ labels, const int* __restrict__ col_off,
                             const int* __restrict__ col_pad, const int* __restrict__ col_vocab, long long n, int L, int T,
                             unsigned int V, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                             unsigned int* __restrict__ totals, int ntotals) {
    const unsigned int sentinel = V + 3u;
    const int L1 = L + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int bt = (unsigned int)(i / L1);
        const int l = (int)(i - (long long)bt * L1);
        unsigned int k = sentinel;
        const unsigned int tflag = (bt % (unsigned int)T) == 0u ? 32u : 0u;
        if (l < L) {
            const int id = ids[(long long)bt * L + l];
            if (id >= 0 && id < col_vocab[l] && id != col_pad[l]) k = (unsigned int)(col_off[l] + id);
            vals[i] = (bt << 6) | tflag | (unsigned int)l;
        } else {
            const int lab = labels[bt];
            if (lab >= 0 && lab <= 2) k = V + (unsigned int)lab;
            vals[i] = (bt << 6) | LABEL_COL;                 // never the "target row" path: no dXemb / dlogit term
        }
        keys[i] = k;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntotals; i += gridDim.x * blockDim.x) totals[i] = 0u;
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const unsigned int* __restrict__ keys, long long n, int shift,
                                                           unsigned int* __restrict__ hist, int nblk,
                                                           unsigned int* __restrict__ totals) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long p = base + r * RS_THREADS + threadIdx.x;
        if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
    if (totals && h[threadIdx.x]) atomicAdd(&totals[threadIdx.x], h[threadIdx.x]);     // integer: order independent
}

// exclusive scan of the digit-major histogram hist[digit][block]: block d of the grid scans row d and adds the
// number of keys with a smaller digit (from the per-digit totals).  256 blocks instead of one.
__global__ void __launch_bounds__(256) k_scan_digits(unsigned int* __restrict__ hist, int nblk,
                                                     const unsigned int* __restrict__ totals) {
    __shared__ unsigned int red[8];
    __shared__ unsigned int carry_s;
    const int d = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int v = (int)threadIdx.x < d ? totals[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int b = 0; for (int i = 0; i < 8; ++i) b += red[i]; carry_s = b; }
    __syncthreads();
    unsigned int* row = hist + (size_t)d * nblk;
    for (int t0 = 0; t0 < nblk; t0 += 256) {
        const int i = t0 + threadIdx.x;
        const unsigned int x = i < nblk ? row[i] : 0u;
        unsigned int inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();                       // red / carry_s of the previous tile fully consumed
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        unsigned int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += red[w];
        const unsigned int base = carry_s;
        if (i < nblk) row[i] = base + wbase + inc - x;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = base + wbase + inc;
    }
}

// exclusive scan of `n` counters in place (single block of 1024 threads)
__global__ void __launch_bounds__(1024) k_scan_exclusive(unsigned int* __restrict__ data, int n) {
    __shared__ unsigned int warp_tot[32];
    const int per = (n + 1023) / 1024;
    const int beg = threadIdx.x * per, end = min(n, beg + per);
    unsigned int s = 0;
    for (int i = beg; i < end; ++i) s += data[i];
    // block exclusive scan of s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_tot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        warp_tot[lane] = winc - w;
    }
    __syncthreads();
    unsigned int run = warp_tot[warp] + inc - s;
    for (int i = beg; i < end; ++i) { unsigned int v = data[i]; data[i] = run; run += v; }
}

// stable scatter: each warp owns a contiguous 256-key slice of the block tile and ranks it with match_any
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const unsigned int* __restrict__ keys_in,
                                                              const unsigned int* __restrict__ vals_in,
                                                              unsigned int* __restrict__ keys_out,
                                                              unsigned int* __restrict__ vals_out, long long n, int shift,
                                                              const unsigned int* __restrict__ hist, int nblk) {
    __shared__ unsigned int cnt[RS_THREADS / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const long long wbase = (long long)blockIdx.x * RS_TILE + (long long)warp * (32 * RS_ITEMS);
    unsigned int key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long p = wbase + r * 32 + lane;
        const bool valid = p < n;
        key[r] = valid ? keys_in[p] : 0xffffffffu;
        val[r] = valid ? vals_in[p] : 0u;
        const unsigned int amask = __ballot_sync(0xffffffffu, valid);
        rank[r] = 0;
        if (valid) {
            const unsigned int d = (key[r] >> shift) & 255u;
            const unsigned int peers = __match_any_sync(amask, d);
            const unsigned int before = __popc(peers & ((1u << lane) - 1u));
            rank[r] = cnt[warp][d] + before;
            __syncwarp(amask);
            if (before == 0) cnt[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps + global base
    {
        const int d = threadIdx.x;
        unsigned int run = hist[(size_t)d * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { unsigned int c = cnt[w][d]; cnt[w][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long p = wbase + r * 32 + lane;
        if (p < n) {
            const unsigned int d = (key[r] >> shift) & 255u;
            const unsigned int dst = cnt[warp][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// ---- segment reduce ------------------------------------------------------------------------------------------
struct SegArgs {
    const unsigned int* keys; const unsigned int* vals; long long n; unsigned int sentinel, V;
    const float* dblock;     // [B,T,N,D]
    const float* dxemb;      // [B,F*D] or nullptr
    const float* dlogit;     // [B] or nullptr
    const int* col_field;    // [L]
    float* g_emb;            // [V,D]
    float* g_lr;             // [V] or nullptr
    float* g_label;          // [3,D] or nullptr
    float* carryF; float* carryL;   // [nchunks][DS]  (DS = round_up(D+1, 4); element D = LR scalar)
    float* carry2;           // [..][DS] second-level records of hot ids
    unsigned int* counters;  // [0] = #items, [1] = #long items, [2] = #second-level records
    uint4* items;            // {first record chunk, count, kind (0 FINAL: carryL[first] + carryF[first+1..first+count],
                             //  1+i PARTIAL of hot id i: carryF[first..first+count-1]), destination (key | carry2 slot)}
    uint4* longs;            // {head chunk, first carry2 slot, #slots, key}
    unsigned int* done;      // [#hot ids] finished PARTIAL items (the warp finishing the last one combines them)
    int T, L, N, D, F, DS;
    FastDiv divT;
    float drop_p; unsigned long long seed; unsigned int stream;
    const unsigned int* step;    // device step counter of the dropout streams (rng_step_ptr)
};

__device__ __forceinline__ float* seg_final_row(const SegArgs& a, unsigned int key) {
    if (key < a.V) return a.g_emb + (size_t)key * a.D;
    return a.g_label ? a.g_label + (size_t)(key - a.V) * a.D : nullptr;
}

constexpr int SEG_G = 5;      // vectors of a gradient row scanned together (D = 40: two groups of five float4)

template <int VW>
__global__ void __launch_bounds__(256) k_segment_scan(SegArgs a) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long p0 = chunk * 32;
    if (p0 >= a.n) return;
    const long long p = p0 + lane;
    const unsigned int key = p < a.n ? a.keys[p] : a.sentinel;
    const unsigned int src = p < a.n ? a.vals[p] : 0u;
    const unsigned int prev_g = p0 > 0 ? a.keys[p0 - 1] : a.sentinel;             // sentinel == "different"
    const unsigned int next_g = p0 + 32 < a.n ? a.keys[p0 + 32] : a.sentinel;
    const unsigned int prevk = __shfl_up_sync(FULL, key, 1);
    const bool head = lane == 0 || key != prevk;
    const unsigned int hm = __ballot_sync(FULL, head);
    const int start = 31 - __clz(hm & (FULL >> (31 - lane)));                    // first lane of my run
    const bool tail = lane == 31 || ((hm >> (lane + 1)) & 1u);
    const bool valid = key != a.sentinel;
    // ---- decode the occurrence
    const unsigned int l = src & 31u, bt = src >> 6;
    const bool t0 = (src & 32u) != 0u, is_lab = l == LABEL_COL;
    const int f = is_lab ? 0 : a.col_field[l];
    const size_t e0 = ((size_t)bt * a.N + (is_lab ? 0 : 1 + f)) * a.D;           // element index inside the block
    const float* grow = a.dblock + e0;
    const float* xrow = nullptr;
    float lr = 0.f;
    if (valid && t0) {
        const unsigned int b = a.divT.div(bt);
        if (a.dxemb) xrow = a.dxemb + ((size_t)b * a.F + f) * a.D;
        if (a.dlogit) lr = a.dlogit[b];
    }
    // ---- where the run that ends at this lane goes
    float* dst = nullptr;
    float* dst_lr = nullptr;
    if (tail && valid) {
        const bool cont = start == 0 && p0 > 0 && prev_g == key;
        const bool fwd = lane == 31 && next_g == key;
        if (cont) { dst = a.carryF + chunk * a.DS; dst_lr = dst + a.D; }
        else if (fwd) { dst = a.carryL + chunk * a.DS; dst_lr = dst + a.D; }
        else { dst = seg_final_row(a, key); dst_lr = (a.g_lr && key < a.V) ? a.g_lr + key : nullptr; }
    }
    bool okk[5];
    unsigned int rounds = 0;                                 // scan rounds some lane of the warp needs (warp-uniform)
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        okk[s] = lane - (1 << s) >= start;
        if (__any_sync(FULL, okk[s])) rounds |= 1u << s;
    }
    const bool drop = a.drop_p > 0.f;
    const float inv_keep = drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    const uint32_t thr = dropout_threshold(a.drop_p);
    const uint32_t dkey = dropout_key(a.seed, rng_stream_of_step(a.stream, a.step)), hk0 = lowbias32(dkey);
    const int DV = a.D / VW;
    const unsigned long long idx0 = e0 / VW;
    for (int v0 = 0; v0 < DV; v0 += SEG_G) {
        float v[SEG_G][VW], x[SEG_G][VW];
#pragma unroll
        for (int g = 0; g < SEG_G; ++g) {                    // every load of the group is issued before any is used
#pragma unroll
            for (int k = 0; k < VW; ++k) { v[g][k] = 0.f; x[g][k] = 0.f; }
            if (valid && v0 + g < DV) vload<VW>(grow + (v0 + g) * VW, v[g]);
            if (xrow != nullptr && v0 + g < DV) vload<VW>(xrow + (v0 + g) * VW, x[g]);
        }
#pragma unroll
        for (int g = 0; g < SEG_G; ++g) {
            if (drop && valid && v0 + g < DV) dropout_chunk<VW>(v[g], idx0 + (v0 + g), dkey, hk0, thr, inv_keep);
#pragma unroll
            for (int k = 0; k < VW; ++k) v[g][k] += x[g][k];
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (!((rounds >> s) & 1u)) continue;
#pragma unroll
            for (int g = 0; g < SEG_G; ++g)
#pragma unroll
                for (int k = 0; k < VW; ++k) {
                    const float t = __shfl_up_sync(FULL, v[g][k], 1 << s);
                    if (okk[s]) v[g][k] += t;
                }
        }
        if (dst != nullptr) {
#pragma unroll
            for (int g = 0; g < SEG_G; ++g)
                if (v0 + g < DV) vstore<VW>(dst + (v0 + g) * VW, v[g]);
        }
    }
    if (a.dlogit) {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (!((rounds >> s) & 1u)) continue;
            const float t = __shfl_up_sync(FULL, lr, 1 << s);
            if (okk[s]) lr += t;
        }
        if (dst_lr != nullptr) *dst_lr = lr;
    }
}

// ---- runs that cross chunk boundaries --------------------------------------------------------------------------
// Chunk c holds the HEAD of a forward-spanning run when its last key continues into chunk c+1 and the run did not
// already enter c from c-1.  The run's last chunk `end` is the first cc > c that is the final chunk or does not end
// inside the run; the result row is carryL[c] + carryF[c+1] + ... + carryF[end], added in chunk order.
struct PlanRunsArgs {
    const unsigned int* keys; long long n; unsigned int sentinel;
    unsigned int* counters; uint4* items; uint4* longs; unsigned int* done;
};
__device__ __forceinline__ bool seg_chunk_stops(const unsigned int* __restrict__ keys, long long n, long long cc,
                                                unsigned int key) {
    const long long cl = cc * 32 + 31;
    return (cl >= n - 1) || keys[cl] != key || keys[cl + 1] != key;
}
__global__ void __launch_bounds__(256) k_plan_runs(PlanRunsArgs a) {
    const int lane = threadIdx.x & 31;
    const long long nchunks = (a.n + 31) / 32;
    const long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nchunks - 1) return;                                       // the last chunk cannot span forward
    const long long last = c * 32 + 31;
    const unsigned int key = a.keys[last];
    if (key == a.sentinel || a.keys[last + 1] != key) return;           // last run does not span forward
    if (a.keys[c * 32] == key && c > 0 && a.keys[c * 32 - 1] == key) return;   // continues from before: not the head
    long long end = -1;
    for (long long base = c + 1; base < nchunks && end < 0; base += 32) {
        const long long cc = base + lane;
        const bool stop = cc < nchunks && seg_chunk_stops(a.keys, a.n, cc, key);
        const unsigned int sm = __ballot_sync(0xffffffffu, stop);
        if (sm) end = base + __ffs(sm) - 1;
    }
    const unsigned int m = (unsigned int)(end - c);                     // carryF records c+1 .. end
    if (m <= 31u) {                                                     // head partial + m records fit one warp
        if (lane == 0) a.items[atomicAdd(&a.counters[0], 1u)] = make_uint4((unsigned int)c, m, 0u, key);
        return;
    }
    const unsigned int nseg = (m + 31u) / 32u;
    unsigned int slot0 = 0, item0 = 0, li = 0;
    if (lane == 0) {
        slot0 = atomicAdd(&a.counters[2], nseg);                        // slot numbers do not affect the arithmetic
        item0 = atomicAdd(&a.counters[0], nseg);
        li = atomicAdd(&a.counters[1], 1u);
        a.longs[li] = make_uint4((unsigned int)c, slot0, nseg, key);
        a.done[li] = 0u;
    }
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    li = __shfl_sync(0xffffffffu, li, 0);
    for (unsigned int j = lane; j < nseg; j += 32)
        a.items[item0 + j] = make_uint4((unsigned int)c + 1u + 32u * j, min(32u, m - 32u * j), 1u + li, slot0 + j);
}

// LANE = RECORD: lane j loads record j of the item (all loads independent: one memory round trip), a fixed xor
// butterfly adds the <= 32 records, lane q of every group of 4 vectors stores vector q.  DS % 4 == 0, 16-byte records.
// The warp that finishes the LAST partial item of a hot id combines that id's second-level records (slot order, 32 at
// a time, same butterfly) after the head partial -- whichever warp it is, the association is the same.
// FX_G groups of 4 float4 are handled together (3: 12 vectors = 48 floats >= DS of D <= 44 in one pass)
template <int FX_G>
__device__ __forceinline__ void seg_store_vecs(const SegArgs& a, const float4 (&v)[FX_G][4], int vb, int NV, int lane,
                                               float* row, float* dst_lr) {
#pragma unroll
    for (int g = 0; g < FX_G; ++g)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int vec = vb + 4 * g + q;
            if (lane != 4 * g + q || vec >= NV) continue;              // lane 4g+q stores vector 4g+q of the pass
            const float f[4] = {v[g][q].x, v[g][q].y, v[g][q].z, v[g][q].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = 4 * vec + k;
                if (e < a.D) { if (row) row[e] = f[k]; }
                else if (e == a.D && dst_lr) *dst_lr = f[k];
            }
        }
}
template <int FX_G>
__device__ __forceinline__ void seg_butterfly(float4 (&v)[FX_G][4]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int g = 0; g < FX_G; ++g)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[g][q].x += __shfl_xor_sync(0xffffffffu, v[g][q].x, o); v[g][q].y += __shfl_xor_sync(0xffffffffu, v[g][q].y, o);
                v[g][q].z += __shfl_xor_sync(0xffffffffu, v[g][q].z, o); v[g][q].w += __shfl_xor_sync(0xffffffffu, v[g][q].w, o);
            }
}
// v[g][q] = vector vb + 4g + q of record `rec` (zeros when rec == nullptr or past the record); CG = bypass L1
template <bool CG, int FX_G>
__device__ __forceinline__ void seg_load_vecs(const float* rec, int vb, int NV, float4 (&v)[FX_G][4]) {
#pragma unroll
    for (int g = 0; g < FX_G; ++g)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int vec = vb + 4 * g + q;
            v[g][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rec != nullptr && vec < NV) {
                const float4* p = reinterpret_cast<const float4*>(rec + 4 * vec);
                v[g][q] = CG ? __ldcg(p) : *p;
            }
        }
}

__global__ void __launch_bounds__(256, 3) k_fixup_items(SegArgs a) {
    const int lane = threadIdx.x & 31;
    const unsigned int nitems = a.counters[0];
    const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
    const int NV = a.DS >> 2;
    for (unsigned int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nitems; i += nwarps) {
        const uint4 it = a.items[i];
        const long long c = it.x;
        const int m = (int)it.y;
        const bool fin = it.z == 0u;
        // FINAL: records carryL[c], carryF[c+1 .. c+m] (m <= 31);  PARTIAL: carryF[c .. c+m-1] (m <= 32)
        const float* rec = nullptr;
        if (fin) { if (lane <= m) rec = (lane == 0 ? a.carryL : a.carryF) + (c + lane) * a.DS; }
        else if (lane < m) rec = a.carryF + (c + lane) * a.DS;
        float* row = nullptr;
        float* dst_lr = nullptr;
        if (fin) { row = seg_final_row(a, it.w); dst_lr = (a.g_lr && it.w < a.V) ? a.g_lr + it.w : nullptr; }
        else { row = a.carry2 + (size_t)it.w * a.DS; dst_lr = row + a.D; }
        for (int vb = 0; vb < NV; vb += 12) {                            // one memory round trip for D <= 44
            float4 v[3][4];
            seg_load_vecs<false, 3>(rec, vb, NV, v);
            seg_butterfly<3>(v);
            seg_store_vecs<3>(a, v, vb, NV, lane, row, dst_lr);
        }
        if (fin) continue;
        // ---- hot id: am I the last partial item of it?
        const unsigned int li = it.z - 1u;
        const uint4 lg = a.longs[li];                                    // {head chunk, first slot, #slots, key}
        unsigned int last = 0;
        __threadfence();                                                 // this warp's second-level record is visible
        __syncwarp();
        if (lane == 0) {
            last = atomicAdd(&a.done[li], 1u) == lg.z - 1u ? 1u : 0u;
            if (last) a.done[li] = 0u;                                   // the plan stays reusable
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) continue;
        __threadfence();
        float* frow = seg_final_row(a, lg.w);
        float* flr = (a.g_lr && lg.w < a.V) ? a.g_lr + lg.w : nullptr;
        for (int vb = 0; vb < NV; vb += 4) {                             // rare path: 4 vectors at a time (registers)
            float4 tot[1][4];
            seg_load_vecs<true, 1>(a.carryL + (size_t)lg.x * a.DS, vb, NV, tot);     // head partial first
            for (unsigned int s0 = 0; s0 < lg.z; s0 += 32) {
                float4 v[1][4];
                seg_load_vecs<true, 1>(s0 + lane < lg.z ? a.carry2 + (size_t)(lg.y + s0 + lane) * a.DS : nullptr, vb, NV, v);
                seg_butterfly<1>(v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tot[0][q].x += v[0][q].x; tot[0][q].y += v[0][q].y; tot[0][q].z += v[0][q].z; tot[0][q].w += v[0][q].w;
                }
            }
            seg_store_vecs<1>(a, tot, vb, NV, lane, frow, flr);
        }
    }
}

static int key_bits(unsigned int maxkey) { int b = 1; while ((maxkey >> b) != 0) ++b; return b; }

}  // namespace rat

using namespace rat;

namespace {
struct ScatterLayout {
    unsigned int *k0, *v0, *k1, *v1, *hist, *totals, *counters;
    uint4 *items, *longs;
    unsigned int* done;
    float *carryF, *carryL, *carry2;
    long long n, nchunks;
    int nblk, DS;
    size_t bytes;
};
ScatterLayout scatter_layout(void* workspace, long long n, int D) {
    ScatterLayout w{};
    w.n = n;
    w.nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    w.nchunks = (n + 31) / 32;
    w.DS = round_up(D + 1, 4);
    const size_t nk = (size_t)((n + 3) / 4 * 4);
    const size_t nc = (size_t)((w.nchunks + 3) / 4 * 4);
    unsigned int* p = (unsigned int*)workspace;
    w.k0 = p; p += nk; w.v0 = p; p += nk; w.k1 = p; p += nk; w.v1 = p; p += nk;
    w.hist = p; p += ((size_t)256 * w.nblk + 3) / 4 * 4;
    w.totals = p; p += 4 * 256;                                     // [4 passes][256] per-digit key counts
    w.counters = p; p += 16;
    const size_t n2 = nc / 16 + 8;                                  // second-level records: sum ceil(m_i / 32), m_i > 32
    w.items = (uint4*)p; p += 4 * (nc + n2);
    w.longs = (uint4*)p; p += 4 * (nc / 32 + 4);
    w.done = p; p += (nc / 32 + 4 + 3) / 4 * 4;
    w.carryF = (float*)p; p += nc * w.DS;
    w.carryL = (float*)p; p += nc * w.DS;
    w.carry2 = (float*)p; p += n2 * w.DS;
    w.bytes = (size_t)((char*)p - (char*)workspace);
    return w;
}
}  // namespace

extern "C" size_t rat_emb_scatter_workspace_bytes(long long n_occ, int D) {
    return scatter_layout(nullptr, n_occ, D).bytes + 64;
}

static int scatter_check(int B, int T, int L, int F, int D, long long V_total, const void* workspace, size_t workspace_bytes,
                         const char* who) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0 && F > 0 && D > 0 && V_total > 0, "%s: bad shape", who);
    RAT_REQUIRE(D <= 128, "%s: D=%d > 128 not supported", who, D);
    const long long n = (long long)B * T * (L + 1);
    RAT_REQUIRE(n < (1ll << 31), "%s: too many occurrences", who);
    RAT_REQUIRE(L <= 31 && (long long)B * T < (1ll << 26), "%s: L=%d (<=31) or B*T too large for the packed occurrence index", who, L);
    RAT_REQUIRE(V_total + 3 < (1ll << 32), "%s: V_total too large", who);
    RAT_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= rat_emb_scatter_workspace_bytes(n, D),
                "%s: workspace missing, unaligned or too small", who);
    return RAT_OK;
}

extern "C" int rat_emb_scatter_plan(const int* ids, const int* labels, const int* col_off, const int* col_pad,
                                    const int* col_vocab, int B, int T, int L, int F, int D, long long V_total,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    int rc = scatter_check(B, T, L, F, D, V_total, workspace, workspace_bytes, "rat_emb_scatter_plan");
    if (rc != RAT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * T * (L + 1);
    ScatterLayout w = scatter_layout(workspace, n, D);
    const unsigned int V = (unsigned int)V_total;
    int grid = (int)min((n + 255) / 256, (long long)num_sms() * 16);
    k_build_keys<<<grid, 256, 0, st>>>(ids, labels, col_off, col_pad, col_vocab, n, L, T, V, w.k0, w.v0, w.totals, 4 * 256 + 16);
    RAT_CHECK_LAUNCH("k_build_keys");
    const int bits = key_bits(V + 3u);
    unsigned int *ki = w.k0, *vi = w.v0, *ko = w.k1, *vo = w.v1;
    for (int shift = 0; shift < bits; shift += 8) {
        unsigned int* tot = w.totals + (shift / 8) * 256;
        k_radix_hist<<<w.nblk, RS_THREADS, 0, st>>>(ki, n, shift, w.hist, w.nblk, tot);
        RAT_CHECK_LAUNCH("k_radix_hist");
        k_scan_digits<<<256, 256, 0, st>>>(w.hist, w.nblk, tot);
        RAT_CHECK_LAUNCH("k_scan_digits");
        k_radix_scatter<<<w.nblk, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, w.hist, w.nblk);
        RAT_CHECK_LAUNCH("k_radix_scatter");
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    PlanRunsArgs pr{ki, n, V + 3u, w.counters, w.items, w.longs, w.done};
    k_plan_runs<<<(int)((w.nchunks + 7) / 8), 256, 0, st>>>(pr);
    RAT_CHECK_LAUNCH("k_plan_runs");
    return RAT_OK;
}

extern "C" int rat_emb_scatter_reduce(const int* ids, const int* labels, const float* dblock, const float* dxemb,
                                      const float* dlogit, const int* col_off, const int* col_pad,
                                      const int* col_vocab, const int* col_field, float* g_emb, float* g_lr,
                                      float* g_label, int B, int T, int L, int F, int D, long long V_total,
                                      float drop_p, unsigned long long seed, unsigned int rng_stream, int planned,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    int rc = scatter_check(B, T, L, F, D, V_total, workspace, workspace_bytes, "rat_emb_scatter_reduce");
    if (rc != RAT_OK) return rc;
    RAT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "rat_emb_scatter_reduce: dropout p=%f", drop_p);
    cudaStream_t st = (cudaStream_t)stream;
    if (!planned) {
        rc = rat_emb_scatter_plan(ids, labels, col_off, col_pad, col_vocab, B, T, L, F, D, V_total, workspace,
                                  workspace_bytes, stream);
        if (rc != RAT_OK) return rc;
    }
    const long long n = (long long)B * T * (L + 1);
    ScatterLayout w = scatter_layout(workspace, n, D);
    const unsigned int V = (unsigned int)V_total;
    const int passes = (key_bits(V + 3u) + 7) / 8;
    const unsigned int* keys = (passes & 1) ? w.k1 : w.k0;
    const unsigned int* vals = (passes & 1) ? w.v1 : w.v0;
    SegArgs a{keys, vals, n, V + 3u, V, dblock, dxemb, dlogit, col_field, g_emb, g_lr, g_label, w.carryF, w.carryL,
              w.carry2, w.counters, w.items, w.longs, w.done, T, L, F + 1, D, F, w.DS, make_fastdiv((uint32_t)T), drop_p, seed, rng_stream, rng_step_ptr()};
    const int sgrid = (int)((w.nchunks + 7) / 8);
    if (D % 4 == 0) k_segment_scan<4><<<sgrid, 256, 0, st>>>(a);
    else if (D % 2 == 0) k_segment_scan<2><<<sgrid, 256, 0, st>>>(a);
    else k_segment_scan<1><<<sgrid, 256, 0, st>>>(a);
    RAT_CHECK_LAUNCH("k_segment_scan");
    const int fgrid = (int)std::min<long long>((w.nchunks + 7) / 8, (long long)num_sms() * 4);
    k_fixup_items<<<fgrid, 256, 0, st>>>(a);
    RAT_CHECK_LAUNCH("k_fixup_items");
    return RAT_OK;
}

// ---- row-sharded tables: sparse exchange of the reduced row gradients (all-to-all #3) ---------------------
// After the segment reduce every rank holds one reduced row per table row its local batch touched, in a dense scratch
// indexed by GLOBAL row.  Each touched row travels once, straight into its owner's receive buffer over NVLink (peer
// stores), instead of a dense reduce-scatter over the whole row space: traffic is proportional to the batch, not to V.
//   receive buffer of an owner: `world` segments; segment r (written by rank r): [0] uint32 count, records from byte 16:
//   {uint32 key, float lr, float row[D]} padded to DS2 = round_up(D + 2, 4) floats.
// Within one segment keys are unique (a rank sends a row once), so the owner can apply a segment fully in parallel; the
// segments are applied one after the other in rank order => the sum over ranks has a fixed association (deterministic).
namespace rat {
struct ShardSendArgs {
    const unsigned int* keys; long long n; unsigned int V; int rows_per_shard, world, rank, D, DS2;
    float* g_emb_full; float* g_lr_full;
    unsigned char* const* recv_peers; long long seg_bytes; unsigned int cap;
    unsigned int* counts;        // [world] local counters (zeroed by the caller's previous k_shard_counts)
    int* err;
};
__global__ void __launch_bounds__(256) k_shard_send_rows(ShardSendArgs a) {
    const int lane = threadIdx.x & 31;
    const long long p0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (p0 >= a.n) return;
    const long long p = p0 + lane;
    const unsigned int key = p < a.n ? a.keys[p] : 0xffffffffu;
    const unsigned int prev = p > 0 && p < a.n ? a.keys[p - 1] : 0xffffffffu;
    const bool head = p < a.n && key < a.V && (p == 0 || key != prev);
    unsigned int hm = __ballot_sync(0xffffffffu, head);
    while (hm) {
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        const unsigned int k = __shfl_sync(0xffffffffu, key, src);
        const int owner = (int)(k / (unsigned int)a.rows_per_shard);
        unsigned int idx = 0;
        if (lane == 0) idx = atomicAdd(&a.counts[owner], 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        float* row = a.g_emb_full + (size_t)k * a.D;
        if (idx >= a.cap) { if (lane == 0 && a.err) atomicOr(a.err, 8); continue; }
        float* rec = reinterpret_cast<float*>(a.recv_peers[owner] + (size_t)a.rank * a.seg_bytes + 16) + (size_t)idx * a.DS2;
        for (int d = lane; d < a.D; d += 32) { rec[2 + d] = row[d]; row[d] = 0.f; }
        if (lane == 0) {
            reinterpret_cast<unsigned int*>(rec)[0] = k;
            rec[1] = a.g_lr_full ? a.g_lr_full[k] : 0.f;
            if (a.g_lr_full) a.g_lr_full[k] = 0.f;
        }
    }
}
// publish the per-owner record counts into the owners' segment headers and clear the local counters
__global__ void k_shard_counts(unsigned int* counts, unsigned char* const* recv_peers, long long seg_bytes, int rank, int world) {
    const int o = threadIdx.x;
    if (o < world) {
        *reinterpret_cast<unsigned int*>(recv_peers[o] + (size_t)rank * seg_bytes) = counts[o];
        counts[o] = 0u;
    }
}
// owner side: g_local[key - row0] += record, one warp per record
__global__ void __launch_bounds__(256) k_shard_apply_rows(const unsigned char* __restrict__ seg, int D, int DS2, unsigned int row0,
                                                          int rows_per_shard, float* __restrict__ g_emb, float* __restrict__ g_lr,
                                                          int* err) {
    const unsigned int count = *reinterpret_cast<const unsigned int*>(seg);
    const float* recs = reinterpret_cast<const float*>(seg + 16);
    const int lane = threadIdx.x & 31;
    const unsigned int nw = gridDim.x * (blockDim.x >> 5);
    for (unsigned int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += nw) {
        const float* rec = recs + (size_t)i * DS2;
        const unsigned int key = reinterpret_cast<const unsigned int*>(rec)[0];
        const unsigned int r = key - row0;
        if (r >= (unsigned int)rows_per_shard) { if (lane == 0 && err) atomicOr(err, 16); continue; }
        for (int d = lane; d < D; d += 32) g_emb[(size_t)r * D + d] += rec[2 + d];
        if (lane == 0 && g_lr) g_lr[r] += rec[1];
    }
}
}  // namespace rat

extern "C" size_t rat_shard_recv_bytes(int B, int T, int L, int D, int world) {
    const size_t cap = (size_t)B * T * (L + 1);
    const size_t seg = 16 + cap * (size_t)round_up(D + 2, 4) * sizeof(float);
    return (size_t)world * ((seg + 127) / 128 * 128);
}

extern "C" int rat_shard_send_rows(void* workspace, size_t workspace_bytes, int B, int T, int L, int F, int D, long long V_total,
                                   int rows_per_shard, int world, int rank, float* g_emb_full, float* g_lr_full,
                                   const void* const* recv_peers, unsigned int* counts, int* err_flag, void* stream) {
    int rc = scatter_check(B, T, L, F, D, V_total, workspace, workspace_bytes, "rat_shard_send_rows");
    if (rc != RAT_OK) return rc;
    RAT_REQUIRE(world >= 1 && world <= 1024 && rank >= 0 && rank < world && rows_per_shard > 0 && recv_peers && counts,
                "rat_shard_send_rows: bad shard description");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * T * (L + 1);
    ScatterLayout w = scatter_layout(workspace, n, D);
    const unsigned int V = (unsigned int)V_total;
    const int passes = (key_bits(V + 3u) + 7) / 8;
    const unsigned int* keys = (passes & 1) ? w.k1 : w.k0;
    const long long seg_bytes = (long long)(rat_shard_recv_bytes(B, T, L, D, world) / world);
    ShardSendArgs a{keys, n, V, rows_per_shard, world, rank, D, round_up(D + 2, 4), g_emb_full, g_lr_full,
                    (unsigned char* const*)recv_peers, seg_bytes, (unsigned int)n, counts, err_flag};
    k_shard_send_rows<<<(int)((w.nchunks + 7) / 8), 256, 0, st>>>(a);
    RAT_CHECK_LAUNCH("k_shard_send_rows");
    k_shard_counts<<<1, 1024, 0, st>>>(counts, (unsigned char* const*)recv_peers, seg_bytes, rank, world);
    RAT_CHECK_LAUNCH("k_shard_counts");
    return RAT_OK;
}

extern "C" int rat_shard_apply_rows(const void* recv_local, int B, int T, int L, int D, int world, int rank, int rows_per_shard,
                                    float* g_emb_local, float* g_lr_local, int* err_flag, void* stream) {
    RAT_REQUIRE(recv_local && world >= 1 && rows_per_shard > 0, "rat_shard_apply_rows: bad arguments");
    const size_t seg_bytes = rat_shard_recv_bytes(B, T, L, D, world) / world;
    const int grid = num_sms() * 4;
    for (int src = 0; src < world; ++src) {           // rank order: the sum over ranks has one fixed association
        k_shard_apply_rows<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)recv_local + (size_t)src * seg_bytes, D,
                                                                   round_up(D + 2, 4), (unsigned int)rank * (unsigned int)rows_per_shard,
                                                                   rows_per_shard, g_emb_local, g_lr_local, err_flag);
        RAT_CHECK_LAUNCH("k_shard_apply_rows");
    }
    return RAT_OK;
}

// exposed for tests: stable sort of (key,val) pairs with the same kernels
extern "C" int rat_radix_sort_pairs(unsigned int* keys, unsigned int* vals, unsigned int* keys_tmp,
                                    unsigned int* vals_tmp, unsigned int* hist, long long n, int bits,
                                    int* result_in_tmp, void* stream) {
    RAT_REQUIRE(n > 0 && bits > 0 && bits <= 32, "rat_radix_sort_pairs: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    unsigned int *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
    int flips = 0;
    for (int shift = 0; shift < bits; shift += 8) {
        k_radix_hist<<<nblk, RS_THREADS, 0, st>>>(ki, n, shift, hist, nblk, nullptr);
        k_scan_exclusive<<<1, 1024, 0, st>>>(hist, 256 * nblk);
        k_radix_scatter<<<nblk, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, hist, nblk);
        RAT_CHECK_LAUNCH("radix pass");
        unsigned int* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
        ++flips;
    }
    *result_in_tmp = flips & 1;
    return RAT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device-side validation metrics.  Replaces the per-pass host work of evaluate_metrics
// (fuxictr/metrics.py:21-41: sklearn roc_auc_score + log_loss with predictions clipped to [1e-7, 1-1e-7]).
//   AUC  = sum over positives of (#negatives with a smaller score + 0.5 #negatives with an equal score) / (P N):
//          stable radix sort of the score bits, prefix count of negatives, tie groups located by binary search;
//          the sum is accumulated as an exact integer (2x contributions, 64-bit atomics: order independent).
//   logloss in float64, per-block partial sums reduced in block order (deterministic).
namespace rat {

__global__ void k_metric_keys(const float* __restrict__ pred, const float* __restrict__ y, long long n,
                              unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                              double* __restrict__ ll_part) {
    __shared__ double wsum[8];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float p = pred[i];
        const unsigned int b = __float_as_uint(p);
        keys[i] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);          // order-preserving map of IEEE floats
        const bool pos = y[i] > 0.5f;
        vals[i] = pos ? 1u : 0u;
        const double pc = fmin(fmax((double)p, 1e-7), 1.0 - 1e-7);
        s -= pos ? log(pc) : log(1.0 - pc);
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
        ll_part[blockIdx.x] = t;
    }
}

// negatives per 1024-element block of the sorted order
__global__ void __launch_bounds__(1024) k_metric_block_neg(const unsigned int* __restrict__ lab, long long n,
                                                           unsigned int* __restrict__ blockneg) {
    __shared__ unsigned int ws[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const unsigned int neg = (i < n && lab[i] == 0u) ? 1u : 0u;
    const unsigned int c = __popc(__ballot_sync(0xffffffffu, neg));
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int t = 0; for (int w = 0; w < 32; ++w) t += ws[w]; blockneg[blockIdx.x] = t; }
}

// cneg[i] = number of negatives at sorted positions < i  (blockneg already exclusive-scanned)
__global__ void __launch_bounds__(1024) k_metric_cneg(const unsigned int* __restrict__ lab, long long n,
                                                      const unsigned int* __restrict__ blockneg,
                                                      unsigned int* __restrict__ cneg) {
    __shared__ unsigned int ws[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int neg = (i < n && lab[i] == 0u) ? 1u : 0u;
    const unsigned int bal = __ballot_sync(0xffffffffu, neg);
    if (lane == 0) ws[warp] = __popc(bal);
    __syncthreads();
    unsigned int base = blockneg[blockIdx.x];
    for (int w = 0; w < warp; ++w) base += ws[w];
    if (i < n) cneg[i] = base + __popc(bal & ((1u << lane) - 1u));
}

// acc[0] += 2 * (#neg below) + (#neg tied) over positives ; acc[1] += #positives
__global__ void k_metric_auc(const unsigned int* __restrict__ keys, const unsigned int* __restrict__ lab,
                             const unsigned int* __restrict__ cneg, long long n, unsigned long long* __restrict__ acc) {
    unsigned long long s = 0, np = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (lab[i] == 0u) continue;
        const unsigned int k = keys[i];
        long long lo = i, hi = i;                                        // tie group [lo, hi]
        if (i > 0 && keys[i - 1] == k) {                                 // lower bound of k in [0, i)
            long long a = 0, b = i;
            while (a < b) { const long long m = (a + b) >> 1; if (keys[m] < k) a = m + 1; else b = m; }
            lo = a;
        }
        if (i + 1 < n && keys[i + 1] == k) {                             // upper bound of k in (i, n)
            long long a = i + 1, b = n;
            while (a < b) { const long long m = (a + b) >> 1; if (keys[m] <= k) a = m + 1; else b = m; }
            hi = a - 1;
        }
        const unsigned long long below = cneg[lo];
        const unsigned long long upto = (unsigned long long)cneg[hi] + (lab[hi] == 0u ? 1u : 0u);   // negatives in [0, hi]
        s += 2ull * below + (upto - below);
        ++np;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); np += __shfl_xor_sync(0xffffffffu, np, o); }
    if ((threadIdx.x & 31) == 0) { if (s) atomicAdd(&acc[0], s); if (np) atomicAdd(&acc[1], np); }
}

__global__ void k_metric_finalize(const unsigned long long* __restrict__ acc, const double* __restrict__ ll_part, int nparts,
                                  long long n, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double ll = 0.0;
    for (int i = 0; i < nparts; ++i) ll += ll_part[i];
    const double P = (double)acc[1], N = (double)n - P;
    out[0] = (P > 0 && N > 0) ? (double)acc[0] / (2.0 * P * N) : nan("");
    out[1] = ll / (double)n;
    out[2] = P;
    out[3] = N;
}

struct MetricLayout { unsigned int *k0, *v0, *k1, *v1, *hist, *blockneg, *cneg; double* ll_part; unsigned long long* acc; size_t bytes; int nblk, nb1k; };
static MetricLayout metric_layout(void* ws, long long n) {
    MetricLayout w{};
    const size_t nk = (size_t)((n + 3) / 4 * 4);
    w.nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    w.nb1k = (int)((n + 1023) / 1024);
    unsigned int* p = (unsigned int*)ws;
    w.k0 = p; p += nk; w.v0 = p; p += nk; w.k1 = p; p += nk; w.v1 = p; p += nk;
    w.hist = p; p += ((size_t)256 * w.nblk + 3) / 4 * 4;
    w.blockneg = p; p += ((size_t)w.nb1k + 4) / 4 * 4;
    w.cneg = p; p += nk;
    w.ll_part = (double*)p; p += 2 * 1024;
    w.acc = (unsigned long long*)p; p += 8;
    w.bytes = (size_t)((char*)p - (char*)ws);
    return w;
}

}  // namespace rat

extern "C" size_t rat_auc_logloss_workspace_bytes(long long n) { return rat::metric_layout(nullptr, n).bytes + 64; }

extern "C" int rat_auc_logloss(const float* y_pred, const float* y_true, long long n, double* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(n > 0 && n < (1ll << 31), "rat_auc_logloss: bad n");
    RAT_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0 && workspace_bytes >= rat_auc_logloss_workspace_bytes(n),
                "rat_auc_logloss: workspace missing, unaligned or too small");
    cudaStream_t st = (cudaStream_t)stream;
    MetricLayout w = metric_layout(workspace, n);
    const int kgrid = (int)std::min<long long>((n + 255) / 256, 1024);
    cudaError_t e = cudaMemsetAsync(w.acc, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(metric accumulators)");
    k_metric_keys<<<kgrid, 256, 0, st>>>(y_pred, y_true, n, w.k0, w.v0, w.ll_part);
    RAT_CHECK_LAUNCH("k_metric_keys");
    unsigned int *ki = w.k0, *vi = w.v0, *ko = w.k1, *vo = w.v1;
    for (int shift = 0; shift < 32; shift += 8) {
        k_radix_hist<<<w.nblk, RS_THREADS, 0, st>>>(ki, n, shift, w.hist, w.nblk, nullptr);
        k_scan_exclusive<<<1, 1024, 0, st>>>(w.hist, 256 * w.nblk);
        k_radix_scatter<<<w.nblk, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, w.hist, w.nblk);
        RAT_CHECK_LAUNCH("metric radix pass");
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    k_metric_block_neg<<<w.nb1k, 1024, 0, st>>>(vi, n, w.blockneg);
    k_scan_exclusive<<<1, 1024, 0, st>>>(w.blockneg, w.nb1k);
    k_metric_cneg<<<w.nb1k, 1024, 0, st>>>(vi, n, w.blockneg, w.cneg);
    k_metric_auc<<<kgrid, 256, 0, st>>>(ki, vi, w.cneg, n, w.acc);
    k_metric_finalize<<<1, 32, 0, st>>>(w.acc, w.ll_part, kgrid, n, out);
    RAT_CHECK_LAUNCH("metric reduce");
    return RAT_OK;
}
