// K6: deterministic sorted segment-reduce of the embedding gradient.
//
// Replaces ATen embedding_dense_backward reached from loss.backward() (base_model.py:223) for every per-field
// nn.Embedding of EmbeddingDictLayer (embedding.py:79-100) and LR_Layer (shallow.py:31), plus the 3-row label
// table (RAT_m2.py:64).  Pipeline per step:
//   k_build_keys        key[i] = table row of occurrence i=(b,t,l)  (padding ids -> sentinel), val[i] = i
//   k_radix_{hist,scan,scatter} x passes   stable LSD radix sort (8-bit digits) of (key, val)
//   k_segment_reduce    one warp per 32 sorted positions: runs wholly inside the chunk are summed in sorted
//                       (= canonical occurrence) order and stored once; runs that cross chunk boundaries leave
//                       partials that k_segment_fixup adds up in chunk order  => bitwise deterministic
//   k_label_grad        per-CTA partial sums of the label-token rows, reduced in fixed order
// The gradient of occurrence (b,t,l) is dBlock[b,t,1+field(l),:] (+ dXemb[b,field(l),:] for the target row t=0,
// the DNN path) and, for the LR table, dlogit[b] for t=0.
#include <algorithm>
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace rat {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block

// val = (bt << 6) | (t == 0 ? 32 : 0) | l : the segment reduce decodes an occurrence with shifts only (L <= 32);
// because (bt, l) is lexicographic in the occurrence index the stable sort still yields the canonical order.
__global__ void k_build_keys(const int* __restrict__ ids, const int* __restrict__ col_off,
                             const int* __restrict__ col_pad, const int* __restrict__ col_vocab, long long n, int L, int T,
                             unsigned int sentinel, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                             unsigned int* __restrict__ totals, int ntotals) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int bt = (unsigned int)(i / L);
        const int l = (int)(i - (long long)bt * L);
        const int id = ids[i];
        unsigned int k = sentinel;
        if (id >= 0 && id < col_vocab[l] && id != col_pad[l]) k = (unsigned int)(col_off[l] + id);
        keys[i] = k;
        vals[i] = (bt << 6) | ((bt % (unsigned int)T) == 0u ? 32u : 0u) | (unsigned int)l;
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
    const unsigned int* keys; const unsigned int* vals; long long n; unsigned int sentinel;
    const float* dblock;     // [B,T,N,D]
    const float* dxemb;      // [B,F*D] or nullptr
    const float* dlogit;     // [B] or nullptr
    const int* col_field;    // [L]
    float* g_emb;            // [V,D]
    float* g_lr;             // [V] or nullptr
    float* carryF; float* carryL;   // [nchunks][D+1]
    int T, L, N, D, F;
};

template <int NPER>   // floats per lane: D <= 32*NPER
__device__ __forceinline__ void seg_load_add(const SegArgs& a, unsigned int src, int lane, float (&acc)[NPER], float& lr) {
    const int l = (int)(src & 31u);
    const long long bt = src >> 6;
    const int t = (src & 32u) ? 0 : 1;                    // only "is this the target row" matters
    const long long b = t == 0 ? bt / a.T : 0;
    const int f = a.col_field[l];
    const float* g = a.dblock + ((bt * a.N) + 1 + f) * a.D;
#pragma unroll
    for (int k = 0; k < NPER; ++k) {
        const int d = lane + 32 * k;
        if (d < a.D) {
            float v = g[d];
            if (t == 0 && a.dxemb) v += a.dxemb[(b * a.F + f) * a.D + d];
            acc[k] += v;
        }
    }
    if (t == 0 && a.dlogit) lr += a.dlogit[b];
}

template <int NPER>
__global__ void __launch_bounds__(256) k_segment_reduce(SegArgs a) {
    const int lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long p0 = chunk * 32;
    if (p0 >= a.n) return;
    const long long p = p0 + lane;
    const unsigned int mykey = p < a.n ? a.keys[p] : a.sentinel;
    const unsigned int mysrc = p < a.n ? a.vals[p] : 0u;
    const unsigned int prev_key = p0 > 0 ? a.keys[p0 - 1] : a.sentinel;           // sentinel == "different"
    const unsigned int next_key = p0 + 32 < a.n ? a.keys[p0 + 32] : a.sentinel;
    float acc[NPER];
    float lr = 0.f;
#pragma unroll
    for (int k = 0; k < NPER; ++k) acc[k] = 0.f;
    int run_start = 0;
    for (int j = 0; j < 32; ++j) {
        const unsigned int kj = __shfl_sync(0xffffffffu, mykey, j);
        if (kj == a.sentinel) break;                                           // sentinels are sorted last
        const unsigned int sj = __shfl_sync(0xffffffffu, mysrc, j);
        seg_load_add<NPER>(a, sj, lane, acc, lr);
        const unsigned int kn = j < 31 ? __shfl_sync(0xffffffffu, mykey, j + 1) : next_key;
        const bool run_ends_here = (j == 31) || (kn != kj);
        if (run_ends_here) {
            const bool continues_from_prev = (run_start == 0) && (p0 > 0) && (prev_key == kj);
            const bool spans_forward = (j == 31) && (next_key == kj);
            float* dst;
            float* dst_lr;
            if (continues_from_prev) { dst = a.carryF + chunk * (a.D + 1); dst_lr = dst + a.D; }
            else if (spans_forward) { dst = a.carryL + chunk * (a.D + 1); dst_lr = dst + a.D; }
            else { dst = a.g_emb + (size_t)kj * a.D; dst_lr = a.g_lr ? a.g_lr + kj : nullptr; }
#pragma unroll
            for (int k = 0; k < NPER; ++k) { const int d = lane + 32 * k; if (d < a.D) dst[d] = acc[k]; acc[k] = 0.f; }
            if (lane == 0 && dst_lr) *dst_lr = lr;
            lr = 0.f;
            run_start = j + 1;
        }
    }
}

// Runs that cross chunk boundaries.  One BLOCK per chunk; only blocks whose chunk holds the HEAD of a forward-spanning
// run do work: the threads find the last chunk of the run in parallel, the 8 warps add contiguous sub-ranges of the
// per-chunk partials (lane = embedding dimension), and the 8 warp sums are combined in warp order.  For a given
// input the partition is fixed => bitwise deterministic; a hot id spanning hundreds of chunks costs ~m/8 dependent
// loads instead of m.
template <int NPER>
__global__ void __launch_bounds__(256) k_segment_fixup(SegArgs a) {
    __shared__ long long end_s;
    __shared__ float part[8][32 * NPER + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nchunks = (a.n + 31) / 32;
    const long long c = blockIdx.x;
    if (c >= nchunks - 1) return;                                   // the last chunk cannot span forward
    const long long last = c * 32 + 31;
    const unsigned int key = a.keys[last];
    if (key == a.sentinel || a.keys[last + 1] != key) return;       // last run does not span forward
    const bool covers_start = a.keys[c * 32] == key;
    if (covers_start && c > 0 && a.keys[c * 32 - 1] == key) return; // it continues from an earlier chunk: not the head
    // ---- last chunk `end` of the run: the first cc > c that is the final chunk or does not end inside the run
    if (threadIdx.x == 0) end_s = nchunks - 1;
    __syncthreads();
    for (long long base = c + 1; base < nchunks; base += 256) {
        const long long cc = base + threadIdx.x;
        bool stop = false;
        if (cc < nchunks) {
            const long long cl = cc * 32 + 31;
            stop = (cl >= a.n - 1) || a.keys[cl] != key || a.keys[cl + 1] != key;
        }
        if (stop) atomicMin(&end_s, cc);
        __syncthreads();
        if (end_s < base + 256) break;
    }
    __syncthreads();
    const long long end = end_s;
    // ---- partial sums: warp w adds chunks c+1+w*per .. (fixed partition), lane = dimension
    const long long m = end - c;                                    // number of carryF records
    const long long per = (m + 7) / 8;
    const long long lo = c + 1 + warp * per, hi = min(end + 1, lo + per);
    float acc[NPER], lr = 0.f;
#pragma unroll
    for (int k = 0; k < NPER; ++k) acc[k] = 0.f;
    for (long long cc = lo; cc < hi; ++cc) {
        const float* f = a.carryF + cc * (a.D + 1);
#pragma unroll
        for (int k = 0; k < NPER; ++k) { const int d = lane + 32 * k; if (d < a.D) acc[k] += f[d]; }
        lr += f[a.D];
    }
#pragma unroll
    for (int k = 0; k < NPER; ++k) part[warp][lane + 32 * k] = acc[k];
    if (lane == 0) part[warp][32 * NPER] = lr;
    __syncthreads();
    if (warp == 0) {
        const float* h = a.carryL + c * (a.D + 1);
        float* dst = a.g_emb + (size_t)key * a.D;
#pragma unroll
        for (int k = 0; k < NPER; ++k) {
            const int d = lane + 32 * k;
            if (d < a.D) {
                float s = h[d];
                for (int w = 0; w < 8; ++w) s += part[w][d];
                dst[d] = s;
            }
        }
        if (lane == 0 && a.g_lr) {
            float s = h[a.D];
            for (int w = 0; w < 8; ++w) s += part[w][32 * NPER];
            a.g_lr[key] = s;
        }
    }
}

// label-token gradient: partial[cta][lab][d] = sum over the cta's (b,t) rows with labels[b,t]==lab of dblock[b,t,0,d]
__global__ void __launch_bounds__(256) k_label_grad(const float* __restrict__ dblock, const int* __restrict__ labels,
                                                    long long nrows, int N, int D, float* __restrict__ partials) {
    extern __shared__ float sm[];           // [8 warps][3][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < nw * 3 * D; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long per = (nrows + gridDim.x - 1) / gridDim.x;
    const long long beg = blockIdx.x * per, end = min(nrows, beg + per);
    for (long long r = beg + warp; r < end; r += nw) {
        const int lab = labels[r];
        if (lab < 0 || lab > 2) continue;
        const float* g = dblock + r * (long long)N * D;
        float* dst = sm + (warp * 3 + lab) * D;
        for (int d = lane; d < D; d += 32) dst[d] += g[d];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += sm[w * 3 * D + i];
        partials[(size_t)blockIdx.x * 3 * D + i] = s;
    }
}
__global__ void k_label_reduce(const float* __restrict__ partials, int nparts, int len, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < nparts; ++c) s += partials[(size_t)c * len + i];
        out[i] = s;
    }
}

static int key_bits(unsigned int maxkey) { int b = 1; while ((maxkey >> b) != 0) ++b; return b; }

}  // namespace rat

using namespace rat;

extern "C" size_t rat_emb_scatter_workspace_bytes(long long n_occ, int D) {
    // keys/vals double buffers + histogram + carries + label partials
    const long long nblk = (n_occ + RS_TILE - 1) / RS_TILE;
    const long long nchunks = (n_occ + 31) / 32;
    size_t b = 0;
    b += 4 * (size_t)round_up((int)n_occ, 4) * sizeof(unsigned int);
    b += (size_t)256 * nblk * sizeof(unsigned int) + 64 + 4 * 256 * sizeof(unsigned int);
    b += 2 * (size_t)nchunks * (D + 1) * sizeof(float) + 64;
    b += (size_t)num_sms() * 3 * D * sizeof(float) + 64;
    return b;
}

extern "C" int rat_emb_scatter_reduce(const int* ids, const int* labels, const float* dblock, const float* dxemb,
                                      const float* dlogit, const int* col_off, const int* col_pad,
                                      const int* col_vocab, const int* col_field, float* g_emb, float* g_lr,
                                      float* g_label, int B, int T, int L, int F, int D, long long V_total,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0 && F > 0 && D > 0 && V_total > 0, "rat_emb_scatter_reduce: bad shape");
    RAT_REQUIRE(D <= 128, "rat_emb_scatter_reduce: D=%d > 128 not supported", D);
    const long long n = (long long)B * T * L;
    RAT_REQUIRE(n < (1ll << 31), "rat_emb_scatter_reduce: too many occurrences");
    RAT_REQUIRE(L <= 32 && (long long)B * T < (1ll << 26), "rat_emb_scatter_reduce: L=%d (<=32) or B*T too large for the packed occurrence index", L);
    RAT_REQUIRE(workspace && workspace_bytes >= rat_emb_scatter_workspace_bytes(n, D), "rat_emb_scatter_reduce: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    const long long nchunks = (n + 31) / 32;
    const size_t nk = (size_t)round_up((int)n, 4);
    unsigned int* k0 = (unsigned int*)workspace;
    unsigned int* v0 = k0 + nk;
    unsigned int* k1 = v0 + nk;
    unsigned int* v1 = k1 + nk;
    unsigned int* hist = v1 + nk;
    unsigned int* totals = hist + (size_t)256 * nblk + 16;          // [4 passes][256] per-digit key counts
    float* carryF = (float*)(totals + 4 * 256);
    float* carryL = carryF + (size_t)nchunks * (D + 1);
    float* lab_part = carryL + (size_t)nchunks * (D + 1) + 16;
    const unsigned int sentinel = (unsigned int)V_total;
    const int N = F + 1;

    int grid = (int)min((n + 255) / 256, (long long)num_sms() * 16);
    k_build_keys<<<grid, 256, 0, st>>>(ids, col_off, col_pad, col_vocab, n, L, T, sentinel, k0, v0, totals, 4 * 256);
    RAT_CHECK_LAUNCH("k_build_keys");
    const int bits = key_bits(sentinel);
    unsigned int *ki = k0, *vi = v0, *ko = k1, *vo = v1;
    for (int shift = 0; shift < bits; shift += 8) {
        unsigned int* tot = totals + (shift / 8) * 256;
        k_radix_hist<<<nblk, RS_THREADS, 0, st>>>(ki, n, shift, hist, nblk, tot);
        RAT_CHECK_LAUNCH("k_radix_hist");
        k_scan_digits<<<256, 256, 0, st>>>(hist, nblk, tot);
        RAT_CHECK_LAUNCH("k_scan_digits");
        k_radix_scatter<<<nblk, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, hist, nblk);
        RAT_CHECK_LAUNCH("k_radix_scatter");
        unsigned int* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    SegArgs a{ki, vi, n, sentinel, dblock, dxemb, dlogit, col_field, g_emb, g_lr, carryF, carryL, T, L, N, D, F};
    const int wpb = 8;
    const int sgrid = (int)((nchunks + wpb - 1) / wpb);
    const int fgrid = (int)std::max<long long>(1, nchunks - 1);      // one block per chunk that may hold a run head
    if (D <= 32) { k_segment_reduce<1><<<sgrid, 256, 0, st>>>(a); RAT_CHECK_LAUNCH("k_segment_reduce");
                   k_segment_fixup<1><<<fgrid, 256, 0, st>>>(a); }
    else if (D <= 64) { k_segment_reduce<2><<<sgrid, 256, 0, st>>>(a); RAT_CHECK_LAUNCH("k_segment_reduce");
                        k_segment_fixup<2><<<fgrid, 256, 0, st>>>(a); }
    else { k_segment_reduce<4><<<sgrid, 256, 0, st>>>(a); RAT_CHECK_LAUNCH("k_segment_reduce");
           k_segment_fixup<4><<<fgrid, 256, 0, st>>>(a); }
    RAT_CHECK_LAUNCH("k_segment_fixup");
    if (g_label) {
        const long long nrows = (long long)B * T;
        const int lgrid = (int)min((nrows + 255) / 256, (long long)num_sms());
        k_label_grad<<<lgrid, 256, (size_t)8 * 3 * D * sizeof(float), st>>>(dblock, labels, nrows, N, D, lab_part);
        RAT_CHECK_LAUNCH("k_label_grad");
        k_label_reduce<<<1, 128, 0, st>>>(lab_part, lgrid, 3 * D, g_label);
        RAT_CHECK_LAUNCH("k_label_reduce");
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
