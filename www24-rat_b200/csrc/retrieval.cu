// BM25 top-K retrieval on the device: the compare-scan + top-K of the reference's
// BM25_topk_retrieval_v4 (fuxictr/datasets/data_utils.py:773-1064):
//     score[b][n] = sum_f (qry[b][f] == db[n][f]) * IDF_f(qry[b][f])            (data_utils.py:945-950, float64, field order)
//     exact-match columns (data_utils.py:862-876): only db rows that agree with the query on ALL of them are candidates
//     and their score is score + 1 (data_utils.py:947) -- or exactly 1 when the caller asks for unit scores
//     (pure exact matching, data_utils.py:912-917,1033-1038);
//     the K best candidates per query, score descending; score 0 = no match -> index -1 (sort_results, :787-797).
// HBM / issue bound integer work: every (query, db row) pair costs F compares + selects + float64 adds; nothing here is
// GEMM shaped.  Layout: db [N][E + F] int32 row-major (the E exact-match columns first), one WARP per query, the lanes
// stride over the db rows of the CTA's range (rows staged once per CTA in shared memory, odd row stride => conflict free),
// the warp's running top-K lives in registers (lane j = j-th best) and is touched only when a row beats the K-th best,
// which after the first few hundred rows is rare.  Ties: the row with the smaller db index wins (rows are visited in index
// order and only a strictly better score displaces), so the result is a pure function of the inputs.  The db range is split
// over gridDim.y CTAs when there are few queries; a second kernel merges the per-split lists.
#include "common.cuh"
#include <cstdint>
#include <algorithm>

namespace rat {

constexpr int BM_WARPS = 8;            // queries per CTA
constexpr int BM_ROWS = 1024;          // db rows staged per iteration
constexpr int BM_MAXC = 24;            // max columns (E + F)

struct Bm25Args {
    const int* db; const int* qry; const double* qry_idf;
    long long N, Q;
    int E, F, K, unit_scores, prefer_last;
    long long rows_per_split;
    double* pval; long long* pidx;      // [Q][nsplit][K] partial lists (or the final outputs when nsplit == 1)
    int nsplit;
};

// insert (v, i) into the warp's sorted list (lane j holds the j-th best); returns the new K-th best score
__device__ __forceinline__ void bm_insert(double& lv, long long& li, double v, long long i, int K, int lane, bool after_equal) {
    const unsigned int FULL = 0xffffffffu;
    // position = number of entries that stay in front of the new one
    const bool front = after_equal ? (lv >= v) : (lv > v);
    const int pos = __popc(__ballot_sync(FULL, lane < K && front));
    const double uv = __shfl_up_sync(FULL, lv, 1);
    const long long ui = __shfl_up_sync(FULL, li, 1);
    if (lane > pos) { lv = uv; li = ui; }
    if (lane == pos) { lv = v; li = i; }
}

__global__ void __launch_bounds__(BM_WARPS * 32) k_bm25_scan(Bm25Args a) {
    extern __shared__ int bm_rows[];                     // [BM_ROWS][CS]  (CS odd)
    const int C = a.E + a.F;
    const int CS = C | 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * BM_WARPS + warp;
    const bool qv = b < a.Q;
    int q[BM_MAXC];
    double w[BM_MAXC];
#pragma unroll
    for (int f = 0; f < BM_MAXC; ++f) {
        q[f] = (qv && f < C) ? a.qry[b * C + f] : -1;
        w[f] = (qv && f >= a.E && f < C) ? a.qry_idf[b * a.F + (f - a.E)] : 0.0;
    }
    double lv = 0.0;                                     // lane j: j-th best score so far (0 = empty)
    long long li = -1;
    double thr = 0.0;                                    // K-th best (0 while the list is not full)
    const long long n0 = (long long)blockIdx.y * a.rows_per_split;
    const long long n1 = min(a.N, n0 + a.rows_per_split);
    for (long long base = n0; base < n1; base += BM_ROWS) {
        const int nr = (int)min((long long)BM_ROWS, n1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * C; i += blockDim.x) {
            const int r = i / C, c = i - r * C;
            bm_rows[r * CS + c] = a.db[(base * C) + i];
        }
        __syncthreads();
        if (!qv) continue;
        for (int r0 = 0; r0 < nr; r0 += 32) {
            const int r = r0 + lane;
            double s = 0.0;
            if (r < nr) {
                const int* row = bm_rows + r * CS;
                bool cand = true;
#pragma unroll
                for (int f = 0; f < BM_MAXC; ++f)
                    if (f < a.E) cand = cand && (row[f] == q[f]);
                if (cand) {
                    // the association of torch's float64 sum(-1) over < 20 contiguous elements (oracle/bm25_oracle.py::_scores):
                    // four interleaved accumulators over the full groups of four, the < 4 leftover terms summed in order
                    // and added to accumulator 0, then ((a0 + a1) + a2) + a3.  Adding 0.0 for a mismatch is exact.
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, tail = 0.0;
                    const int nfull = a.F & ~3;
#pragma unroll
                    for (int f = 0; f < BM_MAXC; ++f)
                        if (f >= a.E && f < C) {
                            const double term = (row[f] == q[f]) ? w[f] : 0.0;
                            const int k = f - a.E;
                            if (k >= nfull) tail += term;
                            else if ((k & 3) == 0) a0 += term;
                            else if ((k & 3) == 1) a1 += term;
                            else if ((k & 3) == 2) a2 += term;
                            else a3 += term;
                        }
                    if (nfull < a.F) a0 += tail;
                    s = ((a0 + a1) + a2) + a3;
                    if (a.E > 0) s += 1.0;
                    if (a.unit_scores) s = 1.0;
                }
            }
            // a row enters the list when it is strictly better than the K-th best (ties keep the earlier row); with
            // prefer_last an equal score enters too and goes in FRONT of its equals (the list ends up holding the last K)
            unsigned int m = __ballot_sync(0xffffffffu, a.prefer_last ? (s > 0.0 && s >= thr) : (s > thr));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const double v = __shfl_sync(0xffffffffu, s, src);
                if (a.prefer_last ? (v >= thr) : (v > thr)) {     // thr may have risen inside this batch
                    bm_insert(lv, li, v, base + r0 + src, a.K, lane, !a.prefer_last);
                    thr = __shfl_sync(0xffffffffu, lv, a.K - 1);
                }
            }
        }
    }
    if (qv && lane < a.K) {
        a.pval[(b * a.nsplit + blockIdx.y) * a.K + lane] = lv;
        a.pidx[(b * a.nsplit + blockIdx.y) * a.K + lane] = lv > 0.0 ? li : -1;
    }
}

// merge the per-split lists of one query (one warp per query): K rounds of (max score, then smallest / largest index)
__global__ void __launch_bounds__(256) k_bm25_merge(const double* pval, const long long* pidx, long long Q, int nsplit, int K,
                                                    int prefer_last, double* values, long long* indices, long long* lens) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * 8 + warp;
    if (b >= Q) return;
    const int total = nsplit * K;
    const double* pv = pval + b * total;
    const long long* pi = pidx + b * total;
    // heads of the nsplit sorted lists: lane s (and s + 32, ...) owns split s; splits are in db-index order
    int found = 0;
    long long last_i = prefer_last ? (1LL << 62) : -1;
    double last_v = 1e300;
    for (int k = 0; k < K; ++k) {
        // best remaining candidate: score desc, then index asc (desc with prefer_last), strictly after (last_v, last_i)
        double bv = 0.0;
        long long bi = -1;
        for (int e = lane; e < total; e += 32) {
            const double v = pv[e];
            const long long i = pi[e];
            if (v <= 0.0 || i < 0) continue;
            const bool after = v < last_v || (v == last_v && (prefer_last ? i < last_i : i > last_i));
            if (!after) continue;
            const bool better = v > bv || (v == bv && (bi < 0 || (prefer_last ? i > bi : i < bi)));
            if (better) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool better = oi >= 0 && (ov > bv || (ov == bv && (bi < 0 || (prefer_last ? oi > bi : oi < bi))));
            if (better) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            values[b * K + k] = bi >= 0 ? bv : 0.0;
            indices[b * K + k] = bi;
        }
        if (bi >= 0) { ++found; last_v = bv; last_i = bi; } else { last_v = -1.0; }
    }
    if (lane == 0) lens[b] = found;
}

static int bm25_nsplit(long long N, long long Q) {
    const long long qblocks = (Q + BM_WARPS - 1) / BM_WARPS;
    long long want = std::max<long long>(1, (4LL * num_sms() + qblocks - 1) / qblocks);
    const long long max_split = std::max<long long>(1, (N + 4 * BM_ROWS - 1) / (4 * BM_ROWS));
    return (int)std::min<long long>(std::min<long long>(want, max_split), 64);
}

}  // namespace rat

using namespace rat;

extern "C" size_t rat_bm25_topk_workspace_bytes(long long N, long long Q, int K) {
    const int ns = bm25_nsplit(N, Q);
    return ns == 1 ? 16 : (size_t)Q * ns * K * (sizeof(double) + sizeof(long long));
}

extern "C" int rat_bm25_topk(const int* db, long long N, const int* qry, const double* qry_idf, long long Q, int E, int F, int K,
                             int unit_scores, int prefer_last, double* values, long long* indices, long long* lens,
                             void* workspace, size_t workspace_bytes, void* stream) {
    RAT_REQUIRE(N >= 0 && Q >= 0 && E >= 0 && F >= 0 && E + F >= 1, "rat_bm25_topk: bad shape");
    RAT_REQUIRE(E + F <= BM_MAXC, "rat_bm25_topk: %d columns > %d not supported", E + F, BM_MAXC);
    RAT_REQUIRE(F <= 19, "rat_bm25_topk: %d scored columns > 19 (the float64 summation order of the reference is only reproduced below 20)", F);
    RAT_REQUIRE(K >= 1 && K <= 32, "rat_bm25_topk: topK=%d must be in [1, 32]", K);
    if (Q == 0) return RAT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = bm25_nsplit(std::max<long long>(N, 1), Q);
    RAT_REQUIRE(workspace_bytes >= rat_bm25_topk_workspace_bytes(N, Q, K), "rat_bm25_topk: workspace too small");
    Bm25Args a{};
    a.db = db; a.qry = qry; a.qry_idf = qry_idf; a.N = N; a.Q = Q; a.E = E; a.F = F; a.K = K;
    a.unit_scores = unit_scores; a.prefer_last = prefer_last; a.nsplit = ns;
    a.rows_per_split = std::max<long long>(1, (N + ns - 1) / ns);
    // with one split the scan writes its sorted list into the (values, indices) outputs and the merge runs in place
    double* pval = ns == 1 ? values : reinterpret_cast<double*>(workspace);
    long long* pidx = ns == 1 ? indices : reinterpret_cast<long long*>(pval + (size_t)Q * ns * K);
    a.pval = pval; a.pidx = pidx;
    const int CS = (E + F) | 1;
    const size_t smem = (size_t)BM_ROWS * CS * sizeof(int);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(k_bm25_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_bm25_scan)");
        attr_smem = smem;
    }
    const long long qblocks = (Q + BM_WARPS - 1) / BM_WARPS;
    RAT_REQUIRE(qblocks <= 0x7fffffffLL, "rat_bm25_topk: too many queries in one call");
    k_bm25_scan<<<dim3((unsigned)qblocks, (unsigned)ns), BM_WARPS * 32, smem, st>>>(a);
    RAT_CHECK_LAUNCH("k_bm25_scan");
    if (ns == 1) {
        // lens = number of valid entries; the list is already final
        k_bm25_merge<<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(pval, pidx, Q, 1, K, prefer_last, values, indices, lens);
    } else {
        k_bm25_merge<<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(pval, pidx, Q, ns, K, prefer_last, values, indices, lens);
    }
    RAT_CHECK_LAUNCH("k_bm25_merge");
    return RAT_OK;
}
