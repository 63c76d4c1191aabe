// K0/K1: retrieval-set assembly + fused FeatureEmbedding gather.
//
// Replaces (reference, /root/reference): Dataset.__getitem__ fuxictr/pytorch/data_generator.py:66-78,
// BaseModel.inputs_to_device base_model.py:125-133, EmbeddingDictLayer.forward embedding.py:158-178,
// MaskedSumPooling sequence.py:36-38, the label-token concat RAT_m2.py:115-126, nn.Dropout RAT_m2.py:135
// and LR_Layer.forward shallow.py:36-45 -- in ONE pass over the output block.
#include "common.cuh"
#include "../../include/rat_b200.h"
#include <algorithm>

namespace rat {

// ---- K0a: wire format (float64 ids/labels from the reference DataLoader) -> int32 ------------------------
__global__ void k_convert_wire(const double* __restrict__ X, const double* __restrict__ y, int* __restrict__ ids,
                               int* __restrict__ labels, float* __restrict__ y_true, int B, int T, int L) {
    long long n_ids = (long long)B * T * L;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_ids; i += stride)
        ids[i] = (int)X[i];                                   // .long() truncation, embedding.py:166
    long long n_lab = (long long)B * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_lab; i += stride) {
        int t = (int)(i % T);
        double v = y[i];
        labels[i] = (t == 0) ? 2 : (int)v;                    // token=2 for the target row, RAT_m2.py:115-123
        if (t == 0) y_true[i / T] = (float)v;                 // y.float(), base_model.py:128
    }
}

// ---- K0b: device-resident assembly: pool[retr_indices[i]] with numpy negative-index wraparound ----------
__global__ void k_assemble(const int* __restrict__ q_ids, const unsigned char* __restrict__ q_labels,
                           const long long* __restrict__ rows, long long row0, const int* __restrict__ pool_ids,
                           const unsigned char* __restrict__ pool_labels, const long long* __restrict__ nbr,
                           long long n_pool, int* __restrict__ ids, int* __restrict__ labels,
                           float* __restrict__ y_true, int B, int T, int L, int* __restrict__ err) {
    long long total = (long long)B * T * (L + 1);
    long long stride = (long long)gridDim.x * blockDim.x;
    const int K = T - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int l = (int)(i % (L + 1));
        long long bt = i / (L + 1);
        int t = (int)(bt % T);
        long long b = bt / T;
        long long r = rows ? rows[b] : row0 + b;
        if (t == 0) {
            if (l < L) ids[bt * L + l] = q_ids[r * L + l];
            else { labels[bt] = 2; y_true[b] = (float)q_labels[r]; }
        } else {
            long long j = nbr[r * K + (t - 1)];
            if (j < 0) j += n_pool;                           // numpy fancy-index wrap: -1 -> last pool row
            if (j < 0 || j >= n_pool) { atomicOr(err, 2); j = 0; }
            if (l < L) ids[bt * L + l] = pool_ids[j * L + l];
            else labels[bt] = (int)pool_labels[j];
        }
    }
}

// ---- K1: fused gather -> block [B,T,N,D], x_emb [B,F*D], lr_out [B] -------------------------------------
struct GatherArgs {
    const float* emb_W; const float* lr_W; const float* label_W;
    const int* ids; const int* labels;
    const int* col_off; const int* col_vocab; const int* field_col0; const int* field_width;
    float* block; float* x_emb; float* lr_out;
    int B, T, L, F, D;
    float drop_p; unsigned long long seed; unsigned int stream;
    const unsigned int* step;      // device step counter of the dropout streams (rng_step_ptr)
    int* err;
    // row-sharded tables (range partition, `vs` rows per rank): the flat parameter buffer of every rank is mapped
    // into this process (symmetric memory over NVLink / NVSwitch); peers[o] + emb_off is rank o's [vs, D] shard.
    const float* const* peers; long long emb_off, lr_off; int vs;
};
__device__ __forceinline__ const float* emb_row_ptr(const GatherArgs& a, int r) {
    if (a.peers == nullptr) return a.emb_W + (size_t)r * a.D;
    const int o = r / a.vs;
    return a.peers[o] + a.emb_off + (size_t)(r - o * a.vs) * a.D;
}
__device__ __forceinline__ float lr_value(const GatherArgs& a, int r) {
    if (a.peers == nullptr) return __ldg(a.lr_W + r);
    const int o = r / a.vs;
    return a.peers[o][a.lr_off + (r - o * a.vs)];
}

template <int VW>
__global__ void __launch_bounds__(256) k_gather(GatherArgs a) {
    const int N = a.F + 1, DV = a.D / VW;
    const long long total = (long long)a.B * a.T * N * DV;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv_keep = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
        const int dv = (int)(v % DV);
        const int n = (int)((v / DV) % N);
        const long long bt = v / ((long long)DV * N);
        const int t = (int)(bt % a.T);
        const long long b = bt / a.T;
        float val[VW];
        if (n == 0) {
            int lab = a.labels[bt];
            if (lab < 0 || lab > 2) { atomicOr(a.err, 4); lab = 0; }
            vload<VW>(a.label_W + (long long)lab * a.D + dv * VW, val);
            if (t == 0 && dv == 0 && a.lr_out) {
                // LR_Layer: per field (sequence: sum of its columns) then sum over fields, shallow.py:37-38
                float tot = 0.f;
                for (int f = 0; f < a.F; ++f) {
                    int c0 = a.field_col0[f], w = a.field_width[f];
                    float s = 0.f;
                    for (int j = 0; j < w; ++j) {
                        int id = a.ids[bt * a.L + c0 + j];
                        if (id < 0 || id >= a.col_vocab[c0 + j]) id = 0;
                        s += __ldg(a.lr_W + a.col_off[c0 + j] + id);
                    }
                    tot += s;
                }
                a.lr_out[b] = tot;
            }
        } else {
            const int f = n - 1;
            const int c0 = a.field_col0[f], w = a.field_width[f];
#pragma unroll
            for (int i = 0; i < VW; ++i) val[i] = 0.f;
            for (int j = 0; j < w; ++j) {
                int id = a.ids[bt * a.L + c0 + j];
                if (id < 0 || id >= a.col_vocab[c0 + j]) { atomicOr(a.err, 1); id = 0; }
                float r[VW];
                vload<VW>(a.emb_W + ((long long)a.col_off[c0 + j] + id) * a.D + dv * VW, r);
#pragma unroll
                for (int i = 0; i < VW; ++i) val[i] = (j == 0) ? r[i] : val[i] + r[i];   // left-to-right sum-pool
            }
            if (t == 0 && a.x_emb)                           // X_emb: target row, never dropped out (RAT_m2.py:120)
                vstore<VW>(a.x_emb + (b * a.F + f) * a.D + dv * VW, val);
        }
        if (a.drop_p > 0.f) {
            const unsigned long long e0 = (unsigned long long)v * VW;
#pragma unroll
            for (int i = 0; i < VW; ++i) val[i] *= dropout_scale(a.seed, rng_stream_of_step(a.stream, a.step), e0 + i, a.drop_p, inv_keep);
        }
        vstore<VW>(a.block + v * VW, val);
    }
}

// ---- K1 (fast path): one THREAD per 16-byte vector chunk of the block, fixed chunk geometry per thread ------
// A block owns RPB consecutive (b,t) rows per pass; thread tid is chunk c = tid % NC of row r = tid / NC of the pass
// (NC = (F+1) * D/VW chunks per row), so everything that depends on the chunk only -- token, vector lane, id column,
// table base, vocabulary, sum-pool width -- is computed ONCE and the pass loop is: id load -> bounds check -> one
// 16-byte table-row load -> (sequence fields: w-1 more) -> dropout -> one 16-byte store.  Consecutive threads write
// consecutive chunks (a warp stores 512 contiguous bytes) and U passes are in flight per thread.  The LR logit
// (t == 0 rows only) is done by extra blocks, one warp per target row.  Same arithmetic as k_gather (left-to-right
// sum-pool, same dropout mask function), so the output is bit-identical.
template <int VW, bool SHARDED, int U>
__global__ void __launch_bounds__(320, 4) k_gather_flat(GatherArgs a, int NC, int RPB, int RB, int SROWS, int main_blocks,
                                                     FastDiv divT) {
    pdl_launch_dependents();        // the first attention kernel may stage its weight images under this kernel's tail (common.cuh)
    const int nrows = a.B * a.T;
    if ((int)blockIdx.x >= main_blocks) {
        // ---- LR_Layer (shallow.py:37-38): one warp per target row, lane l owns id column l, sum in field order
        const int lane = threadIdx.x & 31;
        const int wpb = blockDim.x >> 5;
        const int my_off = lane < a.L ? a.col_off[lane] : 0;
        const int my_vocab = lane < a.L ? a.col_vocab[lane] : 1;
        for (int b = ((int)blockIdx.x - main_blocks) * wpb + (threadIdx.x >> 5); b < a.B;
             b += ((int)gridDim.x - main_blocks) * wpb) {
            float lrv = 0.f;
            if (lane < a.L) {
                int id = __ldg(a.ids + (size_t)b * a.T * a.L + lane);
                if (id < 0 || id >= my_vocab) id = 0;               // flagged by the chunk threads
                lrv = SHARDED ? lr_value(a, my_off + id) : __ldg(a.lr_W + my_off + id);
            }
            float tot = 0.f;
            int col = 0;
            for (int f = 0; f < a.F; ++f) {
                const int w = a.field_width[f];
                float s = 0.f;
                for (int j = 0; j < w; ++j, ++col) s += __shfl_sync(0xffffffffu, lrv, col);
                tot += s;
            }
            if (lane == 0) a.lr_out[b] = tot;
        }
        return;
    }
    extern __shared__ int s_ids[];                          // [SROWS][L + 2]: validated ids | label | b if t == 0 else -1
    const int tid = threadIdx.x;
    const int r = tid / NC, c = tid - r * NC;
    const int DV = a.D / VW;
    const int n = c / DV, dv = c - n * DV;
    // chunk geometry (row independent)
    const bool is_label = n == 0;
    const int c0 = is_label ? 0 : a.field_col0[n - 1];
    const int w = is_label ? 1 : a.field_width[n - 1];
    const int row0 = is_label ? 0 : a.col_off[c0];
    const int idcol = is_label ? a.L : c0;
    const int LS = a.L + 2;
    const float* src = (is_label ? a.label_W : (SHARDED ? (const float*)nullptr : a.emb_W + (size_t)row0 * a.D)) + dv * VW;
    float* xdst = (is_label || a.x_emb == nullptr) ? nullptr : a.x_emb + (size_t)(n - 1) * a.D + dv * VW;
    const size_t xstride = (size_t)a.F * a.D;
    const bool drop = a.drop_p > 0.f;
    const float inv_keep = drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    const uint32_t thr = dropout_threshold(a.drop_p);
    const uint32_t key = dropout_key(a.seed, rng_stream_of_step(a.stream, a.step)), hk0 = lowbias32(key);
    const int row_begin = blockIdx.x * RB, row_end = min(nrows, row_begin + RB);

    for (int s0 = row_begin; s0 < row_end; s0 += SROWS) {
        const int srows = min(SROWS, row_end - s0);
        if (s0 != row_begin) __syncthreads();
        // ---- stage + validate this block's ids / labels once (coalesced), so the chunk loop below has ONE global
        //      load on its critical path (the table row)
        for (int i = tid; i < srows * a.L; i += blockDim.x) {
            const int rr = i / a.L, l = i - rr * a.L;
            int id = __ldg(a.ids + (size_t)s0 * a.L + i);
            if (id < 0 || id >= __ldg(a.col_vocab + l)) { atomicOr(a.err, 1); id = 0; }
            s_ids[rr * LS + l] = id;
        }
        for (int i = tid; i < srows; i += blockDim.x) {
            int lab = __ldg(a.labels + s0 + i);
            if (lab < 0 || lab > 2) { atomicOr(a.err, 4); lab = 0; }
            s_ids[i * LS + a.L] = lab;
            const uint32_t bt = (uint32_t)(s0 + i), b = divT.div(bt);
            s_ids[i * LS + a.L + 1] = (bt == b * (uint32_t)a.T) ? (int)b : -1;
        }
        __syncthreads();
        if (r >= RPB) continue;
        for (int g0 = 0; g0 * RPB < srows; g0 += U) {
            int ri[U];
            bool ok[U];
            float val[U][VW];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ri[u] = (g0 + u) * RPB + r;
                ok[u] = ri[u] < srows;
                if (ok[u]) {
                    const int id = s_ids[ri[u] * LS + idcol];
                    if (SHARDED && !is_label) vload<VW>(emb_row_ptr(a, row0 + id) + dv * VW, val[u]);
                    else vload<VW>(src + (size_t)id * a.D, val[u]);
                }
            }
            if (w > 1) {                                    // sequence field: left-to-right sum-pool of its columns
                for (int j = 1; j < w; ++j) {
#pragma unroll
                    for (int h = 0; h < U; h += 2) {        // two rows in flight (keeps the register count at 48)
                        float rv[2][VW];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (h + q >= U || !ok[h + q]) continue;
                            const int idj = s_ids[ri[h + q] * LS + idcol + j];
                            if (SHARDED) vload<VW>(emb_row_ptr(a, row0 + idj) + dv * VW, rv[q]);
                            else vload<VW>(src + (size_t)idj * a.D, rv[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int k = 0; k < VW; ++k)
                                if (h + q < U && ok[h + q]) val[h + q][k] += rv[q][k];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                if (xdst != nullptr) {                      // X_emb: target row, never dropped out (RAT_m2.py:120)
                    const int b = s_ids[ri[u] * LS + a.L + 1];
                    if (b >= 0) vstore<VW>(xdst + (size_t)b * xstride, val[u]);
                }
                const unsigned long long idx = (unsigned long long)(s0 + ri[u]) * NC + c;
                if (drop) dropout_chunk<VW>(val[u], idx, key, hk0, thr, inv_keep);
                vstore<VW>(a.block + idx * VW, val[u]);
            }
        }
    }
}

// in-place dropout backward on the block gradient (same mask as the gather); one Philox call per 8 elements
__global__ void k_dropout_bwd(float* __restrict__ g, long long n, float p, unsigned long long seed,
                              unsigned int stream0, const unsigned int* __restrict__ step) {
    const unsigned int stream = rng_stream_of_step(stream0, step);
    const float inv_keep = 1.0f / (1.0f - p);
    const uint32_t thr = dropout_threshold(p);
    const long long n8 = (n + 7) >> 3;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n8; c += stride) {
        const uint4 bits = dropout_bits8(seed, stream, (unsigned long long)c);
        const long long e0 = c << 3;
        if (e0 + 8 <= n && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            float4 v0 = *reinterpret_cast<float4*>(g + e0), v1 = *reinterpret_cast<float4*>(g + e0 + 4);
            v0.x *= dropout_lane16(bits, 0) < thr ? 0.f : inv_keep; v0.y *= dropout_lane16(bits, 1) < thr ? 0.f : inv_keep;
            v0.z *= dropout_lane16(bits, 2) < thr ? 0.f : inv_keep; v0.w *= dropout_lane16(bits, 3) < thr ? 0.f : inv_keep;
            v1.x *= dropout_lane16(bits, 4) < thr ? 0.f : inv_keep; v1.y *= dropout_lane16(bits, 5) < thr ? 0.f : inv_keep;
            v1.z *= dropout_lane16(bits, 6) < thr ? 0.f : inv_keep; v1.w *= dropout_lane16(bits, 7) < thr ? 0.f : inv_keep;
            *reinterpret_cast<float4*>(g + e0) = v0; *reinterpret_cast<float4*>(g + e0 + 4) = v1;
        } else {
            for (int k = 0; k < 8 && e0 + k < n; ++k) g[e0 + k] *= dropout_lane16(bits, k) < thr ? 0.f : inv_keep;
        }
    }
}

// dst[r*dst_stride + d] = src[r*src_stride + d], d < D  (token-0 pooling of RAT_m1 and its backward scatter)
__global__ void k_strided_copy(const float* __restrict__ src, float* __restrict__ dst, long long rows, int D,
                               long long src_stride, long long dst_stride) {
    const long long total = rows * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D;
        const int d = (int)(i % D);
        dst[r * dst_stride + d] = src[r * src_stride + d];
    }
}

// dst[(r*group + g)*D + d] = g == 0 ? src[r*D + d] : 0   (vectorised when D % 4 == 0): the gradient of the field-token-0 rows of
// the last RAT block scattered back into the full [B,T,N,D] block gradient, which it also zero-fills (no separate memset)
template <int VW>
__global__ void k_expand_rows(const float* __restrict__ src, float* __restrict__ dst, long long rows, int DV, int group) {
    pdl_launch_dependents();
    const long long total = rows * group * DV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / DV;
        const int dv = (int)(i - row * DV);
        const long long r = row / group;
        float v[VW];
#pragma unroll
        for (int k = 0; k < VW; ++k) v[k] = 0.f;
        if (row - r * group == 0) vload<VW>(src + (r * DV + dv) * VW, v);
        vstore<VW>(dst + i * VW, v);
    }
}

static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    long long cap = (long long)num_sms() * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace rat

using namespace rat;

extern "C" int rat_convert_wire_f64(const double* X, const double* y, int* ids, int* labels, float* y_true, int B,
                                    int T, int L, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0, "rat_convert_wire_f64: bad shape B=%d T=%d L=%d", B, T, L);
    k_convert_wire<<<grid_for((long long)B * T * L, 256), 256, 0, (cudaStream_t)stream>>>(X, y, ids, labels, y_true,
                                                                                          B, T, L);
    RAT_CHECK_LAUNCH("k_convert_wire");
    return RAT_OK;
}

extern "C" int rat_assemble_ids(const int* q_ids, const unsigned char* q_labels, const long long* rows,
                                long long row0, const int* pool_ids, const unsigned char* pool_labels,
                                const long long* nbr, long long n_pool, int* ids, int* labels, float* y_true, int B,
                                int T, int L, int* err_flag, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0 && n_pool > 0, "rat_assemble_ids: bad shape");
    k_assemble<<<grid_for((long long)B * T * (L + 1), 256), 256, 0, (cudaStream_t)stream>>>(
        q_ids, q_labels, rows, row0, pool_ids, pool_labels, nbr, n_pool, ids, labels, y_true, B, T, L, err_flag);
    RAT_CHECK_LAUNCH("k_assemble");
    return RAT_OK;
}

static int launch_gather(GatherArgs a, cudaStream_t st) {
    const int B = a.B, T = a.T, L = a.L, F = a.F, D = a.D;
    const int vw = (D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1;
    long long total = (long long)B * T * (F + 1) * (D / vw);
    int grid = grid_for(total, 256);
    const int nc = (F + 1) * (D / vw);                    // vector chunks per (b,t) row
    if (nc <= 320 && L <= 32 && (long long)B * T < (1ll << 30)) {       // thread-per-chunk fast path
        const int nrows = B * T;
        int rpb = 1, best = 0;                                          // rows per pass: fill the block's warps
        for (int r = 1; r * nc <= 320 && r <= 16; ++r) {
            const int thr = round_up(r * nc, 32);
            const int util = r * nc * 1000 / thr;
            if (util > best + 10 || util >= best) { best = std::max(best, util); rpb = r; }
        }
        const int threads = round_up(rpb * nc, 32);
        const FastDiv dT = make_fastdiv((uint32_t)T);
        const bool sh = a.peers != nullptr;
        const int lr_blocks = a.lr_out ? std::min((B + threads / 32 - 1) / (threads / 32), num_sms()) : 0;
        const int unit = rpb * 4;                                       // rows of one unrolled trip (U = 4)
        // each block owns RB consecutive rows (one resident wave), staged SROWS rows at a time
#define RAT_GATHER_FLAT(VW_, SH_) do {                                                                                 \
            int per_sm = 1;                                                                                            \
            const int srows_max = std::max(unit, 128 / unit * unit);                                                   \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather_flat<VW_, SH_, 4>, threads,                \
                                                          (size_t)srows_max * (L + 2) * 4);                            \
            const int want = num_sms() * std::max(1, per_sm);                                                          \
            const int rb = round_up((nrows + want - 1) / want, rpb);                                                   \
            const int main_blocks = (nrows + rb - 1) / rb;                                                             \
            const int srows = std::min(srows_max, round_up(rb, unit));                                                 \
            k_gather_flat<VW_, SH_, 4><<<main_blocks + lr_blocks, threads, (size_t)srows * (L + 2) * 4, st>>>(         \
                a, nc, rpb, rb, srows, main_blocks, dT);                                                               \
        } while (0)
        if (vw == 4) { if (sh) RAT_GATHER_FLAT(4, true); else RAT_GATHER_FLAT(4, false); }
        else if (vw == 2) { if (sh) RAT_GATHER_FLAT(2, true); else RAT_GATHER_FLAT(2, false); }
        else { if (sh) RAT_GATHER_FLAT(1, true); else RAT_GATHER_FLAT(1, false); }
#undef RAT_GATHER_FLAT
        RAT_CHECK_LAUNCH("k_gather_flat");
        return RAT_OK;
    }
    RAT_REQUIRE(a.peers == nullptr, "rat_gather_fwd_sharded: shape outside the thread-per-chunk path (L=%d <= 32, <= 320 "
                                    "vector chunks per row)", L);
    if (vw == 4) k_gather<4><<<grid, 256, 0, st>>>(a);
    else if (vw == 2) k_gather<2><<<grid, 256, 0, st>>>(a);
    else k_gather<1><<<grid, 256, 0, st>>>(a);
    RAT_CHECK_LAUNCH("k_gather");
    return RAT_OK;
}

extern "C" int rat_gather_fwd(const float* emb_W, const float* lr_W, const float* label_W, const int* ids,
                              const int* labels, const int* col_off, const int* col_vocab, const int* field_col0,
                              const int* field_width, float* block, float* x_emb, float* lr_out, int B, int T, int L,
                              int F, int D, float drop_p, unsigned long long seed, unsigned int rng_stream,
                              int* err_flag, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0 && F > 0 && D > 0, "rat_gather_fwd: bad shape");
    RAT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "rat_gather_fwd: dropout p=%f", drop_p);
    GatherArgs a{emb_W, lr_W, label_W, ids, labels, col_off, col_vocab, field_col0, field_width,
                 block, x_emb, lr_out, B, T, L, F, D, drop_p, seed, rng_stream, rng_step_ptr(), err_flag, nullptr, 0, 0, 1};
    return launch_gather(a, (cudaStream_t)stream);
}

extern "C" int rat_gather_fwd_sharded(const float* const* W_peers, long long emb_off, long long lr_off,
                                      int rows_per_shard, int world, const float* label_W, const int* ids,
                                      const int* labels, const int* col_off, const int* col_vocab,
                                      const int* field_col0, const int* field_width, float* block, float* x_emb,
                                      float* lr_out, int B, int T, int L, int F, int D, float drop_p,
                                      unsigned long long seed, unsigned int rng_stream, int* err_flag, void* stream) {
    RAT_REQUIRE(B > 0 && T > 0 && L > 0 && F > 0 && D > 0, "rat_gather_fwd_sharded: bad shape");
    RAT_REQUIRE(W_peers != nullptr && rows_per_shard > 0 && world > 0, "rat_gather_fwd_sharded: bad shard description");
    RAT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "rat_gather_fwd_sharded: dropout p=%f", drop_p);
    GatherArgs a{nullptr, nullptr, label_W, ids, labels, col_off, col_vocab,
                 field_col0, field_width, block, x_emb, lr_out, B, T, L, F, D, drop_p, seed, rng_stream, rng_step_ptr(), err_flag,
                 W_peers, emb_off, lr_off, rows_per_shard};
    return launch_gather(a, (cudaStream_t)stream);
}

extern "C" int rat_dropout_bwd(float* grad, long long n, float p, unsigned long long seed, unsigned int rng_stream,
                               void* stream) {
    if (p <= 0.f) return RAT_OK;
    k_dropout_bwd<<<grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(grad, n, p, seed, rng_stream, rng_step_ptr());
    RAT_CHECK_LAUNCH("k_dropout_bwd");
    return RAT_OK;
}

extern "C" int rat_expand_rows(const float* src, float* dst, long long rows, int D, int group, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0 && group > 0, "rat_expand_rows: bad shape");
    const bool v4 = (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    if (v4) k_expand_rows<4><<<grid_for(rows * group * (D / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, D / 4, group);
    else k_expand_rows<1><<<grid_for(rows * group * D, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, D, group);
    RAT_CHECK_LAUNCH("k_expand_rows");
    return RAT_OK;
}

extern "C" int rat_strided_copy(const float* src, float* dst, long long rows, int D, long long src_stride,
                                long long dst_stride, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0, "rat_strided_copy: bad shape");
    k_strided_copy<<<grid_for(rows * D, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, D, src_stride, dst_stride);
    RAT_CHECK_LAUNCH("k_strided_copy");
    return RAT_OK;
}
