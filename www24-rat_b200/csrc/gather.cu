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

struct GatherArgs {
    const float* emb_W; const float* lr_W; const float* label_W;
    const int* ids; const int* labels;
    const int* col_off; const int* col_vocab; const int* field_col0; const int* field_width;
    float* block; float* x_emb; float* lr_out;
    int B, T, L, F, D;
    float drop_p; unsigned long long seed; unsigned int stream;
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
            for (int i = 0; i < VW; ++i) val[i] *= dropout_scale(a.seed, a.stream, e0 + i, a.drop_p, inv_keep);
        }
        vstore<VW>(a.block + v * VW, val);
    }
}

// ---- K1 (fast path): one WARP per (b,t) row of the block ---------------------------------------------------
// Lane l < L owns column l of the id matrix: it validates id[bt][l] once and keeps the table row in a register;
// every lane then produces PAIRS of adjacent vector chunks (2*VW contiguous floats = 32 B for VW = 4) of the row's
// N*D contiguous output floats.  Table rows travel by shuffle, so all row loads of a lane are INDEPENDENT and the
// block row is written as one contiguous, fully coalesced stream.  Extra sum-pool terms (sequence fields) are only
// visited by the pair slots that contain a sequence token (warp-uniform test).  With VW = 4 a chunk pair is exactly
// the eight elements of one Philox call (dropout).  Same arithmetic as k_gather (left-to-right sum-pool, same mask).
template <int VW, int PS>
__global__ void __launch_bounds__(256, 3) k_gather_rows(GatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int N = a.F + 1, DV = a.D / VW, NC = N * DV;
    const int nrows = a.B * a.T;
    const int wstride = gridDim.x * (blockDim.x >> 5);
    const float inv_keep = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    const uint32_t thr = dropout_threshold(a.drop_p);
    const int my_off = lane < a.L ? a.col_off[lane] : 0;
    const int my_vocab = lane < a.L ? a.col_vocab[lane] : 1;
    // per-lane chunk geometry is row independent: token n, vector dv, first column c0 and width w of its field
    int cn[PS][2], cdv[PS][2], cc0[PS][2], cw[PS][2], slotw[PS];
#pragma unroll
    for (int i = 0; i < PS; ++i) {
        int mw = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * (lane + 32 * i) + h;
            cn[i][h] = c < NC ? c / DV : -1;
            cdv[i][h] = c < NC ? c - cn[i][h] * DV : 0;
            cc0[i][h] = 0; cw[i][h] = 0;
            if (cn[i][h] > 0) { cc0[i][h] = a.field_col0[cn[i][h] - 1]; cw[i][h] = a.field_width[cn[i][h] - 1]; }
            mw = max(mw, cw[i][h]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mw = max(mw, __shfl_xor_sync(0xffffffffu, mw, o));
        slotw[i] = mw;                                           // warp-uniform: widest field touched by pair slot i
    }
    int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int t = bt % a.T, b = bt / a.T;
    const int dt = wstride % a.T, db = wstride / a.T;
    for (; bt < nrows; bt += wstride) {
        int id = 0;
        if (lane < a.L) {
            id = __ldg(a.ids + (size_t)bt * a.L + lane);
            if (id < 0 || id >= my_vocab) { atomicOr(a.err, 1); id = 0; }
        }
        const int row = my_off + id;                         // table row of column `lane`
        int lab = __ldg(a.labels + bt);
        if (lab < 0 || lab > 2) { if (lane == 0) atomicOr(a.err, 4); lab = 0; }
        if (t == 0 && a.lr_out) {                            // LR_Layer (shallow.py:37-38), field order, lane 0
            const float lrv = lane < a.L ? lr_value(a, row) : 0.f;
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
        float val[PS][2][VW];
#pragma unroll
        for (int i = 0; i < PS; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int k = 0; k < VW; ++k) val[i][h][k] = 0.f;
                if (cn[i][h] == 0) vload<VW>(a.label_W + (size_t)lab * a.D + cdv[i][h] * VW, val[i][h]);
                const int r = __shfl_sync(0xffffffffu, row, min(cc0[i][h], 31));
                if (cw[i][h] > 0) vload<VW>(emb_row_ptr(a, r) + cdv[i][h] * VW, val[i][h]);
            }
#pragma unroll
        for (int i = 0; i < PS; ++i)
            for (int j = 1; j < slotw[i]; ++j)               // sum-pool terms 1.. of sequence fields (rare slots only)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = __shfl_sync(0xffffffffu, row, min(cc0[i][h] + j, 31));
                    if (j < cw[i][h]) {
                        float rv[VW];
                        vload<VW>(emb_row_ptr(a, r) + cdv[i][h] * VW, rv);
#pragma unroll
                        for (int k = 0; k < VW; ++k) val[i][h][k] += rv[k];
                    }
                }
        float* out_row = a.block + (size_t)bt * NC * VW;
#pragma unroll
        for (int i = 0; i < PS; ++i) {
            const int c0 = 2 * (lane + 32 * i);
            uint4 bits = make_uint4(0u, 0u, 0u, 0u);
            const unsigned long long e0 = ((unsigned long long)bt * NC + c0) * VW;
            if (a.drop_p > 0.f && cn[i][0] >= 0) bits = dropout_bits8(a.seed, a.stream, e0 >> 3);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (cn[i][h] < 0) continue;
                if (t == 0 && a.x_emb && cn[i][h] > 0)
                    vstore<VW>(a.x_emb + ((size_t)b * a.F + (cn[i][h] - 1)) * a.D + cdv[i][h] * VW, val[i][h]);
                if (a.drop_p > 0.f) {
                    const unsigned long long eh = e0 + (unsigned long long)h * VW;
                    uint4 bh = bits;
                    if ((eh >> 3) != (e0 >> 3)) bh = dropout_bits8(a.seed, a.stream, eh >> 3);   // never for VW=4, even NC
#pragma unroll
                    for (int k = 0; k < VW; ++k)
                        val[i][h][k] *= dropout_lane16(bh, (int)((eh + k) & 7)) < thr ? 0.0f : inv_keep;
                }
                vstore<VW>(out_row + (size_t)(c0 + h) * VW, val[i][h]);
            }
        }
        t += dt; b += db;
        if (t >= a.T) { t -= a.T; ++b; }
    }
}

// in-place dropout backward on the block gradient (same mask as the gather); one Philox call per 8 elements
__global__ void k_dropout_bwd(float* __restrict__ g, long long n, float p, unsigned long long seed,
                              unsigned int stream) {
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
    const int ps = (nc + 63) / 64;                        // chunk-pair slots per lane
    if (L <= 32 && ps <= 4 && vw >= 2 && (long long)B * T < (1ll << 30)) {       // warp-per-row fast path
        const long long nrows = (long long)B * T;
        const int rgrid = (int)std::min<long long>((nrows + 7) / 8, (long long)num_sms() * 8);
#define RAT_GATHER_ROWS(VW_, PS_) k_gather_rows<VW_, PS_><<<rgrid, 256, 0, st>>>(a)
        if (vw == 4) {
            if (ps <= 1) RAT_GATHER_ROWS(4, 1); else if (ps <= 2) RAT_GATHER_ROWS(4, 2); else if (ps <= 3) RAT_GATHER_ROWS(4, 3); else RAT_GATHER_ROWS(4, 4);
        } else {
            if (ps <= 1) RAT_GATHER_ROWS(2, 1); else if (ps <= 2) RAT_GATHER_ROWS(2, 2); else if (ps <= 3) RAT_GATHER_ROWS(2, 3); else RAT_GATHER_ROWS(2, 4);
        }
#undef RAT_GATHER_ROWS
        RAT_CHECK_LAUNCH("k_gather_rows");
        return RAT_OK;
    }
    RAT_REQUIRE(a.peers == nullptr, "rat_gather_fwd_sharded: shape outside the warp-per-row path (L=%d <= 32, even D, <= 512 "
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
                 block, x_emb, lr_out, B, T, L, F, D, drop_p, seed, rng_stream, err_flag, nullptr, 0, 0, 1};
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
                 field_col0, field_width, block, x_emb, lr_out, B, T, L, F, D, drop_p, seed, rng_stream, err_flag,
                 W_peers, emb_off, lr_off, rows_per_shard};
    return launch_gather(a, (cudaStream_t)stream);
}

extern "C" int rat_dropout_bwd(float* grad, long long n, float p, unsigned long long seed, unsigned int rng_stream,
                               void* stream) {
    if (p <= 0.f) return RAT_OK;
    k_dropout_bwd<<<grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(grad, n, p, seed, rng_stream);
    RAT_CHECK_LAUNCH("k_dropout_bwd");
    return RAT_OK;
}

extern "C" int rat_strided_copy(const float* src, float* dst, long long rows, int D, long long src_stride,
                                long long dst_stride, void* stream) {
    RAT_REQUIRE(rows > 0 && D > 0, "rat_strided_copy: bad shape");
    k_strided_copy<<<grid_for(rows * D, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, D, src_stride, dst_stride);
    RAT_CHECK_LAUNCH("k_strided_copy");
    return RAT_OK;
}
