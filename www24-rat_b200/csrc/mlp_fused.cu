// K3/K4, single-launch BatchNorm / bias-gradient / head-gradient kernels of the DNN head (layers/deep.py:108-141,
// RAT_m2.py:144-150).  Round 2: the DNN head was 50 of the step's 109 launches (0.63 ms of 3.86 ms for 1 % of the FLOPs);
// every BatchNorm was "column partial sums -> finalize -> mean/rstd -> elementwise apply" in 4-5 dependent launches of a few
// microseconds each.  A column statistic needs every row, so a fusion needs a cross-CTA exchange: here a THREAD-BLOCK
// CLUSTER of 8 CTAs owns a 32-column slab (CTA y = row slice y), the slab partials are exchanged through DISTRIBUTED
// SHARED MEMORY (cluster.map_shared_rank) and summed in rank order by every CTA (bitwise run-to-run deterministic), and the
// same CTAs then apply the normalisation to their rows -- the second read of the slab hits L2 (a [4096, 400] fp32
// activation is 6.5 MB).  Grid = (C / 32 slabs) x 8 = 104 CTAs for C = 400.
//
//   k_bn_act_fwd_cl   z -> batch mean / rstd (saved for the backward), running-stat update, out = dropout(relu(bn(z)))
//   k_bn_act_bwd_cl   dout, out, z -> dz = d relu/bn, dgamma, dbeta, bias gradient colsum(dz), max|dz|
//   k_head_bwd_cl     dlogit -> fc.weight / fc.bias gradient, final Linear(K -> 1) weight / bias gradient, d h_last
//
// Data-parallel training (SyncBN-equivalent statistics over the GLOBAL batch): the cluster leader of every slab exchanges
// its 2 x 32 double totals with the same slab's leader on every other rank through NVLink peer memory (one-shot: store
// into every peer's slot, release a flag, acquire the peers' flags, sum the slots in rank order -- the protocol of
// collective.cu, one independent channel per slab) between the two passes, inside the same launch.
#include <algorithm>
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace cg = cooperative_groups;

namespace rat {

constexpr int CL = 8;              // CTAs per cluster = row slices of a column slab
constexpr int CL_RG = 16;          // row groups per CTA (16 warps, lane = column of the slab)
constexpr int CL_THREADS = CL_RG * 32;

struct ClRed {
    double s[2][CL_RG][32];        // per-warp partials of this CTA
    double tot[2][2][32];          // [round][quantity][column]: this CTA's slab partials, read by the whole cluster
    double all[2][32];             // cluster totals
};

// Sum (a, b) over every thread of the cluster that has the same lane (= column).  On return R.all[q][lane] holds the totals
// (identical bits in every CTA: rank-order sum).  `round` selects the exchange buffer so that a second reduction does not
// overwrite partials a slower CTA is still reading; the caller ends the kernel with cluster.sync().
__device__ __forceinline__ void cluster_colsum(cg::cluster_group& cl, ClRed& R, int round, double a, double b, int rg, int lane) {
    R.s[0][rg][lane] = a;
    R.s[1][rg][lane] = b;
    __syncthreads();
    if (rg < 2) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < CL_RG; ++i) t += R.s[rg][i][lane];
        R.tot[round][rg][lane] = t;
    }
    cl.sync();
    if (rg < 2) {
        double t = 0.0;
        for (unsigned int r = 0; r < CL; ++r) t += *cl.map_shared_rank(&R.tot[round][rg][lane], r);
        R.all[rg][lane] = t;
    }
    __syncthreads();
}

// ---- cross-rank exchange of a slab's totals (data-parallel BatchNorm) -------------------------------------------------
// Symmetric buffer of a rank (rat_bn_exchange_workspace_bytes, zero-initialised once): BNX_CHANNELS channels of
//   [0] uint32 seq | [64 ..] uint32 flags[2][64] | [1024 ..] double slots[2][world][64]
// Channel = slab index; each channel is the one-shot protocol of collective.cu on its own counters, so the 13 slab leaders
// of a launch never touch each other's state and consecutive launches alternate slot sets by call parity.
constexpr int BNX_CHANNELS = 64;
constexpr int BNX_HDR = 1024;
__host__ __device__ inline size_t bnx_channel_bytes(int world) { return (size_t)BNX_HDR + (size_t)2 * world * 64 * sizeof(double); }
struct BnExchange { unsigned char* const* peers; int rank, world; };

__device__ __forceinline__ void bnx_st_release(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int bnx_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// On entry R.all holds this rank's totals (every CTA of the cluster); on return it holds the sum over all ranks (rank-order
// association: identical bits on every rank).  Called by all threads of all CTAs of the cluster.
__device__ __forceinline__ void cluster_exchange(cg::cluster_group& cl, ClRed& R, const BnExchange& x, int channel, int rg, int lane) {
    __shared__ unsigned int seq_s;
    if (cl.block_rank() == 0) {
        const size_t chb = bnx_channel_bytes(x.world);
        unsigned char* mine = x.peers[x.rank] + (size_t)channel * chb;
        if (threadIdx.x == 0) {
            unsigned int* seqp = reinterpret_cast<unsigned int*>(mine);
            seq_s = *seqp + 1u;
            *seqp = seq_s;
        }
        __syncthreads();
        const unsigned int seq = seq_s, par = seq & 1u;
        if (rg < 2) {
            const double v = R.all[rg][lane];
            const size_t off = (size_t)channel * chb + BNX_HDR + (((size_t)par * x.world + x.rank) * 64 + rg * 32 + lane) * sizeof(double);
            for (int p = 0; p < x.world; ++p) *reinterpret_cast<double*>(x.peers[p] + off) = v;
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < x.world) {
            bnx_st_release(reinterpret_cast<unsigned int*>(x.peers[threadIdx.x] + (size_t)channel * chb + 64) + par * 64 + x.rank, seq);
            const unsigned int* f = reinterpret_cast<const unsigned int*>(mine + 64) + par * 64 + threadIdx.x;
            unsigned int it = 0;
            while (bnx_ld_acquire(f) != seq) {
                if (++it > (1u << 28)) __trap();      // a peer never arrived: surface an error instead of hanging the GPU
            }
        }
        __syncthreads();
        if (rg < 2) {
            const double* slots = reinterpret_cast<const double*>(mine + BNX_HDR) + (size_t)par * x.world * 64 + rg * 32 + lane;
            double t = 0.0;
            for (int r = 0; r < x.world; ++r) t += slots[(size_t)r * 64];
            R.all[rg][lane] = t;
        }
    }
    cl.sync();                                        // the leader's R.all holds the global totals
    if (cl.block_rank() != 0 && rg < 2) R.all[rg][lane] = *cl.map_shared_rank(&R.all[rg][lane], 0);
    __syncthreads();
}

// NV > 0: the CTA's rows fit NV per thread (rows <= CL * CL_RG * NV): every load of the slab is issued before the first use
// and the second pass runs from registers (one read of z).  NV == 0: generic two-pass loop (the second read hits L2).
template <int NV>
__global__ void __cluster_dims__(1, CL, 1) __launch_bounds__(CL_THREADS)
k_bn_act_fwd_cl(const float* __restrict__ z, int Bn, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ mean_o, float* __restrict__ rstd_o, float* __restrict__ running_mean,
                float* __restrict__ running_var, float momentum, float eps, float* __restrict__ out, float drop_p,
                unsigned long long seed, unsigned int stream0, const unsigned int* __restrict__ step, BnExchange xc) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ ClRed R;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < C;
    const int per = (Bn + CL - 1) / CL, r0 = blockIdx.y * per, r1 = min(Bn, r0 + per);
    double a = 0.0, b = 0.0;
    float zc[NV > 0 ? NV : 1];
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int r = r0 + rg + j * CL_RG;
            zc[j] = (ok && r < r1) ? z[(size_t)r * C + c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) { const double v = zc[j]; a += v; b += v * v; }      // rows past r1 add exact zeros
    } else if (ok) {
#pragma unroll 8
        for (int r = r0 + rg; r < r1; r += CL_RG) { const double v = z[(size_t)r * C + c]; a += v; b += v * v; }
    }
    cluster_colsum(cl, R, 0, a, b, rg, lane);
    if (xc.world > 1) cluster_exchange(cl, R, xc, blockIdx.x, rg, lane);
    if (ok) {
        const double count = (double)Bn * (double)xc.world;
        const double m = R.all[0][lane] / count;
        double var = R.all[1][lane] / count - m * m;
        if (var < 0.0) var = 0.0;
        const float mu = (float)m, rs = (float)(1.0 / sqrt(var + (double)eps));
        if (blockIdx.y == 0 && rg == 0) {
            mean_o[c] = mu;
            rstd_o[c] = rs;
            if (running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        }
        const unsigned int stream = rng_stream_of_step(stream0, step);
        const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
        const float g = gamma[c], bt = beta[c];
        if (NV > 0) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int r = r0 + rg + j * CL_RG;
                if (r >= r1) continue;
                const size_t i = (size_t)r * C + c;
                float v = (zc[j] - mu) * rs * g + bt;
                v = fmaxf(v, 0.f);
                if (drop_p > 0.f) v *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
                out[i] = v;
            }
        } else {
#pragma unroll 8
            for (int r = r0 + rg; r < r1; r += CL_RG) {
                const size_t i = (size_t)r * C + c;
                float v = (z[i] - mu) * rs * g + bt;
                v = fmaxf(v, 0.f);
                if (drop_p > 0.f) v *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
                out[i] = v;
            }
        }
    }
    cl.sync();                      // no CTA may exit while a peer still reads its partials
}

template <int NV>
__global__ void __cluster_dims__(1, CL, 1) __launch_bounds__(CL_THREADS)
k_bn_act_bwd_cl(const float* dout, const float* __restrict__ out, const float* __restrict__ z,
                const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma, int Bn, int C,
                float* dz, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, float drop_p,
                unsigned long long seed, unsigned int stream0, const unsigned int* __restrict__ step,
                float* __restrict__ dz_amax, BnExchange xc) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ ClRed R;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < C;
    const int per = (Bn + CL - 1) / CL, r0 = blockIdx.y * per, r1 = min(Bn, r0 + per);
    const unsigned int stream = rng_stream_of_step(stream0, step);
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const bool bn = mean != nullptr;
    const float mu = (bn && ok) ? mean[c] : 0.f, rs = (bn && ok) ? rstd[c] : 1.f;
    float db = 0.f, dg = 0.f, gr = 1.f;
    float dyc[NV > 0 ? NV : 1], xhc[NV > 0 ? NV : 1];        // dy = d relu / dropout, xhat (register-resident slab)
    if (NV > 0) {
        float oc[NV > 0 ? NV : 1];
#pragma unroll
        for (int j = 0; j < NV; ++j) {                       // all loads of the slab in flight at once
            const int r = r0 + rg + j * CL_RG;
            const bool in = ok && r < r1;
            const size_t i = (size_t)r * C + c;
            oc[j] = in ? out[i] : 0.f;
            dyc[j] = in ? dout[i] : 0.f;
            xhc[j] = (in && bn) ? z[i] : mu;
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int r = r0 + rg + j * CL_RG;
            float dy = oc[j] > 0.f ? dyc[j] : 0.f;
            if (drop_p > 0.f && ok && r < r1) dy *= dropout_scale(seed, stream, (unsigned long long)((size_t)r * C + c), drop_p, inv_keep);
            dyc[j] = dy;
            xhc[j] = (xhc[j] - mu) * rs;
        }
    }
    if (bn) {
        double a = 0.0, b = 0.0;
        if (NV > 0) {
#pragma unroll
            for (int j = 0; j < NV; ++j) { a += dyc[j]; b += (double)dyc[j] * (double)xhc[j]; }     // rows past r1: dy = 0
        } else if (ok) {
#pragma unroll 4
            for (int r = r0 + rg; r < r1; r += CL_RG) {
                const size_t i = (size_t)r * C + c;
                float dy = out[i] > 0.f ? dout[i] : 0.f;
                if (drop_p > 0.f) dy *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
                a += dy;
                b += (double)dy * (double)((z[i] - mu) * rs);
            }
        }
        cluster_colsum(cl, R, 0, a, b, rg, lane);
        if (xc.world > 1) cluster_exchange(cl, R, xc, blockIdx.x, rg, lane);
        if (ok) {
            // data-parallel: the sums are GLOBAL, so every rank writes 1/world of dgamma / dbeta and the later SUM
            // all-reduce of the gradient buffer restores them (as rat_bn_act_bwd_apply's param_grad_scale)
            const double count = (double)Bn * (double)xc.world;
            const float pgs = 1.0f / (float)xc.world;
            db = (float)(R.all[0][lane] / count);
            dg = (float)(R.all[1][lane] / count);
            gr = gamma[c] * rs;
            if (blockIdx.y == 0 && rg == 0) { dgamma[c] = (float)R.all[1][lane] * pgs; dbeta[c] = (float)R.all[0][lane] * pgs; }
        }
    }
    float amax = 0.f;
    double sb = 0.0;
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int r = r0 + rg + j * CL_RG;
            if (!ok || r >= r1) continue;
            float rr = dyc[j];
            if (bn) rr = gr * (dyc[j] - db - xhc[j] * dg);
            dz[(size_t)r * C + c] = rr;
            amax = fmaxf(amax, fabsf(rr));
            sb += rr;
        }
    } else if (ok) {
#pragma unroll 4
        for (int r = r0 + rg; r < r1; r += CL_RG) {
            const size_t i = (size_t)r * C + c;
            float dy = out[i] > 0.f ? dout[i] : 0.f;
            if (drop_p > 0.f) dy *= dropout_scale(seed, stream, (unsigned long long)i, drop_p, inv_keep);
            float rr = dy;
            if (bn) {
                const float xh = (z[i] - mu) * rs;
                rr = gr * (dy - db - xh * dg);
            }
            dz[i] = rr;
            amax = fmaxf(amax, fabsf(rr));
            sb += rr;
        }
    }
    __syncthreads();                // R.s / R.all of the first reduction have been consumed by every thread
    cluster_colsum(cl, R, 1, sb, 0.0, rg, lane);
    if (ok && dbias != nullptr && blockIdx.y == 0 && rg == 0) dbias[c] = (float)R.all[0][lane];
    publish_amax_block(dz_amax, amax);
    cl.sync();
}

// Gradients that hang off dlogit [Bn] (d BCE / d logit, written by rat_head):
//   slabs [0, slabs_h):            g_final_w[c] = sum_r dlogit[r] * h_last[r][c],  dh_last[r][c] = dlogit[r] * w_final[c]
//   slabs [slabs_h, +slabs_e):     g_fc_w[d]   = sum_r dlogit[r] * enc[r * enc_stride + d]
//   last slab:                     g_fc_b = g_final_b = sum_r dlogit[r]
__global__ void __cluster_dims__(1, CL, 1) __launch_bounds__(CL_THREADS)
k_head_bwd_cl(const float* __restrict__ dlogit, int Bn, const float* __restrict__ enc, long long enc_stride, int D,
              float* __restrict__ g_fc_w, float* __restrict__ g_fc_b, const float* __restrict__ h_last, int K,
              const float* __restrict__ w_final, float* __restrict__ g_final_w, float* __restrict__ g_final_b,
              float* __restrict__ dh_last, int slabs_h, int slabs_e) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ ClRed R;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int per = (Bn + CL - 1) / CL, r0 = blockIdx.y * per, r1 = min(Bn, r0 + per);
    const int slab = blockIdx.x;
    double a = 0.0;
    float* dst = nullptr;
    float* dst2 = nullptr;
    if (slab < slabs_h) {
        const int c = slab * 32 + lane;
        if (c < K) {
            const float w = w_final[c];
#pragma unroll 4
            for (int r = r0 + rg; r < r1; r += CL_RG) {
                const float g = dlogit[r];
                a += (double)(g * h_last[(size_t)r * K + c]);
                dh_last[(size_t)r * K + c] = g * w;
            }
            dst = g_final_w + c;
        }
    } else if (slab < slabs_h + slabs_e) {
        const int c = (slab - slabs_h) * 32 + lane;
        if (c < D) {
#pragma unroll 4
            for (int r = r0 + rg; r < r1; r += CL_RG) a += (double)(dlogit[r] * enc[(size_t)r * enc_stride + c]);
            dst = g_fc_w + c;
        }
    } else {
        for (int r = r0 + rg * 32 + lane; r < r1; r += CL_THREADS) a += (double)dlogit[r];
    }
    cluster_colsum(cl, R, 0, a, 0.0, rg, lane);
    if (blockIdx.y == 0 && rg == 0) {
        double t = R.all[0][lane];
        if (slab >= slabs_h + slabs_e) {
            t = warp_sum_d(t);
            if (lane == 0) { dst = g_fc_b; dst2 = g_final_b; }
        }
        if (dst) *dst = (float)t;
        if (dst2) *dst2 = (float)t;
    }
    cl.sync();
}

}  // namespace rat

using namespace rat;

extern "C" size_t rat_bn_exchange_workspace_bytes(int world) { return (size_t)BNX_CHANNELS * bnx_channel_bytes(world < 1 ? 1 : world); }

static int bnx_check(const void* const* peer_bufs, int rank, int world, int C, const char* who) {
    RAT_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "%s: bad rank %d / world %d", who, rank, world);
    RAT_REQUIRE(world == 1 || peer_bufs != nullptr, "%s: world > 1 needs the peers' exchange buffers", who);
    RAT_REQUIRE(world == 1 || ceil_div(C, 32) <= BNX_CHANNELS, "%s: C=%d exceeds %d exchange channels", who, C, BNX_CHANNELS);
    return RAT_OK;
}

extern "C" int rat_bn_act_fwd_train(const float* z, int rows, int C, const float* gamma, const float* beta, float* mean,
                                    float* rstd, float* running_mean, float* running_var, float momentum, float eps,
                                    float* out, float drop_p, unsigned long long seed, unsigned int rng_stream,
                                    const void* const* peer_bufs, int rank, int world, void* stream) {
    RAT_REQUIRE(rows > 0 && C > 0 && z && gamma && beta && mean && rstd && out, "rat_bn_act_fwd_train: bad arguments");
    if (int rc = bnx_check(peer_bufs, rank, world, C, "rat_bn_act_fwd_train")) return rc;
    const BnExchange xc{(unsigned char* const*)peer_bufs, rank, world};
    const dim3 grid(ceil_div(C, 32), CL);
    const int per_thread = ceil_div(ceil_div(rows, CL), CL_RG);
#define RAT_BNF(NV_) k_bn_act_fwd_cl<NV_><<<grid, CL_THREADS, 0, (cudaStream_t)stream>>>(                                    \
        z, rows, C, gamma, beta, mean, rstd, running_mean, running_var, momentum, eps, out, drop_p, seed, rng_stream, rng_step_ptr(), xc)
    if (per_thread <= 8) RAT_BNF(8); else if (per_thread <= 32) RAT_BNF(32); else RAT_BNF(0);
#undef RAT_BNF
    RAT_CHECK_LAUNCH("k_bn_act_fwd_cl");
    return RAT_OK;
}

extern "C" int rat_bn_act_bwd_fused(const float* dout, const float* out, const float* z, const float* mean,
                                    const float* rstd, const float* gamma, int rows, int C, float* dz, float* dgamma,
                                    float* dbeta, float* dbias, float drop_p, unsigned long long seed,
                                    unsigned int rng_stream, float* dz_amax, const void* const* peer_bufs, int rank,
                                    int world, void* stream) {
    RAT_REQUIRE(rows > 0 && C > 0 && dout && out && dz, "rat_bn_act_bwd_fused: bad arguments");
    RAT_REQUIRE(mean == nullptr || (z && rstd && gamma && dgamma && dbeta), "rat_bn_act_bwd_fused: BatchNorm needs z / rstd / gamma / dgamma / dbeta");
    if (int rc = bnx_check(peer_bufs, rank, world, C, "rat_bn_act_bwd_fused")) return rc;
    const BnExchange xc{(unsigned char* const*)peer_bufs, rank, mean ? world : 1};      // no BatchNorm: nothing to exchange
    const dim3 grid(ceil_div(C, 32), CL);
    const int per_thread = ceil_div(ceil_div(rows, CL), CL_RG);
#define RAT_BNB(NV_) k_bn_act_bwd_cl<NV_><<<grid, CL_THREADS, 0, (cudaStream_t)stream>>>(                                    \
        dout, out, z, mean, rstd, gamma, rows, C, dz, dgamma, dbeta, dbias, drop_p, seed, rng_stream, rng_step_ptr(), dz_amax, xc)
    if (per_thread <= 8) RAT_BNB(8); else if (per_thread <= 32) RAT_BNB(32); else RAT_BNB(0);
#undef RAT_BNB
    RAT_CHECK_LAUNCH("k_bn_act_bwd_cl");
    return RAT_OK;
}

extern "C" int rat_head_bwd(const float* dlogit, int B, const float* enc, long long enc_stride, int D, float* g_fc_w,
                            float* g_fc_b, const float* h_last, int K, const float* w_final, float* g_final_w,
                            float* g_final_b, float* dh_last, void* stream) {
    RAT_REQUIRE(B > 0 && D > 0 && dlogit && enc && g_fc_w && g_fc_b, "rat_head_bwd: bad arguments");
    RAT_REQUIRE(h_last == nullptr || (K > 0 && w_final && g_final_w && g_final_b && dh_last), "rat_head_bwd: incomplete DNN arguments");
    const int slabs_h = h_last ? ceil_div(K, 32) : 0, slabs_e = ceil_div(D, 32);
    k_head_bwd_cl<<<dim3(slabs_h + slabs_e + 1, CL), CL_THREADS, 0, (cudaStream_t)stream>>>(
        dlogit, B, enc, enc_stride, D, g_fc_w, g_fc_b, h_last, K, w_final, g_final_w, h_last ? g_final_b : nullptr, dh_last,
        slabs_h, slabs_e);
    RAT_CHECK_LAUNCH("k_head_bwd_cl");
    return RAT_OK;
}
