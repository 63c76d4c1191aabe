// K7/K8: global-norm clipping + dense-equivalent fused Adam over the flat parameter buffer.
//
// Replaces BaseModel.add_regularization's gradient (base_model.py:79-94: lambda/2*||p||^2 on every
// "embedding_layer"-named parameter => grad lambda*W on EVERY row, every step), nn.utils.clip_grad_norm_
// (base_model.py:224) and torch.optim.Adam.step (base_model.py:225, torch_utils.py:41-49: lr, betas (0.9,0.999),
// eps 1e-8, no weight decay, no amsgrad) -- for all parameters in two streaming passes over [W|G|M|V].
//
// Flat layout: elements [0, reg_boundary) are "net" parameters (lambda_net), [reg_boundary, n) are
// embedding-named parameters (lambda_emb).  G holds the data gradient only (the segment-reduced sparse rows, zero
// elsewhere); the regulariser gradient lambda*W is added on the fly, so the dense [V,D] gradient of the reference
// is never materialised.  The Adam pass zeroes G for the next step.
#include "common.cuh"
#include "../../include/rat_b200.h"

namespace rat {

// state (device, float[8]): [0] grad norm, [1] clip coef, [2] lr/bias_corr1, [3] 1/sqrt(bias_corr2), [4] step (as float),
// [5] regularisation loss  sum lambda/2 ||p||^2
// lr (device float[1]) is read on the device so that lr decay needs no re-capture of a CUDA graph.

__global__ void __launch_bounds__(256) k_grad_sqnorm(const float4* __restrict__ G, const float4* __restrict__ W,
                                                     long long n4, long long boundary4, float lam_net, float lam_emb,
                                                     double* __restrict__ partial) {
    // partial[block] = sum (G + lam W)^2 ; partial[gridDim.x + block] = sum lam/2 W^2 (regularisation loss term)
    __shared__ double red[8], red2[8];
    double s = 0.0, rl = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float lam = i < boundary4 ? lam_net : lam_emb;
        float4 g = G[i];
        if (lam != 0.f) {
            const float4 w = W[i];
            g.x = fmaf(lam, w.x, g.x); g.y = fmaf(lam, w.y, g.y); g.z = fmaf(lam, w.z, g.z); g.w = fmaf(lam, w.w, g.w);
            rl += 0.5 * (double)lam * ((double)w.x * w.x + (double)w.y * w.y + (double)w.z * w.z + (double)w.w * w.w);
        }
        s += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
    }
    s = warp_sum_d(s);
    rl = warp_sum_d(rl);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red2[threadIdx.x >> 5] = rl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, t2 = 0.0;
        for (int i = 0; i < 8; ++i) { t += red[i]; t2 += red2[i]; }
        partial[blockIdx.x] = t;
        partial[gridDim.x + blockIdx.x] = t2;
    }
}

// one block: total norm (+ optional externally reduced extra sum), clip coefficient, Adam bias corrections.
// Thread t adds partials t, t+256, ... in order; the 256 thread sums are then combined by a fixed tree: deterministic.
__global__ void __launch_bounds__(256) k_optim_prepare(const double* __restrict__ partial, int nparts,
                                                       const double* __restrict__ extra_sq, float max_norm,
                                                       const float* __restrict__ lr, float beta1, float beta2,
                                                       float* __restrict__ state, int advance_step) {
    __shared__ double st[256], sr[256];
    double t = 0.0, reg = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) { t += partial[i]; reg += partial[nparts + i]; }
    st[threadIdx.x] = t; sr[threadIdx.x] = reg;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { st[threadIdx.x] += st[threadIdx.x + o]; sr[threadIdx.x] += sr[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    t = st[0]; reg = sr[0];
    if (extra_sq) { t += extra_sq[0]; reg += extra_sq[1]; }
    const float norm = (float)sqrt(t);
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.0f;     // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    if (coef > 1.0f) coef = 1.0f;
    float step = state[4];
    if (advance_step) step += 1.0f;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    state[0] = norm;
    state[1] = coef;
    state[2] = (float)((double)lr[0] / bc1);
    state[3] = (float)(1.0 / sqrt(bc2));
    state[4] = step;
    state[5] = (float)reg;
}

__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ W, float4* __restrict__ G, float4* __restrict__ M,
                                              float4* __restrict__ V, long long n4, long long boundary4, float lam_net,
                                              float lam_emb, const float* __restrict__ state, float beta1, float beta2,
                                              float eps) {
    const float coef = state[1], step_size = state[2], inv_sqrt_bc2 = state[3];
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float lam = i < boundary4 ? lam_net : lam_emb;
        float4 w = W[i], g = G[i], m = M[i], v = V[i];
        float* wp = reinterpret_cast<float*>(&w); float* gp = reinterpret_cast<float*>(&g);
        float* mp = reinterpret_cast<float*>(&m); float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gg = fmaf(lam, wp[k], gp[k]) * coef;
            mp[k] = beta1 * mp[k] + omb1 * gg;
            vp[k] = beta2 * vp[k] + omb2 * gg * gg;
            const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
            wp[k] -= step_size * (mp[k] / denom);
        }
        W[i] = w; M[i] = m; V[i] = v;
        G[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// dense gradient with the regulariser folded in (tests / debugging only): out = G + lambda*W
__global__ void k_materialize_grad(const float* __restrict__ G, const float* __restrict__ W, long long n,
                                   long long boundary, float lam_net, float lam_emb, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = G[i] + (i < boundary ? lam_net : lam_emb) * W[i];
}

}  // namespace rat

using namespace rat;

extern "C" int rat_optim_blocks(void) { return num_sms() * 8; }

extern "C" int rat_grad_sqnorm(const float* G, const float* W, long long n, long long reg_boundary, float lambda_net,
                               float lambda_emb, double* partial, void* stream) {
    RAT_REQUIRE(n > 0 && n % 4 == 0 && reg_boundary % 4 == 0, "rat_grad_sqnorm: n and reg_boundary must be multiples of 4");
    const long long n4 = n / 4;
    const int grid = rat_optim_blocks();
    k_grad_sqnorm<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)G, (const float4*)W, n4, reg_boundary / 4,
                                                          lambda_net, lambda_emb, partial);
    RAT_CHECK_LAUNCH("k_grad_sqnorm");
    return RAT_OK;
}

extern "C" int rat_optim_prepare(const double* partial, int nparts, const double* extra_sq, float max_norm,
                                 const float* lr, float beta1, float beta2, float* state, int advance_step,
                                 void* stream) {
    k_optim_prepare<<<1, 256, 0, (cudaStream_t)stream>>>(partial, nparts, extra_sq, max_norm, lr, beta1, beta2, state,
                                                        advance_step);
    RAT_CHECK_LAUNCH("k_optim_prepare");
    return RAT_OK;
}

extern "C" int rat_adam_step(float* W, float* G, float* M, float* V, long long n, long long reg_boundary,
                             float lambda_net, float lambda_emb, const float* state, float beta1, float beta2,
                             float eps, void* stream) {
    RAT_REQUIRE(n > 0 && n % 4 == 0 && reg_boundary % 4 == 0, "rat_adam_step: n and reg_boundary must be multiples of 4");
    const int grid = rat_optim_blocks();
    k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)W, (float4*)G, (float4*)M, (float4*)V, n / 4,
                                                   reg_boundary / 4, lambda_net, lambda_emb, state, beta1, beta2, eps);
    RAT_CHECK_LAUNCH("k_adam");
    return RAT_OK;
}

extern "C" int rat_materialize_grad(const float* G, const float* W, long long n, long long reg_boundary,
                                    float lambda_net, float lambda_emb, float* out, void* stream) {
    const int grid = rat_optim_blocks();
    k_materialize_grad<<<grid, 256, 0, (cudaStream_t)stream>>>(G, W, n, reg_boundary, lambda_net, lambda_emb, out);
    RAT_CHECK_LAUNCH("k_materialize_grad");
    return RAT_OK;
}
