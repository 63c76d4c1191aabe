/* rat_b200.h -- C ABI of librat_b200.so: the B200 (sm_100a) hot path of RAT (WWW'24) retrieval-augmented CTR.
 *
 * Boundary contract:
 *   - plain C, raw DEVICE pointers + explicit sizes; no torch types.  All float tensors are fp32, contiguous,
 *     row-major; ids are int32; every pointer must be 16-byte aligned (torch CUDA allocations are).
 *   - the library never allocates or frees user-visible memory; scratch comes from caller-provided workspaces
 *     whose size is returned by the matching *_workspace_bytes() query.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - return value: 0 = RAT_OK, negative = error (RAT_EINVAL -1, RAT_ECUDA -2, RAT_ESMEM -3); the message is
 *     available from rat_last_error().  No exceptions, no exit().
 *   - there is NO CPU fallback: on a machine without an sm_100 device every entry point fails with RAT_ECUDA.
 *
 * Each entry point cites the reference code it replaces (paths relative to the reference repository root;
 * all of it is eager-PyTorch Python -- the reference has no native code, so this ABI is what a ctypes binding
 * inside the reference's fuxictr/pytorch/models/base_model.py would bind).
 */
#ifndef RAT_B200_H
#define RAT_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAT_ABI_VERSION 1

const char* rat_last_error(void);
int rat_abi_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py's gpu_launches) */
long long rat_launch_count(void);
/* 0 if an sm_100 device is current, else RAT_ECUDA (the product path refuses to run anywhere else) */
int rat_device_check(void);

/* Arithmetic of the RAT-block projections and DNN GEMMs: 2 (default) = 5th-gen tensor cores (tcgen05.mma kind::f16,
 * fp16 operands, fp32 accumulate in TMEM, dynamic power-of-two gradient scaling; residual stream / LayerNorm /
 * softmax / GELU / BatchNorm in fp32); 1 = mma.sync TF32 operands with fp32 accumulate (parity tolerance 3e-3
 * relative); 0 = exact fp32 on the SIMT pipe (parity anchor, ~1e-5). */
int rat_set_precision(int mode);
int rat_get_precision(void);
/* Deferred weight-gradient reductions.  rat_attn_bwd / rat_ff_bwd are two launches: the backward kernel (per-CTA gradient
 * records into `workspace`) and a fixed-order reduction of the records into dW*.  Only the optimizer reads dW*, so with a
 * reduce stream set the reduction is launched THERE (forked from the call's stream by an event) and the next backward
 * kernel does not queue behind it; a later call that reuses the same `workspace` first waits for the reduction that still
 * reads it (alternate two workspaces to get the overlap).  rat_reduce_stream_join makes `stream` wait for every deferred
 * reduction issued so far -- call it before anything reads the gradients and before a stream capture ends.  NULL = off
 * (default).  Events only: safe inside CUDA-graph capture.  No reference counterpart (autograd's engine orders its own
 * accumulation kernels); the contract of rat_attn_bwd / rat_ff_bwd is unchanged after the join. */
int rat_set_reduce_stream(void* stream);
int rat_reduce_stream_join(void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K0: batch assembly
 * ------------------------------------------------------------------------------------------------------- */
/* Wire-format conversion.  Replaces BaseModel.inputs_to_device (fuxictr/pytorch/models/base_model.py:125-133)
 * + the `.long()` casts of EmbeddingDictLayer.forward (fuxictr/pytorch/layers/embedding.py:166,169) and the
 * token=2 construction (fuxictr/pytorch/models/RAT_m2.py:115-123).
 * X [B,T,L] f64 ids, y [B,T] f64 labels (row 0 = target) -> ids [B,T,L] i32, labels [B,T] i32 (labels[:,0]=2),
 * y_true [B] f32. */
int rat_convert_wire_f64(const double* X, const double* y, int* ids, int* labels, float* y_true, int B, int T, int L,
                         void* stream);

/* Device-resident retrieval-set assembly.  Replaces Dataset.__getitem__ + default collate
 * (fuxictr/pytorch/data_generator.py:66-78): ids[b,0,:] = q_ids[rows[b]], ids[b,1+k,:] = pool_ids[nbr[rows[b],k]]
 * with numpy negative-index wraparound (-1 -> last pool row).  rows == NULL means rows[b] = row0 + b.
 * err_flag (device int, caller-zeroed) gets bit 1 set on an out-of-range neighbour index. */
int rat_assemble_ids(const int* q_ids, const unsigned char* q_labels, const long long* rows, long long row0,
                     const int* pool_ids, const unsigned char* pool_labels, const long long* nbr, long long n_pool,
                     int* ids, int* labels, float* y_true, int B, int T, int L, int* err_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K1: fused FeatureEmbedding gather
 * ------------------------------------------------------------------------------------------------------- */
/* Replaces EmbeddingLayer/EmbeddingDictLayer.forward + dict2tensor (fuxictr/pytorch/layers/embedding.py:40-43,
 * 138-178), MaskedSumPooling (layers/sequence.py:36-38), the label embedding + 3 concats (models/RAT_m2.py:
 * 117-126), nn.Dropout(emb_dropout) (RAT_m2.py:135) and LR_Layer.forward (layers/shallow.py:36-45).
 * emb_W [V_total,D] / lr_W [V_total] are the per-field tables concatenated in field order; col_off[l] is the first
 * row of column l's table, col_vocab[l] its vocab size; field f spans columns [field_col0[f], +field_width[f]).
 * Outputs: block [B,T,F+1,D] (dropout applied iff drop_p>0), x_emb [B,F*D] (target row, never dropped; may be
 * NULL), lr_out [B] (may be NULL together with lr_W).  err_flag bit 0: id out of vocabulary, bit 2: bad label. */
int rat_gather_fwd(const float* emb_W, const float* lr_W, const float* label_W, const int* ids, const int* labels,
                   const int* col_off, const int* col_vocab, const int* field_col0, const int* field_width,
                   float* block, float* x_emb, float* lr_out, int B, int T, int L, int F, int D, float drop_p,
                   unsigned long long seed, unsigned int rng_stream, int* err_flag, void* stream);
/* Row-sharded tables (tmall configuration): the concatenated table index space is range-partitioned,
 * rank o owns rows [o*rows_per_shard, (o+1)*rows_per_shard).  W_peers[o] is the base of rank o's flat parameter
 * buffer mapped into this process (symmetric memory over NVLink/NVSwitch); its [rows_per_shard, D] embedding shard
 * starts at float offset emb_off and its LR shard at lr_off.  The gather kernel loads remote rows straight from
 * peer memory (no staging all-to-all); everything else is identical to rat_gather_fwd (bit-exact). */
int rat_gather_fwd_sharded(const float* const* W_peers, long long emb_off, long long lr_off, int rows_per_shard,
                           int world, const float* label_W, const int* ids, const int* labels, const int* col_off,
                           const int* col_vocab, const int* field_col0, const int* field_width, float* block,
                           float* x_emb, float* lr_out, int B, int T, int L, int F, int D, float drop_p,
                           unsigned long long seed, unsigned int rng_stream, int* err_flag, void* stream);
/* One-shot all-reduce (SUM) of a small float64 vector over NVLink peer memory: the data-parallel hook for the raw BatchNorm
 * sums of MLP_Layer (layers/deep.py:128-135 at the GLOBAL batch) and the shard-norm partials.  peer_bufs: device array of
 * `world` pointers to the ranks' symmetric buffers (rat_oneshot_workspace_bytes each, zero-initialised once); every rank
 * must enqueue the call the same number of times.  in == out is allowed.  No host value changes between calls, so the call
 * captures into a CUDA graph. */
size_t rat_oneshot_workspace_bytes(int world);
int rat_oneshot_allreduce_f64(const void* const* peer_bufs, int rank, int world, const double* in, int n, double* out,
                              void* stream);

/* Dropout streams.  Every kernel with dropout derives its mask from (seed, rng_stream + 64 * step, element) where `step` is
 * a DEVICE-resident counter owned by the library: nn.Dropout's "a new mask every training step" (RAT_m2.py:83,135,
 * layers/deep.py:134) without a host-side value baked into the launch, so that a CUDA graph of the whole training step
 * replays with fresh masks.  rat_rng_step_advance enqueues step += 1 (call it once at the start of a training step, inside
 * the captured region); rat_rng_step_set enqueues step = value (tests, checkpoint resume). */
int rat_rng_step_set(unsigned int value, void* stream);
int rat_rng_step_advance(void* stream);
/* backward of the embedding dropout: grad *= mask/(1-p), same mask as rat_gather_fwd */
int rat_dropout_bwd(float* grad, long long n, float p, unsigned long long seed, unsigned int rng_stream,
                    void* stream);

/* dst[r*dst_stride + d] = src[r*src_stride + d] for d < D: the `x[:, 0]` token pooling between the two
 * Transformers of RAT_m1 (models/RAT_m1.py:125-126) and, with the strides swapped, its backward scatter. */
int rat_strided_copy(const float* src, float* dst, long long rows, int D, long long src_stride, long long dst_stride,
                     void* stream);
/* dst [rows*group, D]: dst[r*group + 0] = src[r], every other row zero.  RAT_m2 consumes only token (t=0, n=0) of the
 * encoder output (models/RAT_m2.py:138-140), so the last block's cross attention / FeedForward run on the field-token-0
 * rows only; this scatters their gradient back into the full [B,T,N,D] block gradient (and replaces its memset). */
int rat_expand_rows(const float* src, float* dst, long long rows, int D, int group, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K2: fused RAT block (forward)
 * ------------------------------------------------------------------------------------------------------- */
/* out = res + alpha * (MHA(LayerNorm(x)) Wo^T + bo) on a [B,T,N,D] tensor.  mode 0 ("intra"): sequences are the
 * B*T rows of N tokens; mode 1 ("cross"): sequences are the B*N columns of T tokens (the reference's
 * reshape/transpose/flatten, RAT_m2.py:221-235, becomes index arithmetic).  Replaces PreNorm (RAT_m2.py:155-161)
 * + Attention (RAT_m2.py:176-202; RAT_m3.py:164-196 when Wq/Wk/Wv are separate tensors) + the residual add.
 * Wq/Wk/Wv [heads*dim_head, D], Wo [D, heads*dim_head], bo [D].  res may be NULL (RAT_m3), alpha scales the
 * attention branch (0.5 for RAT_m3's mean of the two branches). x, res and out may alias each other only if
 * res == x == out is NOT used (out must not alias x). */
int rat_attn_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b,
                 const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* bo, int B, int T,
                 int N, int D, int heads, int dim_head, float scale, float alpha, int mode, void* stream);
/* out = res + W2 gelu_erf(W1 LNopt(x) + b1) + b2 over `rows` tokens.  Replaces FeedForward (RAT_m2.py:163-174);
 * ln_w/ln_b non-NULL adds the PreNorm of RAT_m0.py:197-201.  W1 [M,D], W2 [D,M]. */
int rat_ff_fwd(const float* x, const float* res, float* out, const float* ln_w, const float* ln_b, const float* W1,
               const float* b1, const float* W2, const float* b2, long long rows, int D, int M, void* stream);
/* out = LayerNorm(x) (eps 1e-5).  Final norm of the RAT_m0/m1 Transformer (RAT_m0.py:202,208). */
int rat_layernorm_fwd(const float* x, float* out, const float* w, const float* b, long long rows, int D,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K3/K4: DNN head, BatchNorm, loss
 * ------------------------------------------------------------------------------------------------------- */
/* C[M,N] = opA(A) opB(B) (+bias[n]).  trans_a=0: A[m*lda+k], 1: A[k*lda+m]; trans_b=0: B[n*ldb+k] (torch Linear
 * weight), 1: B[k*ldb+n].  Replaces the nn.Linear calls of MLP_Layer (layers/deep.py:126,137) and their
 * autograd backward.  Deterministic split-K when a workspace of rat_sgemm_workspace_bytes() is supplied.
 * rat_sgemm_scaled: backward products (A = dz).  a_amax (device float, may be NULL) holds max|A| (rat_absmax); in the
 * tensor-core precision mode (fp16 operands) A is lifted by the power of two that puts that maximum into [4, 8) while
 * it is converted and the fp32 accumulator is unscaled, so that small gradients stay in the fp16 normal range and
 * large ones cannot overflow.  The fp32 / tf32 kernels ignore it.  rat_sgemm == a_amax NULL.
 * rat_absmax: *out = max(*out, max |x[r*row_stride + c]|) over r < rows, c < cols (caller zeroes *out). */
size_t rat_sgemm_workspace_bytes(int M, int N, int K);
int rat_sgemm_scaled(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                     int ldc, int trans_a, int trans_b, const float* a_amax, float* workspace, size_t workspace_bytes,
                     void* stream);
int rat_absmax(const float* x, long long rows, int cols, long long row_stride, float* out, void* stream);
int rat_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
              int ldc, int trans_a, int trans_b, float* workspace, size_t workspace_bytes, void* stream);
/* BatchNorm1d (layers/deep.py:128-129), train mode: raw per-column sums (double, [2C]: sum, sum of squares) so a
 * data-parallel caller can all-reduce them, then mean/rstd + running-stat update (momentum 0.1, unbiased var). */
int rat_bn_sums(const float* z, int rows, int C, double* sums, void* stream);
int rat_bn_finalize(const double* sums, double count, int C, float* mean, float* rstd, float* running_mean,
                    float* running_var, float momentum, float eps, void* stream);
int rat_bn_eval_stats(const float* running_mean, const float* running_var, int C, float* mean, float* rstd,
                      float eps, void* stream);
/* out = dropout(relu(bn(z))) ; mean==NULL skips the normalisation (batch_norm: false). */
int rat_bn_act_fwd(const float* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                   float* out, int rows, int C, float drop_p, unsigned long long seed, unsigned int rng_stream,
                   void* stream);
int rat_bn_act_bwd_sums(const float* dout, const float* out, const float* z, const float* mean, const float* rstd,
                        int rows, int C, float drop_p, unsigned long long seed, unsigned int rng_stream, double* sums,
                        void* stream);
/* dz_amax (device float, caller-zeroed, may be NULL): receives max|dz| (see rat_sgemm_scaled).
 * param_grad_scale multiplies the emitted dgamma / dbeta: `sums` are GLOBAL sums in data-parallel training (all-reduced
 * by the caller), so each rank writes 1/world of them and the later SUM all-reduce of the gradient restores them. */
int rat_bn_act_bwd_apply(const float* dout, const float* out, const float* z, const float* mean, const float* rstd,
                         const float* gamma, const double* sums, double count, float* dz, float* dgamma,
                         float* dbeta, int rows, int C, float drop_p, unsigned long long seed,
                         unsigned int rng_stream, float* dz_amax, float param_grad_scale, void* stream);
int rat_colsum(const float* A, int rows, int C, int lda, float* out, void* stream);
/* Single-launch forms of the above for single-process training (csrc/mlp_fused.cu: one thread-block cluster of 8 CTAs per
 * 32-column slab, column partials exchanged through distributed shared memory, rank-order sums => deterministic).
 * rat_bn_act_fwd_train == rat_bn_sums + rat_bn_finalize(count = rows) + rat_bn_act_fwd: nn.BatchNorm1d in train mode +
 * ReLU + Dropout of MLP_Layer (layers/deep.py:128-135); mean / rstd are saved for the backward.
 * rat_bn_act_bwd_fused == rat_bn_act_bwd_sums + rat_bn_act_bwd_apply(count = rows, param_grad_scale = 1) + rat_colsum(dz):
 * autograd's backward of the same three modules plus the bias gradient of the nn.Linear in front of them (dbias may be
 * NULL; mean == NULL: no BatchNorm, dz = d relu/dropout).  dz may alias dout.  See below for peer_bufs / rank / world. */
int rat_bn_act_fwd_train(const float* z, int rows, int C, const float* gamma, const float* beta, float* mean,
                         float* rstd, float* running_mean, float* running_var, float momentum, float eps, float* out,
                         float drop_p, unsigned long long seed, unsigned int rng_stream, const void* const* peer_bufs,
                         int rank, int world, void* stream);
int rat_bn_act_bwd_fused(const float* dout, const float* out, const float* z, const float* mean, const float* rstd,
                         const float* gamma, int rows, int C, float* dz, float* dgamma, float* dbeta, float* dbias,
                         float drop_p, unsigned long long seed, unsigned int rng_stream, float* dz_amax,
                         const void* const* peer_bufs, int rank, int world, void* stream);
/* Data-parallel form of the two calls above (world > 1): the statistics are those of the GLOBAL batch (rows * world rows;
 * every rank passes the same `rows`), i.e. what the single-device reference computes at that batch.  peer_bufs = device
 * array of `world` pointers to the ranks' symmetric exchange buffers (rat_bn_exchange_workspace_bytes each, zeroed once;
 * NVLink peer memory); the slab leaders exchange their 2 x 32 double totals inside the launch (one-shot store / flag /
 * rank-order sum), so no separate all-reduce sits between the statistics and the apply pass.  dgamma / dbeta are written
 * as 1/world of the global value (the caller's SUM all-reduce of the gradient buffer restores them); dbias is the LOCAL
 * column sum.  Every rank must enqueue the same sequence of calls.  world == 1: peer_bufs may be NULL. */
size_t rat_bn_exchange_workspace_bytes(int world);
/* Every gradient that hangs off dlogit [B] in ONE launch: g_fc_w[d] = sum_b dlogit[b] enc[b*enc_stride + d] and
 * g_fc_b = sum_b dlogit[b] (self.fc, RAT_m2.py:106,144) and, when h_last != NULL, the final Linear(K -> 1) of the DNN
 * (layers/deep.py:137): g_final_w[k] = sum_b dlogit[b] h_last[b][k], g_final_b = sum_b dlogit[b],
 * dh_last[b][k] = dlogit[b] w_final[k].  Replaces 3 rat_sgemm + 2 rat_colsum calls (10 launches). */
int rat_head_bwd(const float* dlogit, int B, const float* enc, long long enc_stride, int D, float* g_fc_w,
                 float* g_fc_b, const float* h_last, int K, const float* w_final, float* g_final_w, float* g_final_b,
                 float* dh_last, void* stream);
/* logit = fc(enc[b,0,0,:]) + dnn_out[b] + lr_out[b]; y_pred = sigmoid; BCE(mean, log clamp -100) and, when dlogit
 * is non-NULL, dlogit[b] = dBCE/dlogit * inv_count and denc[b,0,0,:] = dlogit[b]*fc_w (the other tokens of denc are the
 * caller's business).  Replaces RAT_m2.py:138-150 + BaseModel.add_loss (base_model.py:74-77).  One warp per sample, D <= 128.
 * loss_part: double [2 * rat_head_blocks(B)] scratch.  denc_amax (device float, may be NULL): receives max|denc| -- the seed
 * of the fp16 mode's gradient-scale chain (K5) -- as a plain store (max|dlogit| * max|fc_w|, exact), no caller zeroing. */
int rat_head_blocks(int B);
int rat_head(const float* enc, long long enc_stride, const float* fc_w, const float* fc_b, const float* dnn_out,
             const float* lr_out, const float* y_true, int B, int D, float* y_pred, float* dlogit, float* denc,
             float inv_count, double* loss_part, float* loss_sum, float* loss_mean, float* denc_amax, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K5: fused RAT block (backward).  Replace autograd's reverse of RAT_m2.py:155-236 (loss.backward(),
 * base_model.py:223).  Every kernel recomputes the sub-block's intermediates in shared memory from the saved
 * sub-block INPUT x and writes  dx = (base ? base : 0) + d(sub-block)/dx ; weight gradients are summed
 * deterministically (per-CTA partials in `workspace`, fixed-order reduction) and STORED to dW* (dWq is
 * accumulated instead when accumulate_wq != 0: RAT_m3 shares W_q between its two attentions).
 * dx may alias dout/base (in place).  Any dW* pointer may be NULL (gradient discarded).
 * dout_amax / dx_amax (device floats, may be NULL): dynamic gradient scaling of the fp16 tensor-core mode.  *dout_amax
 * = max|dout| lets the kernel lift dout by a power of two into the fp16 normal range (results are unscaled in fp32);
 * max|dx| is merged into *dx_amax (caller-zeroed) for the next kernel of the chain.  NULL dout_amax = no scaling.
 * ------------------------------------------------------------------------------------------------------- */
size_t rat_attn_bwd_workspace_bytes(int B, int T, int N, int D, int heads, int dim_head, int mode);
int rat_attn_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                 const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                 float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq, int B,
                 int T, int N, int D, int heads, int dim_head, float scale, float alpha, int mode,
                 const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, void* stream);
/* rat_attn_bwd with the backward of an nn.Dropout on dx fused into its last store: dx *= mask / (1 - p) with the mask of
 * (out_drop_p, seed, rng_stream) over the flattened [B,T,N,D] element index -- for the encoder's first sub-block this is the
 * backward of nn.Dropout(emb_dropout) (models/RAT_m2.py:83,135) with the mask rat_gather_fwd drew, so that the segment
 * reduce reads a ready gradient.  out_drop_p = 0 is rat_attn_bwd. */
int rat_attn_bwd_dropout(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                         const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                         float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq, int B,
                         int T, int N, int D, int heads, int dim_head, float scale, float alpha, int mode,
                         const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, float out_drop_p,
                         unsigned long long seed, unsigned int rng_stream, void* stream);
size_t rat_ff_bwd_workspace_bytes(long long rows, int D, int M);
int rat_ff_bwd(const float* x, const float* dout, const float* base, float* dx, const float* ln_w, const float* ln_b,
               const float* W1, const float* b1, const float* W2, float* dW1, float* db1, float* dW2, float* db2,
               float* dln_w, float* dln_b, long long rows, int D, int M, const float* dout_amax, float* dx_amax,
               float* workspace, size_t workspace_bytes, void* stream);
size_t rat_layernorm_bwd_workspace_bytes(long long rows, int D);
int rat_layernorm_bwd(const float* x, const float* dout, float* dx, const float* w, float* dw, float* db,
                      long long rows, int D, float* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K6: embedding gradient = deterministic sorted segment-reduce.  Replaces ATen embedding_dense_backward for
 * every nn.Embedding of EmbeddingDictLayer (layers/embedding.py:79-100), LR_Layer (layers/shallow.py:31) and the
 * label table (RAT_m2.py:64), and the backward of nn.Dropout(emb_dropout) (RAT_m2.py:135).  Occurrence (b,t,l)
 * contributes mask*dblock[b,t,1+field(l),:] (+ dxemb[b,field(l),:] when t==0) to row col_off[l]+ids[b,t,l] of g_emb,
 * and dlogit[b] (t==0 only) to the same row of g_lr; padding ids contribute nothing (torch padding_idx semantics).
 * g_label [3,D] = sum of mask*dblock[b,t,0,:] by labels[b,t].  mask = the rat_gather_fwd dropout mask of
 * (drop_p, seed, rng_stream) times 1/(1-p); drop_p = 0 means no mask (dblock is never modified).
 * Only touched rows of g_emb / g_lr / g_label are written (each exactly once); the caller keeps the rest zero.
 * rat_emb_scatter_plan builds and sorts the occurrence keys; it needs the ids only, so the caller may run it on
 * another stream while the forward/backward kernels run, then call rat_emb_scatter_reduce(planned=1) with the same
 * workspace.  planned=0 makes rat_emb_scatter_reduce run the plan itself first.  n_occ = B*T*(L+1).
 * ------------------------------------------------------------------------------------------------------- */
size_t rat_emb_scatter_workspace_bytes(long long n_occ, int D);
int rat_emb_scatter_plan(const int* ids, const int* labels, const int* col_off, const int* col_pad,
                         const int* col_vocab, int B, int T, int L, int F, int D, long long V_total, void* workspace,
                         size_t workspace_bytes, void* stream);
int rat_emb_scatter_reduce(const int* ids, const int* labels, const float* dblock, const float* dxemb,
                           const float* dlogit, const int* col_off, const int* col_pad, const int* col_vocab,
                           const int* col_field, float* g_emb, float* g_lr, float* g_label, int B, int T, int L, int F,
                           int D, long long V_total, float drop_p, unsigned long long seed, unsigned int rng_stream,
                           int planned, void* workspace, size_t workspace_bytes, void* stream);

/* Row-sharded tables (all-to-all #3; the tables are the nn.Embedding's of layers/embedding.py:79-100 and
 * layers/shallow.py:31): after rat_emb_scatter_reduce has written this rank's reduced rows into a dense scratch indexed by
 * GLOBAL row (g_emb_full [world*rows_per_shard, D], g_lr_full), rat_shard_send_rows stores every touched row once into its
 * owner's receive buffer over NVLink (recv_peers: device array of the ranks' symmetric receive buffers,
 * rat_shard_recv_bytes each; `counts`: `world` zero-initialised device counters) and clears it in the scratch.  After a
 * cross-rank barrier rat_shard_apply_rows adds the received records into the local shard's gradient, source ranks in rank
 * order (deterministic).  Traffic is proportional to the batch (<= B*T*L rows), not to the vocabulary. */
size_t rat_shard_recv_bytes(int B, int T, int L, int D, int world);
int rat_shard_send_rows(void* workspace, size_t workspace_bytes, int B, int T, int L, int F, int D, long long V_total,
                        int rows_per_shard, int world, int rank, float* g_emb_full, float* g_lr_full,
                        const void* const* recv_peers, unsigned int* counts, int* err_flag, void* stream);
int rat_shard_apply_rows(const void* recv_local, int B, int T, int L, int D, int world, int rank, int rows_per_shard,
                         float* g_emb_local, float* g_lr_local, int* err_flag, void* stream);
/* the stable LSD radix sort used above, exposed for tests: sorts (keys, vals) by the low `bits` bits of keys.
 * hist: scratch of 256*ceil(n/2048) uint32.  *result_in_tmp = 1 if the sorted data ended in the tmp buffers. */
int rat_radix_sort_pairs(unsigned int* keys, unsigned int* vals, unsigned int* keys_tmp, unsigned int* vals_tmp,
                         unsigned int* hist, long long n, int bits, int* result_in_tmp, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K7/K8: regulariser gradient + global-norm clip + dense-equivalent fused Adam over the flat buffers.
 * Replace BaseModel.add_regularization (base_model.py:79-94), nn.utils.clip_grad_norm_ (base_model.py:224) and
 * torch.optim.Adam.step (base_model.py:225; torch_utils.py:41-49).  Elements [0,reg_boundary) use lambda_net,
 * the rest lambda_emb (the `"embedding_layer" in name` rule).  `partial`: double[2*rat_optim_blocks()].
 * `state`: device float[8] = {grad_norm, clip_coef, lr/bc1, 1/sqrt(bc2), step, reg_loss, -, -}; `lr`: device
 * float[1].  extra_sq (device double[2], may be NULL): {sum g^2, reg loss} contributed by other ranks' shards.
 * ------------------------------------------------------------------------------------------------------- */
int rat_optim_blocks(void);
int rat_grad_sqnorm(const float* G, const float* W, long long n, long long reg_boundary, float lambda_net,
                    float lambda_emb, double* partial, void* stream);
int rat_optim_prepare(const double* partial, int nparts, const double* extra_sq, float max_norm, const float* lr,
                      float beta1, float beta2, float* state, int advance_step, void* stream);
int rat_adam_step(float* W, float* G, float* M, float* V, long long n, long long reg_boundary, float lambda_net,
                  float lambda_emb, const float* state, float beta1, float beta2, float eps, void* stream);
int rat_materialize_grad(const float* G, const float* W, long long n, long long reg_boundary, float lambda_net,
                         float lambda_emb, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Validation metrics on the device.  Replaces evaluate_metrics (fuxictr/metrics.py:21-41:
 * sklearn roc_auc_score + log_loss on predictions clipped to [1e-7, 1-1e-7]) for a whole evaluation pass:
 * out[0] = exact tie-aware AUC (integer rank statistics), out[1] = logloss (float64), out[2] = #positives,
 * out[3] = #negatives.  y_true > 0.5 counts as positive.  Deterministic.
 * ------------------------------------------------------------------------------------------------------- */
size_t rat_auc_logloss_workspace_bytes(long long n);
int rat_auc_logloss(const float* y_pred, const float* y_true, long long n, double* out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BM25 top-K retrieval on the device.  Replaces the compare-scan + top-K of
 * BM25_topk_retrieval_v4 (fuxictr/datasets/data_utils.py:773-1064; called from
 * fuxictr/pytorch/data_generator.py:141-168,192-207):
 *   db [N][E + F] int32, qry [Q][E + F] int32 (the E exact-match columns first), qry_idf [Q][F] float64 = IDF of the
 *   query's value in each scored column (0 when the value does not occur in the db; data_utils.py:842-846,879-887).
 *   score = sum_f (qry == db) * qry_idf (float64, in the association torch's sum(-1) uses for < 20 elements: bit-exact
 *   `values`; data_utils.py:950), restricted to db rows that agree with
 *   the query on all E exact-match columns, +1 when E > 0 (data_utils.py:947); unit_scores: every candidate scores 1
 *   (pure exact matching, data_utils.py:912-917,1033-1038).
 *   Outputs (data_utils.py:787-797): values [Q][K] float64 descending, indices [Q][K] int64 (-1 = fewer than K rows with
 *   a non-zero score), lens [Q] int64.  Equal scores: smaller db index first (prefer_last: larger first -- the
 *   `truncating="pre"` of the reference's pad_sequences keeps the LAST K members of an exact-match group).
 *   K <= 32, E + F <= 24, F <= 19.  Deterministic.
 * ------------------------------------------------------------------------------------------------------- */
size_t rat_bm25_topk_workspace_bytes(long long N, long long Q, int K);
int rat_bm25_topk(const int* db, long long N, const int* qry, const double* qry_idf, long long Q, int E, int F, int K,
                  int unit_scores, int prefer_last, double* values, long long* indices, long long* lens,
                  void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAT_B200_H */
