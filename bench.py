#!/usr/bin/env python
"""bench.py -- RAT_m2 training / inference throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape kkbox|ml|tmall] [--batch 4096] [--topk 5]
    python bench.py --impl reference ...     # the reference algorithm on the host CPU cores (oracle port)

One "step" = one full RAT_m2 training step (retrieval-set assembly + gather -> 4 RAT blocks -> DNN head -> loss ->
backward -> sorted segment-reduce -> global-norm clip -> dense-equivalent Adam) on one batch of B=4096 samples PER GPU
(weak scaling) of synthetic data with the kkbox_x1_10fold_retrieval shape (BASELINE.json configs[1]).

Printed JSON line (rank 0):
  value   whole-job train samples/s, inputs (id matrix, pool, neighbour index) resident in HBM
  e2e     the same through the public FuxiCTR API (model.train_step(batch)) with HOST float64 wire-format batches in
          pinned memory: H2D of every batch and a D2H read of the loss inside the timed region
  infer   the same two numbers for model.forward (eval mode)
  roofline / roofline_gather / roofline_adam / kernels   per-kernel CUDA-event timings vs measured peaks
  cpu_baseline   the oracle (CPU restatement of the reference) at the SAME batch size on the host cores, N=1 only
  secondary      the other BASELINE.json workloads measured in the same run (a few steps each): `strict` (TF32 arithmetic),
                 `shapes` (movielens / tmall), `infer_sweep` (K x B inference sweep, configs[4]), `strong` (global batch
                 4096 split over the N GPUs) and, when N > 1, `tmall_sharded` (row-sharded tables, configs[2])
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; every step uses a
different batch and the step's working set (>600 MB of activations) is larger than L2, so no L2 flush is needed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))

import numpy as np
import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="kkbox", choices=["kkbox", "ml", "tmall"])
    ap.add_argument("--batch", type=int, default=4096, help="per-GPU batch size")
    ap.add_argument("--topk", type=int, default=5)
    ap.add_argument("--pool-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=4096, help="batch size of the CPU arms (default: the GPU arm's)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (strict / shapes / sweep / strong / sharded)")
    ap.add_argument("--shard-tables", action="store_true",
                    help="row-shard the embedding / LR tables over the ranks (BASELINE configs[2]; needs --gpus > 1)")
    ap.add_argument("--vocab-scale", type=float, default=1.0, help="scale every vocabulary (scaled-vocab tmall variant)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "tf32", "fp32"],
                    help="projection / DNN GEMM arithmetic: fp16 = tcgen05, fp16 operands + fp32 accumulate (default), "
                         "tf32 = mma.sync, fp32 = SIMT")
    return ap.parse_args()


def ncu_traffic(key, shape, B, K):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py).  Every entry records the sha256 of the kernel's source
    files at capture time and the workload it was taken on: a stale capture (source changed since) or another workload
    reports null instead of a number that no longer belongs to the code."""
    import hashlib
    try:
        db = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = db[key]
        if (e["shape"], e["B"], e["K"]) != (shape, B, K):
            return None
        h = hashlib.sha256()
        for f in e["sources"]:
            h.update(open(os.path.join(ROOT, f), "rb").read())
        return float(e["traffic_bytes"]) if h.hexdigest()[:16] == e["sha16"] else None
    except Exception:
        return None



def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------- algorithmic work
def gather_bytes_per_sample(K, L, F, D):
    T, N = K + 1, F + 1            # nbr idx + ids + labels + table rows + LR scalars + block + x_emb
    return K * 8 + T * L * 4 + T + T * L * D * 4 + L * 4 + T * N * D * 4 + F * D * 4


def encoder_flops_per_sample(K, F, D, H, dh, scale_dim, depth=4):
    T, N, I, M = K + 1, F + 1, H * dh, D * scale_dim
    return depth * (T * N * (16 * D * I + 4 * D * M) + 4 * I * T * N * (N + T))


def attn_flops_per_token(D, I, S):
    return 2 * D * 3 * I + 2 * I * D + 4 * I * S


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # the upper half = samples taken under load
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def timed(fn_step, steps, warmup, dist):
    """W untimed + exactly K timed steps, barrier+sync both sides, CUDA events, max over ranks. Returns seconds."""
    for i in range(warmup):
        fn_step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn_step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / 1e3


# ----------------------------------------------------------------------------------------- our arm
def run_ours(a):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    import rat_native as rn
    from rat_native import shapes
    from fuxictr.pytorch import models
    from fuxictr.pytorch.data_generator import DeviceDataGenerator
    from fuxictr.pytorch.torch_utils import seed_everything
    rn.require_device()
    from rat_native.engine import set_precision
    set_precision(a.precision)
    seed_everything(2021)
    B, K, S = a.batch, a.topk, a.shape
    cfg = shapes.SHAPES[S]
    fm = shapes.make_feature_map(S, vocab_scale=a.vocab_scale)
    params = shapes.model_params(S, K=K, gpu=local)
    sharded = bool(a.shard_tables and world > 1)
    if sharded:
        params["shard_embeddings"] = True
    os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
    model = models.RAT_m2(fm, **params)
    if world > 1 and not sharded:               # identical replicas: broadcast rank 0's initial weights
        dist.broadcast(model._engine.store.W, 0)
    elif sharded:                               # replicated part only; every rank keeps its own table shard
        dist.broadcast(model._engine.store.W[:model._engine.store.emb_off], 0)
    n_params = model.count_parameters()
    hp = cfg["hp"]
    F, L, D, H = fm.num_fields, fm.input_length, hp["embedding_dim"], hp["num_heads"]
    T, N = K + 1, F + 1

    # ---- synthetic data (same seed on every rank; each rank takes its contiguous slice of every global batch)
    pool = shapes.synthetic_array(fm.feature_specs, a.pool_rows, seed=2021, pos_ratio=cfg["pos_ratio"])
    nbr = shapes.synthetic_neighbours(a.pool_rows, a.pool_rows, K, seed=2021)
    gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B * world, shuffle=True, device=f"cuda:{local}", seed=2021,
                              rank=rank, world=world, drop_last=True)
    n_steps_total = a.warmup + a.steps
    it = iter(gen)
    dev_batches = [next(it) for _ in range(n_steps_total)]
    # host wire-format batches (pinned) for the e2e leg: disjoint rows per step and rank
    rng = np.random.default_rng(1234 + rank)
    host_batches = []
    for i in range(n_steps_total):
        rows = rng.integers(0, a.pool_rows, size=B)
        X, y, v, l = shapes.host_wire_batch(pool, pool, nbr, rows)
        host_batches.append(tuple(torch.from_numpy(t).pin_memory() for t in (X, y, v, l)))
    h2d = host_batches[0][0].numel() * 8 + host_batches[0][1].numel() * 8

    model.train()
    model._max_gradient_norm = 10.0
    sampler = ClockSampler(local) if rank == 0 else None
    def launch_count():         # direct C-ABI launches + kernels replayed through the CUDA graph of the training step
        return int(rn.query("rat_launch_count")) + int(model._engine.replayed_launches)
    for i in range(a.warmup):                      # warm-up here so that the launch count brackets exactly the timed steps
        model.train_step(dev_batches[i])
    launches0 = launch_count()
    t_train = timed(lambda i: model.train_step(dev_batches[a.warmup + i]), a.steps, 0, dist)
    launches = launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    model._engine.check_errors()

    def e2e_train(i):
        loss = model.train_step(host_batches[i])
        return float(loss.item())                 # device -> host read of the step's result
    t_e2e = timed(e2e_train, a.steps, a.warmup, dist)

    model.eval()
    with torch.no_grad():
        t_inf = timed(lambda i: model.forward(dev_batches[i]), a.steps, a.warmup, dist)
        out_host = torch.empty(B, 1, dtype=torch.float32).pin_memory()

        def e2e_inf(i):
            rd = model.forward(host_batches[i])
            out_host.copy_(rd["y_pred"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        t_e2e_inf = timed(e2e_inf, a.steps, a.warmup, dist)

    # ---- per-kernel timing pass (CUDA events around every C-ABI call; not part of the headline numbers)
    # Serialised for this pass only: the DNN head normally runs on its own stream NEXT to the RAT encoder and the RAT-block
    # kernels start their prologue under the previous kernel (programmatic dependent launch); either would make an event
    # pair measure queueing / overlap instead of the entry point's own kernels.
    model.train()
    side_env = {k: os.environ.get(k) for k in ("RAT_DNN_SIDE",)}
    os.environ["RAT_DNN_SIDE"] = "0"
    rn.profile_calls(True)
    nprof = min(5, a.steps)
    for i in range(nprof):
        # keep the GPU busy (~25 ms spin) while the host enqueues the whole step, so that each event pair brackets the
        # kernels of one entry point back to back on the device and never the host's launch latency
        torch.cuda._sleep(50_000_000)
        model.train_step(dev_batches[a.warmup + i])
        torch.cuda.synchronize()
    prof = rn.profile_results()
    rn.profile_calls(False)
    for k, v in side_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    pk = peaks()
    total_ms = sum(ms for _, ms in prof.values())
    kernels = {k: {"calls_per_step": n // nprof, "ms_per_step": round(ms / nprof, 4), "share": round(ms / total_ms, 4)}
               for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    rows_tok = B * T * N
    I = H * 10

    def per_call_ms(name):
        n, ms = prof[name]
        return ms / n
    # dominant kernel: fused attention backward (2 per RAT block: intra S=N, cross S=T).  GEMM-shaped work on the tensor
    # cores (mma.sync fragments per warp + tcgen05 weight-gradient / dA products), so it is reported against the measured
    # dense tensor peak; the same launch is also shown against the HBM roofline (x + dout read, dx written, base = dout
    # hits L2), which is the roof that bounds it: the arithmetic intensity (165 FLOP/B) is below the ridge.
    attn_bwd_flops = rows_tok * (attn_flops_per_token(D, I, N) + attn_flops_per_token(D, I, T)) / 2 * 2.75
    ab_ms = per_call_ms("rat_attn_bwd")
    ach_tf = attn_bwd_flops / (ab_ms * 1e-3) / 1e12
    ab_bytes = rows_tok * D * 4 * 3
    tc_mode = a.precision == "fp16"
    rr = tc_mode and os.environ.get("RAT_RR", "1") != "0"
    roofline = {"kernel": ("k_attn_bwd_rr (+k_reduce_attn_tc)" if rr else "k_attn_bwd_tc (+k_reduce_attn_tc)") if tc_mode
                else "k_attn_bwd (+k_reduce_attn)",
                "bound": "tensor", "achieved": round(ach_tf, 3),
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(ach_tf / pk["tf_sustained"], 5),
                "traffic": ncu_traffic("attn_bwd", S, B, K) if tc_mode else None,
                "peak_source": pk["src"] + " bf16 dense (sustained)", "avg_launch_ms": round(ab_ms, 4),
                "flops_per_launch": attn_bwd_flops,
                "hbm_view": {"bytes_per_launch": ab_bytes, "achieved_gbs": round(ab_bytes / (ab_ms * 1e-3) / 1e9, 1),
                             "frac_of_hbm_peak": round(ab_bytes / (ab_ms * 1e-3) / 1e9 / pk["hbm"], 4)},
                "note": "average over the step's 8 launches (4 intra S=14, 3 cross S=6 on all tokens, 1 cross on field token 0 "
                        "only).  Algorithmic FLOPs = 2.75x forward (recompute + dgrad + wgrad).  Every product is a 16x16x(16..48) "
                        "GEMM: one warp owns a 16-row fragment tile end to end (mma.sync m16n8k16, 288 HMMA per tile at 8.1 "
                        "cycles each = 22 % of the legacy tensor pipe), weight gradients and dA on tcgen05; the kernel is bound by "
                        "instruction issue + shared-memory fragment loads + HMMA latency (~100 cycles), see "
                        "profiles/r02_attn_rr_ncu_full.txt and DESIGN.md section 3"}
    gb = B * gather_bytes_per_sample(K, L, F, D)
    g_step_ms = per_call_ms("rat_gather_fwd_sharded" if sharded else "rat_gather_fwd")
    # A 30 us kernel bracketed by its own event pair also pays ~4 us of event + launch turnaround on the device, so
    # the gather / scatter-reduce launch durations are taken from 24 back-to-back launches between ONE event pair
    # (rotating output blocks: 6 x 55 MB > L2); the single-launch in-step figure is reported next to it.
    eng = model._engine
    ws = eng._workspace(B, T, True)
    st_ = rn.current_stream()
    blocks = [torch.empty_like(ws["acts"][0]) for _ in range(6)]

    def loop_ms(fn, n=24):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    drop = float(eng.spec.emb_dropout)
    if sharded:
        g_ms = g_step_ms
    else:
        g_ms = loop_ms(lambda i: rn.call(
            "rat_gather_fwd", eng.store.emb_W, eng.store.lr_W, eng.p["label_embedding_layer.weight"], ws["ids"], ws["labels"],
            eng.col_off, eng.col_vocab, eng.field_col0, eng.field_width, blocks[i % 6], ws["x_emb"], ws["lr_out"], B, T, L,
            F, D, drop, eng.spec.seed, 7, eng.err_flag, st_))
    roofline_gather = {"kernel": "k_gather_flat" + (" (rows loaded from the owners' shards over NVLink peer pointers)" if sharded else ""), "bound": "hbm", "achieved": round(gb / (g_ms * 1e-3) / 1e9, 1),
                       "peak": pk["hbm"], "unit": "GB/s", "frac": round(gb / (g_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                       "traffic": None if sharded else ncu_traffic("gather", S, B, K),
                       "bytes_per_launch": gb, "peak_source": pk["src"],
                       "avg_launch_ms": round(g_ms, 4), "in_step_single_launch_ms": round(g_step_ms, 4)}
    # scatter: the critical-path call (segment scan + fix-ups; dropout backward fused).  Algorithmic bytes:
    # block gradient read once + sorted key/occurrence index + one gradient row per occurrence written... i.e.
    # B*T*N*D*4 (block grad) + B*T*(L+1)*8 (sorted keys, vals) + B*T*L*D*4 (per-occurrence row reads, sequence
    # columns re-read their token from L2).  The key build + radix sort (rat_emb_scatter_plan) only needs the ids and
    # runs on a side stream under the forward kernels; its duration is reported next to it.
    sc_step_ms = per_call_ms("rat_emb_scatter_reduce")
    sw = ws["scatter_ws"]
    if sharded:
        sc_ms = sc_step_ms
    else:
        gs_ = eng.store
        rn.call("rat_emb_scatter_plan", ws["ids"], ws["labels"], eng.col_off, eng.col_pad, eng.col_vocab, B, T, L, F, D,
                eng.spec.V, sw, sw.numel() * 4, st_)
        g_emb = gs_.G[gs_.emb_off:gs_.emb_off + eng.spec.V * D]
        g_lr = gs_.G[gs_.lr_off:gs_.lr_off + eng.spec.V] if eng.spec.use_wide else None
        sc_ms = loop_ms(lambda i: rn.call(
            "rat_emb_scatter_reduce", ws["ids"], ws["labels"], blocks[i % 6], ws["dxemb"], ws["dlogit"] if eng.spec.use_wide
            else None, eng.col_off, eng.col_pad, eng.col_vocab, eng.col_field, g_emb, g_lr,
            gs_.grad_views["label_embedding_layer.weight"], B, T, L, F, D, eng.spec.V, drop, eng.spec.seed, 7, 1, sw,
            sw.numel() * 4, st_))
        gs_.G.zero_()
    plan_ms = per_call_ms("rat_emb_scatter_plan") if "rat_emb_scatter_plan" in prof else 0.0
    sc_bytes = B * T * N * D * 4 + B * T * (L + 1) * 8 + B * T * L * D * 4
    roofline_scatter = {"kernel": "rat_emb_scatter_reduce (k_segment_scan + k_fixup_items)", "bound": "hbm",
                        "achieved": round(sc_bytes / (sc_ms * 1e-3) / 1e9, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(sc_bytes / (sc_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                        "traffic": ncu_traffic("scatter", S, B, K),
                        "bytes_per_launch": sc_bytes, "peak_source": pk["src"], "avg_launch_ms": round(sc_ms, 4),
                        "in_step_single_call_ms": round(sc_step_ms, 4),
                        "plan_ms_side_stream": round(plan_ms, 4),
                        "frac_incl_plan": round(sc_bytes / ((sc_ms + plan_ms) * 1e-3) / 1e9 / pk["hbm"], 4)}
    P = model._engine.store.total
    ad_ms = per_call_ms("rat_adam_step")
    roofline_adam = {"kernel": "k_adam", "bound": "hbm", "achieved": round(P * 32 / (ad_ms * 1e-3) / 1e9, 1),
                     "peak": pk["hbm"], "unit": "GB/s", "frac": round(P * 32 / (ad_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                     "traffic": None, "bytes_per_launch": P * 32, "peak_source": pk["src"],
                     "avg_launch_ms": round(ad_ms, 4),
                     "note": "W, G, M, V (P*32 B = 151 MB at kkbox) were just touched by the gradient-norm / scatter kernels and "
                             "largely sit in the 126 MB L2, so the effective rate can exceed the DRAM copy peak"}

    # free the headline model's buffers before the secondary workloads
    model._engine._ws.clear()
    model._engine._graphs.clear()
    del blocks, dev_batches, gen
    torch.cuda.empty_cache()
    secondary = {} if a.no_secondary else secondary_workloads(a, rank, world, local, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    gB = B * world
    line = {
        "metric": "RAT_m2 train samples/sec", "value": round(a.steps * gB / t_train, 1), "unit": "samples/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(t_train / a.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp16": "f16", "tf32": "tf32", "fp32": "f32"}[a.precision], "data": "synthetic",
        "config": {"workload": f"RAT_m2 {cfg['dataset_id']} shape, K={K}, B={B}/GPU, train step (fwd+bwd+clip+Adam)",
                   "precision": a.precision,
                   "arithmetic": {"fp16": "tcgen05 / mma.sync fp16 operands (10-bit mantissa) with fp32 accumulation and dynamic "
                                          "power-of-two gradient scaling; residual stream, LayerNorm, softmax statistics, GELU, "
                                          "BatchNorm, loss and optimizer in fp32",
                                  "tf32": "mma.sync TF32 operands, fp32 accumulation; everything else fp32",
                                  "fp32": "fp32 SIMT"}[a.precision],
                   "global_batch": gB, "topK": K, "fields": F, "input_length": L, "embedding_dim": D, "heads": H,
                   "params": n_params, "pool_rows": a.pool_rows, "parallelism": f"dp{world}" + ("+row-sharded tables (NVLink peer loads, reduce-scatter)" if sharded else ""),
                   "vocab_scale": a.vocab_scale,
                   "l2": "every step uses a new batch; per-step working set (~660 MB activations) exceeds the 126 MB L2"},
        "e2e": {"value": round(a.steps * gB / t_e2e, 1), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "api": "fuxictr.pytorch.models.RAT_m2.train_step(host f64 wire batch)"},
        "infer": {"value": round(a.steps * gB / t_inf, 1), "unit": "samples/s",
                  "ms_per_step": round(t_inf / a.steps * 1e3, 4),
                  "e2e": {"value": round(a.steps * gB / t_e2e_inf, 1), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                          "d2h_bytes_per_step": B * 4}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline, "roofline_gather": roofline_gather, "roofline_scatter": roofline_scatter,
        "roofline_adam": roofline_adam,
        "kernels": kernels,
        "kernels_note": "per-entry-point CUDA-event times of a separate, serialised pass (eager launches, DNN head on the main "
                        "stream); in the timed steps the whole step is one CUDA graph, the DNN head and the scatter plan run on "
                        "side streams next to the RAT-block kernels and those start their prologue under the previous kernel "
                        "(programmatic dependent launch), so the entries add up to more than ms_per_step",
        "reference_derived": {"note": "BASELINE.md derived (not published) reference-GPU numbers, unknown GPU, incl. dataloader",
                              "train_samples_per_s": {"kkbox": 8800, "ml": 52000, "tmall": 3300}[S],
                              "infer_samples_per_s": {"kkbox": 37500, "ml": 110000, "tmall": 22900}[S]},
    }
    line["secondary"] = secondary
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, S, K)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def small_workload(shape, B, K, precision, rank, world, local, dist, steps=6, warmup=3, sharded=False, vocab_scale=1.0,
                   pool_rows=262144, legs=("train", "infer"), variant="RAT_m2"):
    """train / inference samples/s of one secondary workload (device-resident inputs, CUDA events, max over ranks)."""
    import rat_native as rn
    from rat_native import shapes
    from rat_native.engine import set_precision
    from fuxictr.pytorch import models
    from fuxictr.pytorch.data_generator import DeviceDataGenerator
    set_precision(precision)
    fm = shapes.make_feature_map(shape, vocab_scale=vocab_scale)
    params = shapes.model_params(shape, K=K, gpu=local)
    if sharded:
        params["shard_embeddings"] = True
    os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
    model = getattr(models, variant)(fm, **params)
    if world > 1:
        eng = model._engine
        dist.broadcast(eng.store.W[:eng.store.emb_off] if sharded else eng.store.W, 0)
    pool = shapes.synthetic_array(fm.feature_specs, pool_rows, seed=77, pos_ratio=shapes.SHAPES[shape]["pos_ratio"])
    nbr = shapes.synthetic_neighbours(pool_rows, pool_rows, K, seed=77)
    gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B * world, shuffle=True, device=f"cuda:{local}", seed=77, rank=rank,
                              world=world, drop_last=True)
    it = iter(gen)
    batches = [next(it) for _ in range(min(len(gen), steps + warmup))]
    out = {"model": variant, "shape": shape, "B_per_gpu": B, "K": K, "precision": precision, "n_gpus": world,
           "params": model.count_parameters()}
    if "train" in legs:
        model.train()
        t = timed(lambda i: model.train_step(batches[i % len(batches)]), steps, warmup, dist)
        model._engine.check_errors()
        out["train_samples_per_s"] = round(steps * B * world / t, 1)
        out["train_ms_per_step"] = round(t / steps * 1e3, 4)
    if "infer" in legs:
        model.eval()
        with torch.no_grad():
            t = timed(lambda i: model.forward(batches[i % len(batches)]), steps, warmup, dist)
        out["infer_samples_per_s"] = round(steps * B * world / t, 1)
        out["infer_ms_per_step"] = round(t / steps * 1e3, 4)
    model._engine._ws.clear()
    model._engine._graphs.clear()
    del model, gen, batches
    import gc
    torch.cuda.synchronize()
    gc.collect()                # models sit in reference cycles: free their (symmetric-memory) buffers now, not inside a later capture
    torch.cuda.empty_cache()
    set_precision("fp16")
    return out


def secondary_workloads(a, rank, world, local, dist):
    """the other BASELINE.json configs, measured with the same code in the same run (rank 0 keeps the results)."""
    sec = {}
    # configs[1] in strict arithmetic: TF32 tensor-core operands (mma.sync), everything else fp32
    sec["strict"] = small_workload(a.shape, a.batch, a.topk, "tf32", rank, world, local, dist, steps=4, warmup=2, legs=("train",))
    sec["strict"]["note"] = "same workload as the headline line with precision=tf32 (mma.sync TF32 operands, fp32 accumulate)"
    # configs[0] / configs[2] shapes, replicated tables
    sec["shapes"] = [small_workload(sh, a.batch, a.topk, a.precision, rank, world, local, dist) for sh in ("ml", "tmall")
                     if sh != a.shape]
    # strong scaling: the reference's global batch of 4096 split over the ranks
    if world > 1 and a.batch % world == 0:
        sec["strong"] = small_workload(a.shape, a.batch // world, a.topk, a.precision, rank, world, local, dist, steps=10, warmup=4,
                                       legs=("train",))
        sec["strong"]["scaling"] = "strong"
        sec["strong"]["global_batch"] = a.batch
    # configs[2]: tmall shape, embedding / LR tables row-sharded over the ranks (NVLink peer loads + row-gradient exchange)
    if world > 1:
        sec["tmall_sharded"] = small_workload("tmall", a.batch, a.topk, a.precision, rank, world, local, dist, sharded=True,
                                              legs=("train",))
        sec["tmall_sharded_x50"] = small_workload("tmall", a.batch, a.topk, a.precision, rank, world, local, dist, sharded=True,
                                                  vocab_scale=50.0, legs=("train",))
    # configs[4]: K x B inference sweep on this many GPUs (data-parallel replicas; B is per GPU)
    sweep = []
    for K in (1, 5, 16, 64):
        for B in (256, 4096, 65536):
            if B * (K + 1) * 14 * 40 * 4 * 3 > 60e9:
                continue
            r = small_workload(a.shape, B, K, a.precision, rank, world, local, dist, steps=3 if B >= 65536 else 10, warmup=3,
                               pool_rows=max(262144, B * world), legs=("infer",))
            sweep.append({"K": K, "B_per_gpu": B, "infer_samples_per_s": r["infer_samples_per_s"],
                          "ms_per_batch": r["infer_ms_per_step"]})
    sec["infer_sweep"] = {"shape": a.shape, "n_gpus": world, "precision": a.precision, "grid": sweep}
    # configs[3]: the other model variants on the headline shape (RAT_m0: one flat 84-token sequence -> tile / mma.sync kernels;
    # RAT_m1: intra Transformer + cross Transformer on the pooled tokens; RAT_m3: parallel branches, head width 20)
    sec["variants"] = []
    for v in ("RAT_m0", "RAT_m1", "RAT_m3"):
        try:
            sec["variants"].append(small_workload(a.shape, a.batch, a.topk, a.precision, rank, world, local, dist, steps=4,
                                                  warmup=2, variant=v))
        except Exception as exc:            # a secondary line must never take the headline line down
            sec["variants"].append({"model": v, "error": repr(exc)[:300]})
    if rank == 0:
        sec["bm25"] = bm25_workload()
    return sec


def bm25_workload(N=1_000_000, Q=8192, C=13, K=5):
    """BM25 top-K pre-retrieval (the step BEFORE the training path) on one GPU: Q queries against an N-row
    pool of C categorical columns (kkbox-like id ranges), through the drop-in BM25_topk_retrieval_v4 (host IDF tables + H2D
    inside the timed region) and the device kernel alone.  Integer compare-scan: the unit is (query, db row) pairs per second."""
    import rat_native as rn
    from fuxictr.datasets.data_utils import BM25_topk_retrieval_v4, _dense_ids, _idf_tables, _idf_of_queries
    g = np.random.default_rng(0)
    ranges = [40, 6, 3000, 12, 900, 25, 7, 150, 60, 5, 2000, 33, 480][:C]
    db = np.stack([g.integers(0, r, N) for r in ranges], axis=1)
    qry = np.stack([g.integers(0, r, Q) for r in ranges], axis=1)
    t0 = time.perf_counter()
    res = BM25_topk_retrieval_v4(db, qry, topK=K)
    torch.cuda.synchronize()
    t_api = time.perf_counter() - t0
    dev = torch.device("cuda", torch.cuda.current_device())
    db_o, q_o = _dense_ids(db, qry)
    db_d, q_d = torch.from_numpy(db_o).to(dev), torch.from_numpy(q_o).to(dev)
    w_d = torch.from_numpy(_idf_of_queries(_idf_tables(db), qry)).to(dev)
    vals = torch.zeros(Q, K, dtype=torch.float64, device=dev); inds = torch.zeros(Q, K, dtype=torch.int64, device=dev)
    lens = torch.zeros(Q, dtype=torch.int64, device=dev)
    ws = torch.empty(int(rn.query("rat_bm25_topk_workspace_bytes", N, Q, K)) // 8 + 2, dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for i in range(3):
        e0.record()
        rn.call("rat_bm25_topk", db_d, N, q_d, w_d, Q, 0, C, K, 0, 0, vals, inds, lens, ws, ws.numel() * 8, rn.current_stream())
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    k_ms = min(ms)
    return {"workload": f"BM25 top-{K}: {Q} queries x {N} pool rows x {C} columns", "kernel_ms": round(k_ms, 3),
            "pairs_per_s": round(Q * N / (k_ms * 1e-3), 1), "column_compares_per_s": round(Q * N * C / (k_ms * 1e-3), 1),
            "api_seconds_incl_host_idf_and_copies": round(t_api, 3), "queries_per_s_api": round(Q / t_api, 1),
            "mean_len": float(res.lens.mean()),
            "bound": "instruction issue (1 compare + 1 select + 1 float64 add per column pair); db rows are staged once per 8 queries "
                     "in shared memory, so HBM traffic is N*C*4 B per 8 queries"}


# ----------------------------------------------------------------------------------------- CPU arms (oracle port)
# The reference is pure Python / PyTorch and cannot travel to the GPU box (/root/reference does not exist there), so the CPU
# arm is the oracle: the same torch-CPU fp32 arithmetic, pinned against the imported reference by tests/golden (BASELINE.md
# section 2 protocol: pre-collated synthetic batches of the SAME batch size, all host threads and one thread, a training
# step = zero_grad -> loss -> backward -> clip -> Adam, and an eval forward under no_grad; the retrieval-set assembly of
# Dataset.__getitem__ + collate is timed separately).
def _oracle_setup(S, K, Bc):
    from oracle import rat_oracle as O
    spec = O.shape_spec(S)
    params = O.init_params(spec, 0)
    bufs = O.init_buffers(spec)
    pool = O.synthetic_pool(spec, 20000, seed=1)
    nbr = O.synthetic_neighbours(Bc * 2, 20000, K, seed=1)
    return O, spec, params, bufs, pool, nbr


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _cpu_steps(O, spec, params, bufs, pool, nbr, Bc, threads, n_train, n_eval):
    """samples/s of n_train training steps and n_eval eval forwards at batch Bc on `threads` host threads (1 warm-up each)."""
    torch.set_num_threads(threads)
    st = O.AdamState()
    batches = []
    t0 = time.perf_counter()
    for j in range(2):
        X, y = O.assemble_batch(pool[j * Bc:(j + 1) * Bc], pool, nbr[j * Bc:(j + 1) * Bc], np.arange(Bc))
        batches.append((torch.from_numpy(X), torch.from_numpy(y)))
    t_asm = (time.perf_counter() - t0) / 2
    res = {"threads": threads, "batch": Bc, "assembly_samples_per_s": round(Bc / t_asm, 1)}
    if n_train:
        O.train_step(params, bufs, spec, st, *batches[0])
        t0 = time.perf_counter()
        for i in range(n_train):
            O.train_step(params, bufs, spec, st, *batches[(i + 1) % 2])
        res["train_samples_per_s"] = round(n_train * Bc / (time.perf_counter() - t0), 1)
    if n_eval:
        with torch.no_grad():
            O.forward(params, bufs, spec, *batches[0], training=False)
            t0 = time.perf_counter()
            for i in range(n_eval):
                O.forward(params, bufs, spec, *batches[(i + 1) % 2], training=False)
            res["infer_samples_per_s"] = round(n_eval * Bc / (time.perf_counter() - t0), 1)
    return res


def cpu_baseline(a, S, K):
    """the CPU arm on the GPU box's host cores, bounded to ~30 s: all threads at the GPU arm's batch size (train + eval), one
    thread on a smaller batch (a 1-thread step at B=4096 alone takes 40 s on the kkbox shape)."""
    Bc = a.cpu_batch
    nthr = os.cpu_count() or 1
    O, spec, params, bufs, pool, nbr = _oracle_setup(S, K, Bc)
    t0 = time.perf_counter()
    allt = _cpu_steps(O, spec, params, bufs, pool, nbr, Bc, nthr, 2, 2)
    one = _cpu_steps(O, spec, params, bufs, pool, nbr, min(Bc, 256), 1, 1, 1)
    torch.set_num_threads(nthr)
    t = time.perf_counter() - t0
    return {"value": allt["train_samples_per_s"], "unit": "samples/s", "cores": nthr, "kind": "port",
            "cpu_model": _cpu_model(), "torch": torch.__version__,
            "sample": f"oracle (torch-CPU fp32 restatement of the reference, pinned by tests/golden) on {nthr} host threads: 2 timed "
                      f"training steps + 2 timed eval forwards at B={Bc} (same batch size as the GPU arm), pre-collated batches; "
                      f"1-thread figures at B={min(Bc, 256)}; {t:.0f} s of CPU work",
            "all_threads": allt, "one_thread": one}


def run_reference(a):
    """--impl reference: the reference algorithm on the host cores (the Python reference cannot travel to the GPU box;
    the oracle port is the same torch-CPU arithmetic, pinned against it by tests/golden), same workload / batch size."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    S, K, Bc = a.shape, a.topk, a.cpu_batch
    nthr = os.cpu_count() or 1
    torch.set_num_threads(nthr)
    O, spec, params, bufs, pool, nbr = _oracle_setup(S, K, Bc)
    st = O.AdamState()
    # every step is one full batch of the GPU arm's workload (B=4096: 1.5 - 8 s of CPU work), so the step counts are
    # bounded to keep the run within a few minutes
    steps, warmup = max(1, min(a.steps, 6)), max(0, min(a.warmup, 1))
    batches = []
    for j in range(2):
        X, y = O.assemble_batch(pool[j * Bc:(j + 1) * Bc], pool, nbr[j * Bc:(j + 1) * Bc], np.arange(Bc))
        batches.append((torch.from_numpy(X), torch.from_numpy(y)))
    for i in range(warmup):
        O.train_step(params, bufs, spec, st, *batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        O.train_step(params, bufs, spec, st, *batches[(warmup + i) % 2])
    t = time.perf_counter() - t0
    v = round(steps * Bc / t, 1)
    with torch.no_grad():
        O.forward(params, bufs, spec, *batches[0], training=False)
        t1 = time.perf_counter()
        O.forward(params, bufs, spec, *batches[1], training=False)
        v_inf = round(Bc / (time.perf_counter() - t1), 1)
    from rat_native import shapes
    cfg = shapes.SHAPES[S]
    line = {"impl": "reference", "metric": "RAT_m2 train samples/sec", "value": v, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(t / steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"RAT_m2 {cfg['dataset_id']} shape, K={K}, B={a.batch}/GPU, train step (fwd+bwd+clip+Adam)",
                       "cpu_sample": f"each CPU step is one batch of B={Bc} samples (the GPU arm's batch size) on all {nthr} host threads; "
                                     f"step count bounded to {steps} (+{warmup} warm-up); fp32 torch-CPU oracle port of the reference"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": nthr, "kind": "port", "cpu_model": _cpu_model(),
                             "sample": f"{steps} oracle training steps at B={Bc} on {nthr} host threads",
                             "infer_samples_per_s": v_inf},
            "infer": {"value": v_inf, "unit": "samples/s"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
