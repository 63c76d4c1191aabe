"""Host-side engine of the RAT hot path: owns the flat parameter / gradient / Adam-state buffers in HBM and
sequences the C-ABI kernels of librat_b200.so for one forward, or one full training step.

PyTorch is used for device memory, streams and (data-parallel) torch.distributed plumbing only; every
arithmetic step of the path is a kernel of librat_b200.so.  There is no CPU / eager fallback.

HBM layout (DESIGN.md "Data layout"): ONE fp32 buffer `W` holds every parameter, `G`, `M`, `V` mirror it
(gradient, Adam moments).  [ net region | embedding-named region ]; the second region (label table, per-field
tables concatenated to emb_W [V_total, D], LR tables concatenated to lr_W [V_total]) gets the embedding
regulariser, exactly the `"embedding_layer" in name` rule of base_model.py:86.  state_dict tensors are views.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import call, current_stream, query, require_device


class _no_gc:
    """No automatic garbage collection while a CUDA graph is being captured: a collected engine of an EARLIER model frees
    its symmetric-memory buffers (cuMemUnmap / cudaFree), which CUDA forbids during a global-mode capture and which
    aborts the process from a destructor.  torch.cuda.graph.__enter__ runs gc.collect() itself before the capture starts."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()
        return False

EMB = "embedding_layer.embedding_layer.embedding_layer."
LRP = "lr_layer.embedding_layer.embedding_layer.embedding_layer."


@dataclass
class FeatureSpec:
    name: str
    type: str                   # categorical | sequence
    vocab_size: int
    max_len: int = 1
    padding_idx: Optional[int] = None

    @property
    def width(self):
        return self.max_len if self.type == "sequence" else 1

    @property
    def pad(self):
        return self.vocab_size - 1 if self.type == "sequence" else self.padding_idx


@dataclass
class EngineSpec:
    features: List[FeatureSpec]
    model: str = "RAT_m2"
    embedding_dim: int = 10
    num_heads: int = 1
    dim_head: int = 10
    scale_dim: int = 4
    depth: int = 4
    dnn_hidden_units: Sequence[int] = (64, 64, 64)
    batch_norm: bool = False
    use_wide: bool = False
    emb_dropout: float = 0.0
    net_dropout: float = 0.0
    embedding_regularizer: float = 0.0
    net_regularizer: float = 0.0
    learning_rate: float = 1e-3
    max_gradient_norm: float = 10.0
    seed: int = 2021
    shard_tables: bool = False      # row-shard the embedding / LR tables over the data-parallel ranks

    @property
    def F(self):
        return len(self.features)

    @property
    def L(self):
        return sum(f.width for f in self.features)

    @property
    def V(self):
        return sum(f.vocab_size for f in self.features)


_PREC = {"fp32": 0, "tf32": 1, "fp16": 2}


def set_precision(mode: str):
    """'fp16' (default): tcgen05 projections + DNN GEMMs with fp16 operands (10-bit mantissa), fp32 accumulation in TMEM and
    dynamic power-of-two gradient scaling; 'tf32': mma.sync TF32 projections; 'fp32': exact SIMT twin (parity anchor).
    Process-wide (the library holds one mode); models expose it as the `precision` keyword / config key."""
    if mode not in _PREC:
        raise ValueError(f"precision={mode!r}: choose one of {sorted(_PREC)} (there is no bf16 mode: bf16 operands missed "
                         f"the 1e-3 AUC bar, DESIGN.md section 1)")
    call("rat_set_precision", _PREC[mode])


def get_precision() -> str:
    return {0: "fp32", 1: "tf32", 2: "fp16"}[int(query("rat_get_precision"))]


def _align4(n):
    return (n + 3) // 4 * 4


def dnn_layout(spec) -> Tuple[List[Tuple[int, Optional[int]]], int]:
    """indices inside the reference's nn.Sequential `dnn.dnn` (deep.py:126-137)."""
    out, i = [], 0
    for _ in spec.dnn_hidden_units:
        lin, bn = i, None
        i += 1
        if spec.batch_norm:
            bn = i
            i += 1
        i += 1                      # activation
        if spec.net_dropout > 0:
            i += 1
        out.append((lin, bn))
    return out, i


def param_shapes(spec: EngineSpec) -> Tuple["OrderedDict[str, tuple]", "OrderedDict[str, tuple]"]:
    """(net params, embedding-named params) with the reference's state_dict names and shapes."""
    D, H, dh, F = spec.embedding_dim, spec.num_heads, spec.dim_head, spec.F
    I, M = H * dh, D * spec.scale_dim
    net: "OrderedDict[str, tuple]" = OrderedDict()
    emb: "OrderedDict[str, tuple]" = OrderedDict()
    emb["label_embedding_layer.weight"] = (3, D)
    for f in spec.features:
        emb[EMB + f.name + ".weight"] = (f.vocab_size, D)
    if spec.use_wide:
        for f in spec.features:
            emb[LRP + f.name + ".weight"] = (f.vocab_size, 1)
    net["query_proj.weight"] = (F * D, F * D)
    net["query_proj.bias"] = (F * D,)

    def attn(pre, qkv=True):
        net[pre + "norm.weight"] = (D,)
        net[pre + "norm.bias"] = (D,)
        if qkv:
            net[pre + "fn.to_qkv.weight"] = (3 * I, D)
        net[pre + "fn.to_out.0.weight"] = (D, I)
        net[pre + "fn.to_out.0.bias"] = (D,)

    def ff(pre):
        net[pre + "net.0.weight"] = (M, D)
        net[pre + "net.0.bias"] = (M,)
        net[pre + "net.3.weight"] = (D, M)
        net[pre + "net.3.bias"] = (D,)

    def transformer(pre):
        for l in range(spec.depth):
            attn(f"{pre}layers.{l}.0.")
            net[f"{pre}layers.{l}.1.norm.weight"] = (D,)
            net[f"{pre}layers.{l}.1.norm.bias"] = (D,)
            ff(f"{pre}layers.{l}.1.fn.")
        net[pre + "norm.weight"] = (D,)
        net[pre + "norm.bias"] = (D,)

    if spec.model == "RAT_m2":
        for l in range(spec.depth):
            attn(f"encoder.encoder.{l}.cross_attention.")
            attn(f"encoder.encoder.{l}.intra_attention.")
            ff(f"encoder.encoder.{l}.mlp.")
    elif spec.model == "RAT_m0":
        transformer("encoder.")
    elif spec.model == "RAT_m1":
        transformer("intra_transformer.")
        transformer("cross_transformer.")
    elif spec.model == "RAT_m3":
        for l in range(spec.depth):
            pre = f"encoder.encoder.{l}."
            for nm in ("W_q", "W_k_s", "W_v_s", "W_k_t", "W_v_t"):
                net[pre + nm + ".weight"] = (I, D)
            attn(pre + "intra_attention.", qkv=False)
            attn(pre + "cross_attention.", qkv=False)
            ff(pre + "mlp.")
    else:
        raise NotImplementedError(f"model={spec.model}")
    units = [F * D] + list(spec.dnn_hidden_units)
    layers, final = dnn_layout(spec)
    for (lin, bn), a, b in zip(layers, units[:-1], units[1:]):
        net[f"dnn.dnn.{lin}.weight"] = (b, a)
        net[f"dnn.dnn.{lin}.bias"] = (b,)
        if bn is not None:
            net[f"dnn.dnn.{bn}.weight"] = (b,)
            net[f"dnn.dnn.{bn}.bias"] = (b,)
    if len(spec.dnn_hidden_units) > 0:
        net[f"dnn.dnn.{final}.weight"] = (1, units[-1])
        net[f"dnn.dnn.{final}.bias"] = (1,)
    net["fc.weight"] = (1, D)
    net["fc.bias"] = (1,)
    return net, emb


def shard_rows(V: int, world: int) -> int:
    """rows per rank of the range partition of the concatenated table index space (multiple of 4)."""
    return _align4((V + world - 1) // world)


class ParamStore:
    """Flat fp32 parameter buffer + gradient + Adam moments, with named views.

    Replicated mode: [ net | label table | emb_W [V, D] | lr_W [V] ].
    Row-sharded mode (`shard=(rank, world)`): [ net | label table | emb shard [Vs, D] | lr shard [Vs] ], rank o owns
    global rows [o*Vs, (o+1)*Vs); W lives in symmetric memory so that every rank's gather kernel can load any
    shard over NVLink, and per-field tables exist only as (global row offset, vocab) ranges."""

    def __init__(self, spec: EngineSpec, device, shard=None):
        if shard is not None:
            self._init_sharded(spec, device, *shard)
            return
        self.shard = None
        net, emb = param_shapes(spec)
        self.offsets: "OrderedDict[str, Tuple[int, tuple]]" = OrderedDict()
        off = 0
        for k, shp in net.items():
            self.offsets[k] = (off, shp)
            off = _align4(off + math.prod(shp))
        self.net_end = off
        D, V = spec.embedding_dim, spec.V
        k = "label_embedding_layer.weight"
        self.offsets[k] = (off, emb[k])
        off = _align4(off + 3 * D)
        self.emb_off = off
        for f in spec.features:                      # contiguous -> emb_W [V, D]
            k = EMB + f.name + ".weight"
            self.offsets[k] = (off, emb[k])
            off += f.vocab_size * D
        off = _align4(off)
        self.lr_off = off
        if spec.use_wide:
            for f in spec.features:                  # contiguous -> lr_W [V]
                k = LRP + f.name + ".weight"
                self.offsets[k] = (off, emb[k])
                off += f.vocab_size
            off = _align4(off)
        self.total = off
        self.W = torch.zeros(off, dtype=torch.float32, device=device)
        self.G = torch.zeros(off, dtype=torch.float32, device=device)
        self.M = torch.zeros(off, dtype=torch.float32, device=device)
        self.Vv = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = OrderedDict((k, self.W[o:o + math.prod(s)].view(s)) for k, (o, s) in self.offsets.items())
        self.grad_views = OrderedDict((k, self.G[o:o + math.prod(s)].view(s)) for k, (o, s) in self.offsets.items())
        self.emb_W = self.W[self.emb_off:self.emb_off + V * D].view(V, D)
        self.lr_W = self.W[self.lr_off:self.lr_off + V] if spec.use_wide else None

    def _init_sharded(self, spec, device, rank, world):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        net, emb = param_shapes(spec)
        self.shard = (rank, world)
        self.offsets = OrderedDict()
        off = 0
        for k, shp in net.items():
            self.offsets[k] = (off, shp)
            off = _align4(off + math.prod(shp))
        self.net_end = off
        D, V = spec.embedding_dim, spec.V
        k = "label_embedding_layer.weight"
        self.offsets[k] = (off, emb[k])
        off = _align4(off + 3 * D)
        self.rows_per_shard = Vs = shard_rows(V, world)
        self.row0 = rank * Vs
        self.emb_off = off
        off = _align4(off + Vs * D)
        self.lr_off = off
        if spec.use_wide:
            off = _align4(off + Vs)
        self.total = off
        # per-field tables: global row ranges (field order = concatenation order)
        self.table_rows = OrderedDict()
        r = 0
        for f in spec.features:
            self.table_rows[f.name] = (r, f.vocab_size)
            r += f.vocab_size
        self.W = symm.empty(off, dtype=torch.float32, device=device)
        self.W.zero_()
        handle = symm.rendezvous(self.W, dist.group.WORLD.group_name)
        self.peer_ptrs = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=device)
        self._symm_handle = handle
        self.G = torch.zeros(off, dtype=torch.float32, device=device)
        self.M = torch.zeros(off, dtype=torch.float32, device=device)
        self.Vv = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = OrderedDict((k, self.W[o:o + math.prod(s)].view(s)) for k, (o, s) in self.offsets.items())
        self.grad_views = OrderedDict((k, self.G[o:o + math.prod(s)].view(s)) for k, (o, s) in self.offsets.items())
        self.emb_W = self.W[self.emb_off:self.emb_off + Vs * D].view(Vs, D)            # LOCAL shard
        self.lr_W = self.W[self.lr_off:self.lr_off + Vs] if spec.use_wide else None
        # local dense gradient over the padded global row space; reduce-scattered to the owners every step
        self.G_emb_full = torch.zeros(world * Vs * D, dtype=torch.float32, device=device)
        self.G_lr_full = torch.zeros(world * Vs, dtype=torch.float32, device=device) if spec.use_wide else None

    def numel_params(self):
        return sum(math.prod(s) for _, s in self.offsets.values())

    def g(self, name):
        return self.grad_views[name]

    def p(self, name):
        return self.views[name]


class RatEngine:
    def __init__(self, spec: EngineSpec, device="cuda:0"):
        require_device()
        self.spec = spec
        self.device = torch.device(device)
        shard = None
        if spec.shard_tables:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("shard_tables=True needs an initialised torch.distributed process group (NCCL)")
            shard = (dist.get_rank(), dist.get_world_size())
        self.store = ParamStore(spec, self.device, shard)
        self.p = self.store.views
        D = spec.embedding_dim
        # schema arrays
        col_off, col_vocab, col_pad, f_col0, f_w, col_field = [], [], [], [], [], []
        row = 0
        col = 0
        for fi, f in enumerate(spec.features):
            f_col0.append(col)
            f_w.append(f.width)
            for _ in range(f.width):
                col_off.append(row)
                col_vocab.append(f.vocab_size)
                col_pad.append(-1 if f.pad is None else f.pad)
                col_field.append(fi)
            row += f.vocab_size
            col += f.width
        mk = lambda v: torch.tensor(v, dtype=torch.int32, device=self.device)
        self.col_off, self.col_vocab, self.col_pad = mk(col_off), mk(col_vocab), mk(col_pad)
        self.field_col0, self.field_width, self.col_field = mk(f_col0), mk(f_w), mk(col_field)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        # BatchNorm buffers
        self.buffers: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        layers, _ = dnn_layout(spec)
        for (lin, bn), w in zip(layers, spec.dnn_hidden_units):
            if bn is not None:
                self.buffers[f"dnn.dnn.{bn}.running_mean"] = torch.zeros(w, device=self.device)
                self.buffers[f"dnn.dnn.{bn}.running_var"] = torch.ones(w, device=self.device)
                self.buffers[f"dnn.dnn.{bn}.num_batches_tracked"] = torch.zeros((), dtype=torch.long, device=self.device)
        self._ws: Dict[tuple, dict] = {}
        self.rng_step = 0
        nb = int(query("rat_optim_blocks"))
        self.opt_partial = torch.zeros(2 * nb, dtype=torch.float64, device=self.device)
        self.opt_state = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.lr = torch.full((1,), float(spec.learning_rate), dtype=torch.float32, device=self.device)
        self.world = 1
        self.dist_group = None
        self._side = None
        self.graph_inference = True         # CUDA-graph replay of the eval forward (see forward_ids)
        self.graph_training = True          # ... and of the whole training step (see train_step_ids)
        self.replayed_launches = 0          # kernels of librat_b200.so launched through graph replays (rat_launch_count
                                            # only sees direct C-ABI calls)
        self._graphs: Dict[tuple, object] = {}
        self.amax = torch.zeros(256, dtype=torch.float32, device=self.device)
        self.denc_amax = torch.zeros(1, dtype=torch.float32, device=self.device)     # written by rat_head every training step
        self._amax_next = 0
        self._amax_of = {}
        self._amax_on = False
        self.rank = 0
        self._os_peers = None               # one-shot all-reduce over NVLink peer memory (csrc/collective.cu)
        self._bnx_peers = None              # in-kernel exchange of the BatchNorm slab totals (csrc/mlp_fused.cu)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.world = dist.get_world_size()
                self.rank = dist.get_rank()
        except Exception:
            pass
        if self.world > 1:
            self._init_oneshot()

    def _init_oneshot(self):
        """symmetric buffer for the small-message all-reduce (raw BatchNorm sums, shard-norm partials); NCCL stays the
        fallback when symmetric memory is not available."""
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            n = (int(query("rat_oneshot_workspace_bytes", self.world)) + 3) // 4
            buf = symm.empty(n, dtype=torch.int32, device=self.device)
            buf.zero_()
            handle = symm.rendezvous(buf, dist.group.WORLD.group_name)
            torch.cuda.synchronize()
            dist.barrier()                  # every rank's buffer is zeroed before any peer stores into it
            self._os_buf, self._os_handle = buf, handle
            self._os_peers = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=self.device)
            n2 = (int(query("rat_bn_exchange_workspace_bytes", self.world)) + 3) // 4
            buf2 = symm.empty(n2, dtype=torch.int32, device=self.device)
            buf2.zero_()
            handle2 = symm.rendezvous(buf2, dist.group.WORLD.group_name)
            torch.cuda.synchronize()
            dist.barrier()
            self._bnx_buf, self._bnx_handle = buf2, handle2
            self._bnx_peers = torch.tensor([int(p) for p in handle2.buffer_ptrs], dtype=torch.int64, device=self.device)
        except Exception as exc:            # pragma: no cover - depends on the platform
            import logging
            logging.warning("one-shot all-reduce unavailable (%s): using NCCL for the BatchNorm sums", exc)
            self._os_peers = None
            self._bnx_peers = None

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, B: int, T: int, training: bool) -> dict:
        key = (B, T, training)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        s, dev = self.spec, self.device
        N, D, F, L = s.F + 1, s.embedding_dim, s.F, s.L
        f32 = dict(dtype=torch.float32, device=dev)
        ws = dict(
            ids=torch.empty(B, T, L, dtype=torch.int32, device=dev),
            labels=torch.empty(B, T, dtype=torch.int32, device=dev),
            y_true=torch.empty(B, **f32),
            x_emb=torch.empty(B, F * D, **f32),
            lr_out=torch.empty(B, **f32) if s.use_wide else None,
            y_pred=torch.empty(B, **f32),
            loss_part=torch.empty(2 * int(query("rat_head_blocks", B)), dtype=torch.float64, device=dev),
            loss=torch.zeros(2, **f32),
            dnn_out=torch.empty(B, **f32) if len(s.dnn_hidden_units) else None,
        )
        n_act = self._num_acts() if (training or s.model in ("RAT_m0", "RAT_m1")) else 3
        ws["acts"] = [torch.empty(B, T, N, D, **f32) for _ in range(n_act)]
        ws["acts_c"] = [torch.empty(B, T, D, **f32) for _ in range(n_act)] if s.model == "RAT_m1" else None
        ws["enc_stride"] = T * D if s.model in ("RAT_m1", "RAT_m2") else T * N * D
        # RAT_m2, last block: only token (t=0, n=0) of the encoder output is consumed (RAT_m2.py:138-140), so its cross
        # attention runs on the B sequences of field token 0 and its FeedForward on those B*T rows (dead-token
        # elimination; the reference computes all B*T*N rows and drops them)
        ws["last_c"] = [torch.empty(B, T, D, **f32) for _ in range(3)] if s.model == "RAT_m2" else None
        units = list(s.dnn_hidden_units)
        ws["z"] = [torch.empty(B, u, **f32) for u in units]
        ws["h"] = [torch.empty(B, u, **f32) for u in units]
        ws["bn_mean"] = [torch.empty(u, **f32) for u in units]
        ws["bn_rstd"] = [torch.empty(u, **f32) for u in units]
        ws["bn_sums"] = [torch.empty(2 * u, dtype=torch.float64, device=dev) for u in units]
        nbytes = 0
        dims = [F * D] + units
        for a, b in zip(dims[:-1], dims[1:]):
            nbytes = max(nbytes, int(query("rat_sgemm_workspace_bytes", B, b, a)),
                         int(query("rat_sgemm_workspace_bytes", b, a, B)), int(query("rat_sgemm_workspace_bytes", B, a, b)))
        ws["gemm_ws"] = torch.empty(max(nbytes // 4, 4), **f32)
        if training:
            H, dh, M = s.num_heads, s.dim_head, D * s.scale_dim
            ws["dact"] = torch.empty(B, T, N, D, **f32)
            ws["dact2"] = torch.empty(B, T, N, D, **f32) if s.model == "RAT_m3" else None
            ws["dlogit"] = torch.empty(B, **f32)
            ws["dact_c"] = torch.empty(B, T, D, **f32) if s.model in ("RAT_m1", "RAT_m2") else None
            ws["denc"] = ws["dact_c"] if s.model in ("RAT_m1", "RAT_m2") else ws["dact"]
            ws["dxemb"] = torch.zeros(B, F * D, **f32)
            ws["dh"] = [torch.empty(B, u, **f32) for u in units]
            hh, dd = (max(1, int(H / 2)), (H * dh) // max(1, int(H / 2))) if s.model == "RAT_m3" else (H, dh)
            shapes = {"RAT_m0": [(B, 1, T * N, 0)], "RAT_m1": [(B, T, N, 0), (B, 1, T, 0)]}.get(
                s.model, [(B, T, N, 0), (B, T, N, 1)])
            sizes = [int(query("rat_attn_bwd_workspace_bytes", b_, t_, n_, D, hh, dd, m_)) for b_, t_, n_, m_ in shapes]
            if min(sizes) == 0:
                raise RuntimeError("RAT attention backward: a sequence tile does not fit in shared memory")
            nb = max(max(sizes),
                     int(query("rat_ff_bwd_workspace_bytes", B * T * N, D, M)),
                     int(query("rat_layernorm_bwd_workspace_bytes", B * T * N, D)))
            if nb == 0:
                raise RuntimeError("RAT backward kernels: tile does not fit in shared memory for this shape")
            ws["bwd_ws"] = torch.empty(nb // 4 + 4, **f32)
            if self._defer_reduce():                               # second record buffer: deferred reductions (_bwd_ws)
                ws["bwd_ws2"] = torch.empty(nb // 4 + 4, **f32)
            sb = int(query("rat_emb_scatter_workspace_bytes", B * T * (L + 1), D))
            ws["scatter_ws"] = torch.empty(sb // 4 + 4, dtype=torch.int32, device=dev)
        self._ws[key] = ws
        return ws

    def _num_acts(self) -> int:
        s = self.spec
        if s.model in ("RAT_m2", "RAT_m3"):
            return 3 * s.depth + 1
        return 2 * s.depth + 2          # RAT_m0 / RAT_m1 Transformer: (attn, ff) per layer + final LayerNorm

    # ------------------------------------------------------------------ inputs
    def load_wire(self, X: torch.Tensor, y: torch.Tensor, training: bool):
        """X [B,T,L] float64 (device), y [B,T] float64 (device) -- the reference wire format."""
        B, T, L = X.shape
        assert L == self.spec.L, f"input_length {L} != schema {self.spec.L}"
        ws = self._workspace(B, T, training)
        self.err_flag.zero_()
        call("rat_convert_wire_f64", X, y, ws["ids"], ws["labels"], ws["y_true"], B, T, L, current_stream())
        return ws

    # ------------------------------------------------------------------ forward
    def _attn(self, x, res, out, pre, mode, B, T, N, alpha=1.0, heads=None, dh=None, wq=None, wk=None, wv=None):
        s, p = self.spec, self.p
        H = s.num_heads if heads is None else heads
        d_h = s.dim_head if dh is None else dh
        I = s.num_heads * s.dim_head
        if wq is None:
            w = p[pre + "fn.to_qkv.weight"]
            wq, wk, wv = w[:I], w[I:2 * I], w[2 * I:]
        call("rat_attn_fwd", x, res, out, p[pre + "norm.weight"], p[pre + "norm.bias"], wq, wk, wv,
             p[pre + "fn.to_out.0.weight"], p[pre + "fn.to_out.0.bias"], B, T, N, s.embedding_dim, H, d_h,
             float(s.dim_head ** -0.5), float(alpha), mode, current_stream())

    def _ff(self, x, res, out, pre, rows, ln=None):
        s, p = self.spec, self.p
        D, M = s.embedding_dim, s.embedding_dim * s.scale_dim
        lw = p[ln + "weight"] if ln else None
        lb = p[ln + "bias"] if ln else None
        call("rat_ff_fwd", x, res, out, lw, lb, p[pre + "net.0.weight"], p[pre + "net.0.bias"],
             p[pre + "net.3.weight"], p[pre + "net.3.bias"], rows, D, M, current_stream())

    def encode(self, ws, B, T, training):
        """block acts[0] -> final activations; returns the tensor whose [b,0,0,:] is the pooled token."""
        s = self.spec
        N = s.F + 1
        acts = ws["acts"]
        rows = B * T * N
        if s.model == "RAT_m2":
            cur = 0
            for l in range(s.depth):
                pre = f"encoder.encoder.{l}."
                if training:
                    i0, i1, i2, i3 = 3 * l, 3 * l + 1, 3 * l + 2, 3 * l + 3
                else:
                    i0, i1, i2, i3 = cur, (cur + 1) % 3, (cur + 2) % 3, cur
                self._attn(acts[i0], acts[i0], acts[i1], pre + "intra_attention.", 0, B, T, N)
                if l == s.depth - 1:            # last block: field token 0 only (see _workspace)
                    c = ws["last_c"]
                    call("rat_strided_copy", acts[i1], c[0], B * T, s.embedding_dim, N * s.embedding_dim, s.embedding_dim,
                         current_stream())
                    self._attn(c[0], c[0], c[1], pre + "cross_attention.", 1, B, T, 1)
                    self._ff(c[1], c[1], c[2], pre + "mlp.", B * T)
                    return c[2]
                self._attn(acts[i1], acts[i1], acts[i2], pre + "cross_attention.", 1, B, T, N)
                self._ff(acts[i2], acts[i2], acts[i3], pre + "mlp.", rows)
                cur = i3
            return acts[cur]
        if s.model == "RAT_m3":
            cur = 0
            h2 = max(1, int(s.num_heads / 2))
            dh2 = (s.num_heads * s.dim_head) // h2
            for l in range(s.depth):
                pre = f"encoder.encoder.{l}."
                if training:
                    i0, i1, i2, i3 = 3 * l, 3 * l + 1, 3 * l + 2, 3 * l + 3
                else:
                    i0, i1, i2, i3 = cur, (cur + 1) % 3, (cur + 2) % 3, (cur + 1) % 3
                p = self.p
                self._attn(acts[i0], None, acts[i1], pre + "intra_attention.", 0, B, T, N, 0.5, h2, dh2,
                           p[pre + "W_q.weight"], p[pre + "W_k_s.weight"], p[pre + "W_v_s.weight"])
                self._attn(acts[i0], acts[i1], acts[i2], pre + "cross_attention.", 1, B, T, N, 0.5, h2, dh2,
                           p[pre + "W_q.weight"], p[pre + "W_k_t.weight"], p[pre + "W_v_t.weight"])
                self._ff(acts[i2], acts[i0], acts[i3], pre + "mlp.", rows)
                cur = i3
            return acts[cur]
        if s.model == "RAT_m0":         # one flat sequence of T*N tokens per sample (RAT_m0.py:123-127)
            return self._transformer_fwd(acts, "encoder.", B, 1, T * N)
        if s.model == "RAT_m1":         # intra Transformer -> token 0 of every row -> cross Transformer
            z = self._transformer_fwd(acts, "intra_transformer.", B, T, N)
            c = ws["acts_c"]
            call("rat_strided_copy", z, c[0], B * T, s.embedding_dim, N * s.embedding_dim, s.embedding_dim,
                 current_stream())
            return self._transformer_fwd(c, "cross_transformer.", B, 1, T)
        raise NotImplementedError(s.model)

    def _transformer_fwd(self, a, prefix, B, T, N):
        """vit-style pre-norm Transformer (RAT_m0.py:193-208) over [B,T,N,D] with sequences = rows of N tokens."""
        s = self.spec
        rows = B * T * N
        for l in range(s.depth):
            self._attn(a[2 * l], a[2 * l], a[2 * l + 1], f"{prefix}layers.{l}.0.", 0, B, T, N)
            self._ff(a[2 * l + 1], a[2 * l + 1], a[2 * l + 2], f"{prefix}layers.{l}.1.fn.", rows,
                     ln=f"{prefix}layers.{l}.1.norm.")
        out = a[2 * s.depth + 1]
        call("rat_layernorm_fwd", a[2 * s.depth], out, self.p[prefix + "norm.weight"], self.p[prefix + "norm.bias"],
             rows, s.embedding_dim, current_stream())
        return out

    def _transformer_bwd(self, ws, a, d, prefix, B, T, N):
        s, g = self.spec, self.store.grad_views
        rows = B * T * N
        bw = self._bwd_ws(ws)
        call("rat_layernorm_bwd", a[2 * s.depth], d, d, self.p[prefix + "norm.weight"], g[prefix + "norm.weight"],
             g[prefix + "norm.bias"], rows, s.embedding_dim, bw, bw.numel() * 4, current_stream())
        self._amax_forget(d)                     # written by a kernel that does not publish max|dx|
        for l in reversed(range(s.depth)):
            self._ff_bwd(ws, a[2 * l + 1], d, d, d, f"{prefix}layers.{l}.1.fn.", rows, ln=f"{prefix}layers.{l}.1.norm.")
            self._attn_bwd(ws, a[2 * l], d, d, d, f"{prefix}layers.{l}.0.", 0, B, T, N)

    def _dnn_forward(self, ws, B, training):
        s, p, st = self.spec, self.p, current_stream()
        layers, final = dnn_layout(s)
        h = ws["x_emb"]
        K = s.F * s.embedding_dim
        gw = ws["gemm_ws"]
        for li, ((lin, bn), width) in enumerate(zip(layers, s.dnn_hidden_units)):
            z, out = ws["z"][li], ws["h"][li]
            call("rat_sgemm", h, p[f"dnn.dnn.{lin}.weight"], z, p[f"dnn.dnn.{lin}.bias"], B, width, K, K, K, width,
                 0, 0, gw, gw.numel() * 4, st)
            mean = rstd = gamma = beta = None
            if bn is not None:
                mean, rstd = ws["bn_mean"][li], ws["bn_rstd"][li]
                gamma, beta = p[f"dnn.dnn.{bn}.weight"], p[f"dnn.dnn.{bn}.bias"]
                rm, rv = self.buffers[f"dnn.dnn.{bn}.running_mean"], self.buffers[f"dnn.dnn.{bn}.running_var"]
                if training and self._dnn_fused():     # one cluster kernel: statistics + running stats + apply
                    call("rat_bn_act_fwd_train", z, B, width, gamma, beta, mean, rstd, rm, rv, 0.1, 1e-5, out,
                         float(s.net_dropout), s.seed, self._rng_stream(16 + li), self._bnx_peers, self.rank,
                         self.world, st)
                    self.buffers[f"dnn.dnn.{bn}.num_batches_tracked"] += 1
                    h, K = out, width
                    continue
                if training:
                    call("rat_bn_sums", z, B, width, ws["bn_sums"][li], st)
                    count = float(B) * self._allreduce_sums(ws["bn_sums"][li])
                    call("rat_bn_finalize", ws["bn_sums"][li], count, width, mean, rstd, rm, rv, 0.1, 1e-5, st)
                    self.buffers[f"dnn.dnn.{bn}.num_batches_tracked"] += 1
                else:
                    call("rat_bn_eval_stats", rm, rv, width, mean, rstd, 1e-5, st)
            drop = s.net_dropout if training else 0.0
            call("rat_bn_act_fwd", z, mean, rstd, gamma, beta, out, B, width, float(drop), s.seed,
                 self._rng_stream(16 + li), st)
            h, K = out, width
        call("rat_sgemm", h, p[f"dnn.dnn.{final}.weight"], ws["dnn_out"], p[f"dnn.dnn.{final}.bias"], B, 1, K, K, K, 1,
             0, 0, gw, gw.numel() * 4, st)

    def _dnn_fused(self) -> bool:
        """single-launch BatchNorm kernels (csrc/mlp_fused.cu); data-parallel: the slab totals are exchanged over NVLink
        peer memory inside the launch, which needs the symmetric exchange buffer."""
        return (self.world == 1 or self._bnx_peers is not None) and os.environ.get("RAT_DNN_FUSED", "1") != "0"

    def _allreduce_sums(self, t) -> int:
        """data-parallel hook (SyncBN-equivalent): all-reduce raw BN sums; returns the world size."""
        if self.world > 1:
            if self._os_peers is not None and t.dtype == torch.float64 and t.numel() <= 2048:
                call("rat_oneshot_allreduce_f64", self._os_peers, self.rank, self.world, t, t.numel(), t, current_stream())
            else:
                import torch.distributed as dist
                dist.all_reduce(t, group=self.dist_group)
        return self.world

    def _rng_stream(self, slot: int) -> int:
        """dropout stream of a call site; the training-step counter is device-resident (rat_rng_step_advance), so the
        value passed to the kernels is step-independent and a CUDA graph of the step replays with fresh masks."""
        return slot

    def forward_ids(self, ws, B, T, training=False, with_loss=False, inv_count=None):
        """ids/labels already in ws -> y_pred [B]; keeps activations in ws when training.

        Inference (eval, no loss) is launch-bound at small batches (~25 kernels of a few microseconds each), so the
        second call with the same (B, T, precision) captures the kernel sequence into a CUDA graph and later calls
        replay it; every buffer the kernels touch (workspace, flat parameter buffer, BN buffers) is persistent."""
        import rat_native as _rn
        if (self.graph_inference and not training and not with_loss and self.store.shard is None
                and _rn._profile is None and not torch.cuda.is_current_stream_capturing()):
            key = (B, T, int(query("rat_get_precision")))
            entry = self._graphs.get(key)
            if entry is None or entry[1] is not ws:         # first call (or a new workspace): eager, warms up lazy state
                self._graphs[key] = ("warm", ws)
            else:
                if entry[0] == "warm":
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    n0 = int(query("rat_launch_count"))
                    with _no_gc(), torch.cuda.graph(g):
                        self._forward_ids_impl(ws, B, T, False, False, None)
                    entry = (g, ws, int(query("rat_launch_count")) - n0)     # the graph keeps its workspace alive
                    self._graphs[key] = entry
                entry[0].replay()
                self.replayed_launches += entry[2]
                return ws["y_pred"]
        return self._forward_ids_impl(ws, B, T, training, with_loss, inv_count)

    def _forward_ids_impl(self, ws, B, T, training, with_loss, inv_count):
        s, st = self.spec, current_stream()
        D, F, L = s.embedding_dim, s.F, s.L
        drop = s.emb_dropout if training else 0.0
        if self.store.shard is None:
            call("rat_gather_fwd", self.store.emb_W, self.store.lr_W, self.p["label_embedding_layer.weight"], ws["ids"],
                 ws["labels"], self.col_off, self.col_vocab, self.field_col0, self.field_width, ws["acts"][0],
                 ws["x_emb"], ws["lr_out"], B, T, L, F, D, float(drop), s.seed, self._rng_stream(0), self.err_flag, st)
        else:       # rows are loaded straight from the owners' shards over NVLink (peer pointers)
            gs = self.store
            call("rat_gather_fwd_sharded", gs.peer_ptrs, gs.emb_off, gs.lr_off, gs.rows_per_shard, gs.shard[1],
                 self.p["label_embedding_layer.weight"], ws["ids"], ws["labels"], self.col_off, self.col_vocab,
                 self.field_col0, self.field_width, ws["acts"][0], ws["x_emb"], ws["lr_out"], B, T, L, F, D,
                 float(drop), s.seed, self._rng_stream(0), self.err_flag, st)
        if training:                # after the gather is queued: the sort overlaps the RAT-block kernels, not the gather
            self._plan_scatter(ws, B, T)
        dnn_done = None
        if len(s.dnn_hidden_units) and self._dnn_on_side_stream():
            side2, main = self._side_stream(1), torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)                          # x_emb is written by the gather
            side2.wait_event(ready)
            with torch.cuda.stream(side2):
                self._dnn_forward(ws, B, training)
                dnn_done = torch.cuda.Event()
                dnn_done.record(side2)
        enc = self.encode(ws, B, T, training)
        ws["enc_out"] = enc
        if dnn_done is not None:
            torch.cuda.current_stream().wait_event(dnn_done)
        elif len(s.dnn_hidden_units):
            self._dnn_forward(ws, B, training)
        N = F + 1
        want_loss = with_loss or training
        call("rat_head", enc, ws["enc_stride"], self.p["fc.weight"], self.p["fc.bias"], ws["dnn_out"], ws["lr_out"],
             ws["y_true"] if want_loss else None, B, D, ws["y_pred"], ws.get("dlogit") if training else None,
             ws.get("denc") if training else None, float(inv_count if inv_count else 1.0 / B),
             ws["loss_part"] if want_loss else None, ws["loss"][0:1] if want_loss else None,
             ws["loss"][1:2] if want_loss else None, self.denc_amax if training else None, st)
        return ws["y_pred"]

    def check_errors(self):
        flag = int(self.err_flag.item())
        if flag:
            raise RuntimeError(f"RAT gather: invalid input (flag={flag}: 1=id out of vocabulary, "
                               f"2=neighbour index out of range, 4=label not in {{0,1,2}})")


    # ------------------------------------------------------------------ backward
    # Dynamic gradient scaling (fp16 tensor-core mode): every backward kernel publishes max|dx| into a device slot and
    # the kernel consuming that tensor reads it (include/rat_b200.h, K5).  Slots are zeroed once per step.
    def _amax_reset(self):
        self._amax_on = int(query("rat_get_precision")) == 2
        if self._amax_on:
            self.amax.zero_()
        self._amax_next = 0
        self._amax_of = {}

    def _amax_new(self):
        if not self._amax_on:
            return None
        i = self._amax_next
        if i >= self.amax.numel():
            raise RuntimeError(f"gradient-scale slots exhausted ({self.amax.numel()}): depth too large for the amax buffer")
        self._amax_next += 1
        return self.amax[i:i + 1]

    def _amax_in(self, t):
        """slot holding max|t|; measured with rat_absmax when the last writer of t did not publish it."""
        if not self._amax_on:
            return None
        slot = self._amax_of.get(t.data_ptr())
        if slot is None:
            slot = self._amax_new()
            D = t.shape[-1]
            call("rat_absmax", t, t.numel() // D, D, D, slot, current_stream())
            self._amax_of[t.data_ptr()] = slot
        return slot

    def _amax_out(self, t):
        if not self._amax_on:
            return None
        slot = self._amax_new()
        self._amax_of[t.data_ptr()] = slot
        return slot

    def _amax_forget(self, t):
        self._amax_of.pop(t.data_ptr(), None)

    def _bwd_ws(self, ws):
        """record workspace of the next backward kernel.  With deferred reductions (rat_set_reduce_stream) two buffers
        alternate: the reduction of call i reads one while the kernel of call i+1 writes the other."""
        self._bwd_flip = not getattr(self, "_bwd_flip", False)
        alt = ws.get("bwd_ws2")
        return alt if (alt is not None and self._bwd_flip and self._defer_reduce()) else ws["bwd_ws"]

    def _defer_reduce(self) -> bool:
        return os.environ.get("RAT_DEFER_REDUCE", "0") == "1"     # measured: no gain, off

    def _dnn_on_side_stream(self) -> bool:
        """run the DNN head (independent of the RAT encoder between the gather and the logit) on its own stream.
        Needs the single-launch BatchNorm kernels (the split ones share a library-internal scratch buffer)."""
        return self._dnn_fused() and os.environ.get("RAT_DNN_SIDE", "1") != "0"

    def _side_stream(self, which: int):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        if which == 0:
            return self._side
        if getattr(self, "_side2", None) is None:
            self._side2 = torch.cuda.Stream(device=self.device)
        return self._side2

    def _attn_bwd(self, ws, x, d_in, base, d_out, pre, mode, B, T, N, alpha=1.0, heads=None, dh=None,
                  wq=None, wk=None, wv=None, names=None, acc_wq=0):
        s, p, g = self.spec, self.p, self.store.grad_views
        H = s.num_heads if heads is None else heads
        d_h = s.dim_head if dh is None else dh
        I = s.num_heads * s.dim_head
        if wq is None:
            w, gw = p[pre + "fn.to_qkv.weight"], g[pre + "fn.to_qkv.weight"]
            wq, wk, wv = w[:I], w[I:2 * I], w[2 * I:]
            gq, gk, gv = gw[:I], gw[I:2 * I], gw[2 * I:]
        else:
            gq, gk, gv = (g[n] for n in names)
        bw = self._bwd_ws(ws)
        call("rat_attn_bwd", x, d_in, base, d_out, p[pre + "norm.weight"], p[pre + "norm.bias"], wq, wk, wv,
             p[pre + "fn.to_out.0.weight"], gq, gk, gv, g[pre + "fn.to_out.0.weight"], g[pre + "fn.to_out.0.bias"],
             g[pre + "norm.weight"], g[pre + "norm.bias"], acc_wq, B, T, N, s.embedding_dim, H, d_h,
             float(s.dim_head ** -0.5), float(alpha), mode, self._amax_in(d_in), self._amax_out(d_out), bw,
             bw.numel() * 4, current_stream())

    def _ff_bwd(self, ws, x, d_in, base, d_out, pre, rows, ln=None):
        s, p, g = self.spec, self.p, self.store.grad_views
        D, M = s.embedding_dim, s.embedding_dim * s.scale_dim
        bw = self._bwd_ws(ws)
        call("rat_ff_bwd", x, d_in, base, d_out, p[ln + "weight"] if ln else None, p[ln + "bias"] if ln else None,
             p[pre + "net.0.weight"], p[pre + "net.0.bias"], p[pre + "net.3.weight"], g[pre + "net.0.weight"],
             g[pre + "net.0.bias"], g[pre + "net.3.weight"], g[pre + "net.3.bias"], g[ln + "weight"] if ln else None,
             g[ln + "bias"] if ln else None, rows, D, M, self._amax_in(d_in), self._amax_out(d_out), bw,
             bw.numel() * 4, current_stream())

    def encode_backward(self, ws, B, T):
        """ws['dact'] holds d(loss)/d(encoder output); on return it holds d(loss)/d(block after dropout)."""
        s = self.spec
        N = s.F + 1
        acts, d = ws["acts"], ws["dact"]
        rows = B * T * N
        if s.model == "RAT_m2":
            for l in reversed(range(s.depth)):
                pre = f"encoder.encoder.{l}."
                if l == s.depth - 1:            # last block ran on field token 0 only: dc [B,T,D] -> d[b,t,0,:], zeros elsewhere
                    c, dc = ws["last_c"], ws["dact_c"]
                    self._ff_bwd(ws, c[1], dc, dc, dc, pre + "mlp.", B * T)
                    self._attn_bwd(ws, c[0], dc, dc, dc, pre + "cross_attention.", 1, B, T, 1)
                    call("rat_expand_rows", dc, d, B * T, s.embedding_dim, N, current_stream())
                    if dc.data_ptr() in self._amax_of:
                        self._amax_of[d.data_ptr()] = self._amax_of[dc.data_ptr()]
                else:
                    self._ff_bwd(ws, acts[3 * l + 2], d, d, d, pre + "mlp.", rows)
                    self._attn_bwd(ws, acts[3 * l + 1], d, d, d, pre + "cross_attention.", 1, B, T, N)
                self._attn_bwd(ws, acts[3 * l], d, d, d, pre + "intra_attention.", 0, B, T, N)
            return d
        if s.model == "RAT_m3":
            d2 = ws["dact2"]
            h2 = max(1, int(s.num_heads / 2))
            dh2 = (s.num_heads * s.dim_head) // h2
            for l in reversed(range(s.depth)):
                pre = f"encoder.encoder.{l}."
                # x_out = x + FF(u), u = .5 A_s(x) + .5 A_t(x):  du = FFbwd(d) ; dx = d + .5 dA_t(du) + .5 dA_s(du)
                self._ff_bwd(ws, acts[3 * l + 2], d, None, d2, pre + "mlp.", rows)
                self._attn_bwd(ws, acts[3 * l], d2, d, d, pre + "cross_attention.", 1, B, T, N, 0.5, h2, dh2,
                               self.p[pre + "W_q.weight"], self.p[pre + "W_k_t.weight"], self.p[pre + "W_v_t.weight"],
                               (pre + "W_q.weight", pre + "W_k_t.weight", pre + "W_v_t.weight"), 0)
                self._attn_bwd(ws, acts[3 * l], d2, d, d, pre + "intra_attention.", 0, B, T, N, 0.5, h2, dh2,
                               self.p[pre + "W_q.weight"], self.p[pre + "W_k_s.weight"], self.p[pre + "W_v_s.weight"],
                               (pre + "W_q.weight", pre + "W_k_s.weight", pre + "W_v_s.weight"), 1)
            return d
        if s.model == "RAT_m0":
            self._transformer_bwd(ws, acts, d, "encoder.", B, 1, T * N)
            return d
        if s.model == "RAT_m1":
            dc = ws["dact_c"]
            self._transformer_bwd(ws, ws["acts_c"], dc, "cross_transformer.", B, 1, T)
            call("rat_strided_copy", dc, d, B * T, s.embedding_dim, s.embedding_dim, N * s.embedding_dim,
                 current_stream())          # d was zeroed by train_step_ids: only token 0 of each row gets gradient
            self._amax_forget(d)
            self._transformer_bwd(ws, acts, d, "intra_transformer.", B, T, N)
            return d
        raise NotImplementedError(s.model)

    def _dnn_backward(self, ws, B):
        s, p, g, st = self.spec, self.p, self.store.grad_views, current_stream()
        layers, final = dnn_layout(s)
        units = list(s.dnn_hidden_units)
        gw = ws["gemm_ws"]
        gwb = gw.numel() * 4
        # the final Linear(K -> 1) is handled by rat_head_bwd (backward()) together with the fc gradients: dh[-1] is ready
        count = float(B * self.world)
        fused = self._dnn_fused()
        for li in reversed(range(len(units))):
            lin, bn = layers[li]
            u = units[li]
            dh, out, z = ws["dh"][li], ws["h"][li], ws["z"][li]
            drop = float(s.net_dropout)
            am = self._amax_new()       # max|dz| (published by bn_act_bwd_apply): the fp16 GEMMs lift dz by a power of two
            if fused:                   # sums + apply + bias gradient in one cluster kernel
                if bn is not None:
                    call("rat_bn_act_bwd_fused", dh, out, z, ws["bn_mean"][li], ws["bn_rstd"][li], p[f"dnn.dnn.{bn}.weight"],
                         B, u, dh, g[f"dnn.dnn.{bn}.weight"], g[f"dnn.dnn.{bn}.bias"], g[f"dnn.dnn.{lin}.bias"], drop,
                         s.seed, self._rng_stream(16 + li), am, self._bnx_peers, self.rank, self.world, st)
                else:
                    call("rat_bn_act_bwd_fused", dh, out, z, None, None, None, B, u, dh, None, None,
                         g[f"dnn.dnn.{lin}.bias"], drop, s.seed, self._rng_stream(16 + li), am, None, 0, 1, st)
            elif bn is not None:
                call("rat_bn_act_bwd_sums", dh, out, z, ws["bn_mean"][li], ws["bn_rstd"][li], B, u, drop, s.seed,
                     self._rng_stream(16 + li), ws["bn_sums"][li], st)
                self._allreduce_sums(ws["bn_sums"][li])
                call("rat_bn_act_bwd_apply", dh, out, z, ws["bn_mean"][li], ws["bn_rstd"][li], p[f"dnn.dnn.{bn}.weight"],
                     ws["bn_sums"][li], count, dh, g[f"dnn.dnn.{bn}.weight"], g[f"dnn.dnn.{bn}.bias"], B, u, drop,
                     s.seed, self._rng_stream(16 + li), am, 1.0 / self.world, st)
            else:
                call("rat_bn_act_bwd_apply", dh, out, z, None, None, None, None, count, dh, None, None, B, u, drop,
                     s.seed, self._rng_stream(16 + li), am, 1.0, st)
            h_prev = ws["h"][li - 1] if li > 0 else ws["x_emb"]
            Kin = units[li - 1] if li > 0 else s.F * s.embedding_dim
            d_prev = ws["dh"][li - 1] if li > 0 else ws["dxemb"]
            call("rat_sgemm_scaled", dh, h_prev, g[f"dnn.dnn.{lin}.weight"], None, u, Kin, B, u, Kin, Kin, 1, 1, am,
                 gw, gwb, st)
            if not fused:
                call("rat_colsum", dh, B, u, u, g[f"dnn.dnn.{lin}.bias"], st)
            call("rat_sgemm_scaled", dh, p[f"dnn.dnn.{lin}.weight"], d_prev, None, B, Kin, u, u, Kin, Kin, 0, 1, am,
                 gw, gwb, st)

    def backward(self, ws, B, T):
        """after forward_ids(training=True): fill self.store.G with the data gradient of the mean BCE."""
        s, g, st = self.spec, self.store.grad_views, current_stream()
        N, D, F, L = s.F + 1, s.embedding_dim, s.F, s.L
        enc = ws["enc_out"]
        self._last_bwd = (ws, B, T)
        self._amax_reset()
        # d(loss)/d(encoder output) is non-zero only in the pooled token [b,0,0,:] (written by rat_head)
        if self._amax_on:               # max|denc| was stored by rat_head itself (max|dlogit| * max|fc.weight|)
            self._amax_of[ws["denc"].data_ptr()] = self.denc_amax
        # fc + final Linear(K -> 1) of the DNN: every gradient that hangs off dlogit, one launch
        has_dnn = len(s.dnn_hidden_units) > 0
        if has_dnn:
            _, final = dnn_layout(s)
            Kl = s.dnn_hidden_units[-1]
            call("rat_head_bwd", ws["dlogit"], B, enc, ws["enc_stride"], D, g["fc.weight"], g["fc.bias"], ws["h"][-1], Kl,
                 self.p[f"dnn.dnn.{final}.weight"], g[f"dnn.dnn.{final}.weight"], g[f"dnn.dnn.{final}.bias"],
                 ws["dh"][-1], st)
        else:
            call("rat_head_bwd", ws["dlogit"], B, enc, ws["enc_stride"], D, g["fc.weight"], g["fc.bias"], None, 0, None,
                 None, None, None, st)
        dnn_done = None
        if has_dnn:
            if self._dnn_on_side_stream():
                side2, main = self._side_stream(1), torch.cuda.current_stream()
                ready = torch.cuda.Event()
                ready.record(main)
                side2.wait_event(ready)
                with torch.cuda.stream(side2):
                    self._dnn_backward(ws, B)
                    dnn_done = torch.cuda.Event()
                    dnn_done.record(side2)
            else:
                self._dnn_backward(ws, B)
        defer = self._defer_reduce()
        if defer:           # weight-gradient record reductions leave the critical path (include/rat_b200.h)
            call("rat_set_reduce_stream", self._side_stream(0).cuda_stream)
        try:
            d = self.encode_backward(ws, B, T)
        finally:
            if defer:
                call("rat_reduce_stream_join", st)
                call("rat_set_reduce_stream", None)
        if dnn_done is not None:
            torch.cuda.current_stream().wait_event(dnn_done)
        sw = ws["scatter_ws"]
        gs = self.store
        self._net_work = None
        if self.world > 1 and gs.shard is None:     # dense-net gradients are final: their all-reduce runs under the scatter
            import torch.distributed as dist
            self._net_work = dist.all_reduce(gs.G[:gs.net_end], group=self.dist_group, async_op=True)
        if gs.shard is None:
            g_emb = gs.G[gs.emb_off:gs.emb_off + s.V * D]
            g_lr = gs.G[gs.lr_off:gs.lr_off + s.V] if s.use_wide else None
        else:       # local dense gradient over the global row space; optimizer_step reduce-scatters it to the owners
            g_emb, g_lr = gs.G_emb_full, gs.G_lr_full
        planned = ws.get("plan_event") is not None
        if planned:                 # the sorted occurrence index was built on the side stream during fwd / bwd
            torch.cuda.current_stream().wait_event(ws["plan_event"])
            ws["plan_event"] = None
        call("rat_emb_scatter_reduce", ws["ids"], ws["labels"], d, ws["dxemb"] if has_dnn else None,
             ws["dlogit"] if s.use_wide else None, self.col_off, self.col_pad, self.col_vocab, self.col_field,
             g_emb, g_lr, g["label_embedding_layer.weight"], B, T, L, F, D, s.V, float(s.emb_dropout), s.seed,
             self._rng_stream(0), 1 if planned else 0, sw, sw.numel() * 4, st)      # dropout backward fused (same mask)

    def _shard_recv(self, B, T):
        """symmetric receive buffer of the sparse row-gradient exchange (one per (B, T))."""
        key = ("recv", B, T)
        ent = self._ws.get(key)
        if ent is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            s = self.spec
            nbytes = int(query("rat_shard_recv_bytes", B, T, s.L, s.embedding_dim, self.world))
            buf = symm.empty((nbytes + 3) // 4, dtype=torch.int32, device=self.device)
            buf.zero_()
            handle = symm.rendezvous(buf, dist.group.WORLD.group_name)
            peers = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=self.device)
            counts = torch.zeros(max(self.world, 4), dtype=torch.int32, device=self.device)
            torch.cuda.synchronize()
            dist.barrier()
            ent = dict(buf=buf, handle=handle, peers=peers, counts=counts)
            self._ws[key] = ent
        return ent

    def _barrier_sums(self, t):
        """all-reduce of a small float64 vector that doubles as a cross-rank barrier of the compute streams."""
        if self._os_peers is not None:
            call("rat_oneshot_allreduce_f64", self._os_peers, self.rank, self.world, t, t.numel(), t, current_stream())
        else:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.dist_group)

    def _optimizer_step_sharded(self, ws=None, B=None, T=None):
        """net + label gradients: all-reduce; table gradients: every touched row is sent once to its owner over NVLink
        (rat_shard_send_rows / rat_shard_apply_rows: traffic proportional to the batch, not to the vocabulary);
        global-norm clip with the shard norms all-reduced; Adam on [net | label | local shard]; a final barrier orders the
        W update before the peers' next gather."""
        import torch.distributed as dist
        s, gs, st = self.spec, self.store, current_stream()
        if ws is None:
            ws, B, T = self._last_bwd                   # optimizer_step() called on its own after backward()
        Vs, D = gs.rows_per_shard, s.embedding_dim
        net_work = dist.all_reduce(gs.G[:gs.emb_off], group=self.dist_group, async_op=True)
        rv = self._shard_recv(B, T)
        sw = ws["scatter_ws"]
        call("rat_shard_send_rows", sw, sw.numel() * 4, B, T, s.L, s.F, D, s.V, Vs, self.world, self.rank, gs.G_emb_full,
             gs.G_lr_full, rv["peers"], rv["counts"], self.err_flag, st)
        nb = int(query("rat_optim_blocks"))
        if not hasattr(self, "_shard_partial"):
            self._shard_partial = torch.zeros(2 * nb, dtype=torch.float64, device=self.device)
            self._shard_extra = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._shard_extra.zero_()
        self._barrier_sums(self._shard_extra)               # every rank's records have landed in their owners' buffers
        call("rat_shard_apply_rows", rv["buf"], B, T, s.L, D, self.world, self.rank, Vs,
             gs.G[gs.emb_off:gs.emb_off + Vs * D], gs.G[gs.lr_off:gs.lr_off + Vs] if s.use_wide else None, self.err_flag, st)
        net_work.wait()
        lam_n, lam_e = float(s.net_regularizer or 0.0), float(s.embedding_regularizer or 0.0)
        # replicated part (identical on every rank: counted once) and the local shard (summed over ranks)
        call("rat_grad_sqnorm", gs.G, gs.W, gs.emb_off, gs.net_end, lam_n, lam_e, self.opt_partial, st)
        call("rat_grad_sqnorm", gs.G[gs.emb_off:], gs.W[gs.emb_off:], gs.total - gs.emb_off, 0, lam_n, lam_e,
             self._shard_partial, st)
        torch.sum(self._shard_partial.view(2, nb), dim=1, out=self._shard_extra)
        self._barrier_sums(self._shard_extra)
        call("rat_optim_prepare", self.opt_partial, nb, self._shard_extra, float(s.max_gradient_norm), self.lr, 0.9, 0.999,
             self.opt_state, 1, st)
        call("rat_adam_step", gs.W, gs.G, gs.M, gs.Vv, gs.total, gs.net_end, lam_n, lam_e, self.opt_state, 0.9, 0.999,
             1e-8, st)
        self._shard_extra.zero_()
        self._barrier_sums(self._shard_extra)               # barrier: shards updated before any peer's next gather

    def optimizer_step(self, ws=None, B=None, T=None):
        s, gs, st = self.spec, self.store, current_stream()
        if gs.shard is not None:
            return self._optimizer_step_sharded(ws, B, T)
        if self.world > 1:
            import torch.distributed as dist
            if getattr(self, "_net_work", None) is not None:
                dist.all_reduce(gs.G[gs.net_end:], group=self.dist_group)     # label / embedding / LR table gradients
                self._net_work.wait()
                self._net_work = None
            else:
                dist.all_reduce(gs.G, group=self.dist_group)
        nb = int(query("rat_optim_blocks"))
        lam_n, lam_e = float(s.net_regularizer or 0.0), float(s.embedding_regularizer or 0.0)
        call("rat_grad_sqnorm", gs.G, gs.W, gs.total, gs.net_end, lam_n, lam_e, self.opt_partial, st)
        call("rat_optim_prepare", self.opt_partial, nb, None, float(s.max_gradient_norm), self.lr, 0.9, 0.999,
             self.opt_state, 1, st)
        call("rat_adam_step", gs.W, gs.G, gs.M, gs.Vv, gs.total, gs.net_end, lam_n, lam_e, self.opt_state, 0.9, 0.999,
             1e-8, st)

    def train_step_ids(self, ws, B, T):
        """one full training step on the ids/labels/y_true already in ws. Returns ws['loss'] (device):
        [sum BCE, mean BCE of the local shard]; opt_state[5] holds the regularisation loss."""
        import rat_native as _rn
        if (self.graph_training and self.store.shard is None and _rn._profile is None
                and not torch.cuda.is_current_stream_capturing()):
            key = ("train", B, T, int(query("rat_get_precision")), float(self.spec.max_gradient_norm))
            entry = self._graphs.get(key)
            if entry is None or entry[1] is not ws:      # first call (or a new workspace): eager, warms up lazy state
                self._graphs[key] = ("warm", ws)
            else:
                if entry[0] == "warm":
                    try:
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        n0 = int(query("rat_launch_count"))
                        with _no_gc(), torch.cuda.graph(g):
                            self._train_step_impl(ws, B, T)
                        entry = (g, ws, int(query("rat_launch_count")) - n0)
                    except Exception as exc:             # capture not possible here: stay eager for this key
                        import logging
                        logging.warning("CUDA-graph capture of the training step failed (%s); running eagerly", exc)
                        torch.cuda.synchronize()
                        entry = ("eager", ws, 0)
                    self._graphs[key] = entry
                if entry[0] != "eager":
                    entry[0].replay()
                    self.replayed_launches += entry[2]
                    return ws["loss"]
        return self._train_step_impl(ws, B, T)

    def _train_step_impl(self, ws, B, T):
        st = current_stream()
        call("rat_rng_step_advance", st)
        self.rng_step += 1
        if self.spec.model != "RAT_m2":
            ws["dact"].zero_()          # RAT_m2: written in full by rat_expand_rows (encode_backward)
        if ws.get("dact_c") is not None:
            ws["dact_c"].zero_()
        self.forward_ids(ws, B, T, training=True, inv_count=1.0 / (B * self.world))
        self.backward(ws, B, T)
        self.optimizer_step(ws, B, T)
        return ws["loss"]

    def _plan_scatter(self, ws, B, T):
        """key build + radix sort of the step's occurrences (ids only) on a side stream: off the critical path."""
        s = self.spec
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)                              # ids / labels of this step are in ws
        self._side.wait_event(ready)
        sw = ws["scatter_ws"]
        with torch.cuda.stream(self._side):
            call("rat_emb_scatter_plan", ws["ids"], ws["labels"], self.col_off, self.col_pad, self.col_vocab, B, T, s.L,
                 s.F, s.embedding_dim, s.V, sw, sw.numel() * 4, self._side.cuda_stream)
            done = torch.cuda.Event()
            done.record(self._side)
        ws["plan_event"] = done

    def auc_logloss(self, y_pred: torch.Tensor, y_true: torch.Tensor):
        """(AUC, logloss) of device vectors y_pred / y_true [n] -- fuxictr/metrics.py:21-41 on the device."""
        n = int(y_pred.numel())
        y_pred = y_pred.reshape(-1).float().contiguous()
        y_true = y_true.reshape(-1).float().contiguous()
        nb = int(query("rat_auc_logloss_workspace_bytes", n))
        ws = torch.empty(nb // 4 + 4, dtype=torch.int32, device=self.device)
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        call("rat_auc_logloss", y_pred, y_true, n, out, ws, ws.numel() * 4, current_stream())
        r = out.cpu()
        return float(r[0]), float(r[1])

    def materialize_grads(self) -> Dict[str, torch.Tensor]:
        """dense gradients incl. the regulariser (what the reference's .grad holds before clipping). Tests only."""
        s, gs = self.spec, self.store
        out = torch.empty_like(gs.G)
        call("rat_materialize_grad", gs.G, gs.W, gs.total, gs.net_end, float(s.net_regularizer or 0.0),
             float(s.embedding_regularizer or 0.0), out, current_stream())
        import math as _m
        return OrderedDict((k, out[o:o + _m.prod(shp)].view(shp)) for k, (o, shp) in gs.offsets.items())

    # ------------------------------------------------------------------ row-sharded tables
    def init_sharded_tables(self, std: float, seed: int):
        """N(0, std) rows, padding rows of sequence fields zero (embedding.py:96-100); every rank draws its own shard."""
        gs, s = self.store, self.spec
        rank, world = gs.shard
        g = torch.Generator(device=self.device).manual_seed(int(seed) * 1000003 + rank)
        gs.emb_W.normal_(0.0, std, generator=g)
        if gs.lr_W is not None:
            gs.lr_W.normal_(0.0, std, generator=g)
        lo, hi = gs.row0, gs.row0 + gs.rows_per_shard
        gs.emb_W[max(0, min(s.V, hi) - lo):].zero_()                  # rows beyond V_total (padding of the partition)
        if gs.lr_W is not None:
            gs.lr_W[max(0, min(s.V, hi) - lo):].zero_()
        for f in s.features:
            if f.pad is None:
                continue
            r = gs.table_rows[f.name][0] + f.pad
            if lo <= r < hi:
                gs.emb_W[r - lo].zero_()
                if gs.lr_W is not None:
                    gs.lr_W[r - lo] = 0.0
        self.shard_barrier()

    def shard_barrier(self):
        import torch.distributed as dist
        t = torch.zeros(1, device=self.device)
        dist.all_reduce(t, group=self.dist_group)

    def gather_tables(self) -> Dict[str, torch.Tensor]:
        """full per-field tables (state_dict names) assembled from all shards -- checkpoints stay reference-compatible."""
        import torch.distributed as dist
        gs, s = self.store, self.spec
        rank, world = gs.shard
        Vs, D = gs.rows_per_shard, s.embedding_dim
        full = torch.empty(world * Vs, D, device=self.device)
        dist.all_gather_into_tensor(full, gs.emb_W.contiguous(), group=self.dist_group)
        out = OrderedDict()
        for f in s.features:
            r0, v = gs.table_rows[f.name]
            out[EMB + f.name + ".weight"] = full[r0:r0 + v].clone()
        if s.use_wide:
            fl = torch.empty(world * Vs, device=self.device)
            dist.all_gather_into_tensor(fl, gs.lr_W.contiguous(), group=self.dist_group)
            for f in s.features:
                r0, v = gs.table_rows[f.name]
                out[LRP + f.name + ".weight"] = fl[r0:r0 + v].clone().view(v, 1)
        return out

    def scatter_tables(self, sd: Dict[str, torch.Tensor]):
        """load full per-field tables (every rank passes the same state dict) into the local shard."""
        gs, s = self.store, self.spec
        lo, hi = gs.row0, gs.row0 + gs.rows_per_shard
        for f in s.features:
            r0, v = gs.table_rows[f.name]
            a, b = max(lo, r0), min(hi, r0 + v)
            if a >= b:
                continue
            for prefix, dst in ((EMB, gs.emb_W), (LRP, gs.lr_W)):
                k = prefix + f.name + ".weight"
                if dst is None or k not in sd:
                    continue
                src = sd[k].to(device=self.device, dtype=torch.float32)
                dst[a - lo:b - lo].copy_(src[a - r0:b - r0].view(dst[a - lo:b - lo].shape))
        self.shard_barrier()

    # ------------------------------------------------------------------ state
    def load_params(self, sd: Dict[str, torch.Tensor], strict=True):
        missing = []
        for k, v in self.p.items():
            if k in sd:
                v.copy_(sd[k].to(device=self.device, dtype=torch.float32).view(v.shape))
            elif not k.startswith("query_proj"):
                missing.append(k)
        for k, v in self.buffers.items():
            if k in sd:
                v.copy_(sd[k].to(self.device))
        if self.store.shard is not None:
            self.scatter_tables(sd)
        if strict and missing:
            raise KeyError(f"missing parameters: {missing[:5]}...")
