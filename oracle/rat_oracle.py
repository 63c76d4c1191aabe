"""CPU oracle for the RAT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (no nn.Module) restatement, in plain torch-CPU / numpy, of the reference's
retrieval-augmented CTR step.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product path (``www24-rat_b200/``) never does and fails loudly without its CUDA library.

Parity pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned by EXECUTING the reference itself (imported read-only from /root/reference
with module stubs) on seeded synthetic inputs: ``tests/golden/make_golden.py`` wrote the fixtures
in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file against them
(forward, loss, gradients, post-Adam weights, parameter counts 1,337,241 / 4,714,649 /
16,970,282 from the reference's own training logs).

Every function cites the reference file:line it follows (paths relative to /root/reference).
All parameters live in a flat ``dict[str, Tensor]`` keyed by the reference's state_dict names so
that a reference checkpoint is directly usable.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EMB = "embedding_layer.embedding_layer.embedding_layer."          # embedding.py:26-43,46-101
LR = "lr_layer.embedding_layer.embedding_layer.embedding_layer."   # shallow.py:23-33


# --------------------------------------------------------------------------------------
# schema
# --------------------------------------------------------------------------------------
@dataclass
class Feature:
    """One entry of FeatureMap.feature_specs (features.py:36-57)."""
    name: str
    type: str = "categorical"            # "categorical" | "sequence"
    vocab_size: int = 2
    max_len: int = 1                     # sequence only: number of consecutive id columns
    padding_idx: Optional[int] = None    # sequence: always vocab_size-1 (embedding.py:89)

    @property
    def width(self) -> int:
        return self.max_len if self.type == "sequence" else 1

    @property
    def pad(self) -> Optional[int]:
        if self.type == "sequence":
            return self.vocab_size - 1
        return self.padding_idx


@dataclass
class ModelSpec:
    """Hyper-parameters of one RAT experiment (configs/RAT_m2/*/model_config.yaml)."""
    features: List[Feature]
    model: str = "RAT_m2"                # RAT_m0 | RAT_m1 | RAT_m2 | RAT_m3
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

    @property
    def num_fields(self) -> int:
        return len(self.features)

    @property
    def input_length(self) -> int:          # features.py:46-57
        return sum(f.width for f in self.features)

    @property
    def columns(self) -> List[List[int]]:
        cols, i = [], 0
        for f in self.features:
            cols.append(list(range(i, i + f.width)))
            i += f.width
        return cols

    @property
    def total_vocab(self) -> int:
        return sum(f.vocab_size for f in self.features)


# --------------------------------------------------------------------------------------
# parameter construction (shapes + initialisers of the reference)
# --------------------------------------------------------------------------------------
def _xavier(gen, out_f, in_f):
    std = math.sqrt(2.0 / (in_f + out_f))                       # base_model.py:119-120
    return torch.randn(out_f, in_f, generator=gen) * std


def _attn_params(p, prefix, D, I, gen, project_out=True):
    p[prefix + "norm.weight"] = torch.ones(D)
    p[prefix + "norm.bias"] = torch.zeros(D)
    p[prefix + "fn.to_qkv.weight"] = _xavier(gen, 3 * I, D)       # RAT_m2.py:185
    if project_out:
        p[prefix + "fn.to_out.0.weight"] = _xavier(gen, D, I)     # RAT_m2.py:187-190
        p[prefix + "fn.to_out.0.bias"] = torch.zeros(D)


def _ff_params(p, prefix, D, M, gen):
    p[prefix + "net.0.weight"] = _xavier(gen, M, D)               # RAT_m2.py:166-172
    p[prefix + "net.0.bias"] = torch.zeros(M)
    p[prefix + "net.3.weight"] = _xavier(gen, D, M)
    p[prefix + "net.3.bias"] = torch.zeros(D)


def _transformer_params(p, prefix, spec, gen):
    """RAT_m0.py:193-202 (vit-style Transformer with final LayerNorm)."""
    D, I, M = spec.embedding_dim, spec.num_heads * spec.dim_head, spec.embedding_dim * spec.scale_dim
    project_out = not (spec.num_heads == 1 and spec.dim_head == D)
    for l in range(spec.depth):
        _attn_params(p, f"{prefix}layers.{l}.0.", D, I, gen, project_out)
        p[f"{prefix}layers.{l}.1.norm.weight"] = torch.ones(D)
        p[f"{prefix}layers.{l}.1.norm.bias"] = torch.zeros(D)
        _ff_params(p, f"{prefix}layers.{l}.1.fn.", D, M, gen)
    p[prefix + "norm.weight"] = torch.ones(D)
    p[prefix + "norm.bias"] = torch.zeros(D)


def init_params(spec: ModelSpec, seed: int = 0) -> "OrderedDict[str, Tensor]":
    """Random parameters with the reference's shapes / key names / init distributions
    (base_model.py:101-123; RAT_m2.py:57-99).  NOT bit-identical to a seeded reference model
    (module-traversal RNG order differs) -- parity tests copy weights."""
    gen = torch.Generator().manual_seed(seed)
    D, H, dh = spec.embedding_dim, spec.num_heads, spec.dim_head
    I, M, F_ = H * dh, D * spec.scale_dim, spec.num_fields
    p: "OrderedDict[str, Tensor]" = OrderedDict()
    for f in spec.features:
        w = torch.randn(f.vocab_size, D, generator=gen) * 1e-4      # base_model.py:45,103-116
        if f.pad is not None:
            w[f.pad] = 0.0
        p[EMB + f.name + ".weight"] = w
    p["label_embedding_layer.weight"] = torch.randn(3, D, generator=gen)   # RAT_m2.py:64 (N(0,1) kept)
    p["query_proj.weight"] = _xavier(gen, F_ * D, F_ * D)                   # RAT_m2.py:66 (dead)
    p["query_proj.bias"] = torch.zeros(F_ * D)
    project_out = not (H == 1 and dh == D)
    if spec.model == "RAT_m2":
        for l in range(spec.depth):                                         # RAT_m2.py:204-217
            _attn_params(p, f"encoder.encoder.{l}.cross_attention.", D, I, gen, project_out)
            _attn_params(p, f"encoder.encoder.{l}.intra_attention.", D, I, gen, project_out)
            _ff_params(p, f"encoder.encoder.{l}.mlp.", D, M, gen)
    elif spec.model == "RAT_m0":
        _transformer_params(p, "encoder.", spec, gen)                        # RAT_m0.py:71
    elif spec.model == "RAT_m1":
        _transformer_params(p, "intra_transformer.", spec, gen)              # RAT_m1.py:70-71
        _transformer_params(p, "cross_transformer.", spec, gen)
    elif spec.model == "RAT_m3":
        for l in range(spec.depth):                                         # RAT_m3.py:200-221
            pre = f"encoder.encoder.{l}."
            for nm in ("W_q", "W_k_s", "W_v_s", "W_k_t", "W_v_t"):
                p[pre + nm + ".weight"] = _xavier(gen, I, D)
            for att in ("intra_attention.", "cross_attention."):
                p[pre + att + "norm.weight"] = torch.ones(D)
                p[pre + att + "norm.bias"] = torch.zeros(D)
                p[pre + att + "fn.to_out.0.weight"] = _xavier(gen, D, I)
                p[pre + att + "fn.to_out.0.bias"] = torch.zeros(D)
            _ff_params(p, pre + "mlp.", D, M, gen)
    else:
        raise ValueError(spec.model)
    if spec.use_wide:
        for f in spec.features:
            w = torch.randn(f.vocab_size, 1, generator=gen) * 1e-4
            if f.pad is not None:
                w[f.pad] = 0.0
            p[LR + f.name + ".weight"] = w
    units = [F_ * D] + list(spec.dnn_hidden_units)
    i = 0
    for a, b in zip(units[:-1], units[1:]):                                  # deep.py:126-135
        p[f"dnn.dnn.{i}.weight"] = _xavier(gen, b, a)
        p[f"dnn.dnn.{i}.bias"] = torch.zeros(b)
        i += 1
        if spec.batch_norm:
            p[f"dnn.dnn.{i}.weight"] = torch.ones(b)
            p[f"dnn.dnn.{i}.bias"] = torch.zeros(b)
            i += 1
        i += 1                                                               # activation slot
        if spec.net_dropout > 0:
            i += 1                                                           # dropout slot
    p[f"dnn.dnn.{i}.weight"] = _xavier(gen, 1, units[-1])
    p[f"dnn.dnn.{i}.bias"] = torch.zeros(1)
    p["fc.weight"] = _xavier(gen, 1, D)                                      # RAT_m2.py:98
    p["fc.bias"] = torch.zeros(1)
    return p


def init_buffers(spec: ModelSpec) -> "OrderedDict[str, Tensor]":
    """BatchNorm1d running statistics (deep.py:128-129)."""
    b: "OrderedDict[str, Tensor]" = OrderedDict()
    if not spec.batch_norm:
        return b
    for i, width in dnn_bn_slots(spec):
        b[f"dnn.dnn.{i}.running_mean"] = torch.zeros(width)
        b[f"dnn.dnn.{i}.running_var"] = torch.ones(width)
        b[f"dnn.dnn.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


def dnn_layout(spec: ModelSpec) -> Tuple[List[Tuple[int, Optional[int]]], int]:
    """Indices inside ``dnn.dnn`` (nn.Sequential): [(linear_idx, bn_idx|None)...], final_linear_idx."""
    out, i = [], 0
    for _ in spec.dnn_hidden_units:
        lin, bn = i, None
        i += 1
        if spec.batch_norm:
            bn = i
            i += 1
        i += 1
        if spec.net_dropout > 0:
            i += 1
        out.append((lin, bn))
    return out, i


def dnn_bn_slots(spec: ModelSpec) -> List[Tuple[int, int]]:
    layers, _ = dnn_layout(spec)
    return [(bn, w) for (lin, bn), w in zip(layers, spec.dnn_hidden_units) if bn is not None]


def count_parameters(params: Dict[str, Tensor]) -> int:
    """base_model.py:293-301 (all registered parameters incl. the dead query_proj)."""
    return sum(int(v.numel()) for v in params.values())


def is_embedding_named(name: str) -> bool:
    """base_model.py:86 -- substring match, so label_embedding_layer and LR tables count."""
    return "embedding_layer" in name


# --------------------------------------------------------------------------------------
# a1: retrieval-set assembly  (data_generator.py:66-78)
# --------------------------------------------------------------------------------------
def assemble_batch(darray: np.ndarray, pool: np.ndarray, retr_indices: np.ndarray,
                   rows: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """X [B,1+K,L] , y [B,1+K]; numpy fancy indexing => index -1 wraps to the last pool row."""
    tgt = darray[rows][:, None, :]                                # data_generator.py:67,70
    nbr = pool[retr_indices[rows]]                                # :69  (negative index wraps)
    full = np.concatenate([tgt, nbr], axis=1)                     # :71
    return full[..., :-1], full[..., -1]                          # :72-73


# --------------------------------------------------------------------------------------
# a3/a4: embedding gather + label token  (embedding.py:158-178, RAT_m2.py:115-126)
# --------------------------------------------------------------------------------------
def embed_rows(params: Dict[str, Tensor], spec: ModelSpec, ids: Tensor, prefix: str = EMB) -> Tensor:
    """ids [..., L] int64 -> [..., F, D'] ; sequence fields are sum-pooled (sequence.py:36-38)."""
    outs = []
    for f, cols in zip(spec.features, spec.columns):
        w = params[prefix + f.name + ".weight"]
        if f.type == "sequence":
            e = F.embedding(ids[..., cols], w, padding_idx=f.pad).sum(dim=-2)
        else:
            e = F.embedding(ids[..., cols[0]], w, padding_idx=f.pad)
        outs.append(e)
    return torch.stack(outs, dim=-2)                              # embedding.py:156


def feature_block(params: Dict[str, Tensor], spec: ModelSpec, ids: Tensor, labels: Tensor) -> Tensor:
    """[B,T,F+1,D]: token 0 of every row is the label embedding; the target row gets id 2."""
    lab = labels.long().clone()
    lab[:, 0] = 2                                                 # RAT_m2.py:115-116,123
    lab_emb = F.embedding(lab, params["label_embedding_layer.weight"])[:, :, None, :]
    return torch.cat([lab_emb, embed_rows(params, spec, ids)], dim=2)   # RAT_m2.py:124-126


# --------------------------------------------------------------------------------------
# a6/a7: attention + encoders
# --------------------------------------------------------------------------------------
def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def mha(z: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, heads: int, scale: float,
        wo: Optional[Tensor], bo: Optional[Tensor]) -> Tensor:
    """RAT_m2.py:192-202.  z [S,len,D] -> [S,len,D]."""
    S, n, _ = z.shape
    q = (z @ wq.t()).view(S, n, heads, -1).transpose(1, 2)
    k = (z @ wk.t()).view(S, n, heads, -1).transpose(1, 2)
    v = (z @ wv.t()).view(S, n, heads, -1).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(S, n, -1)
    if wo is None:
        return o
    return o @ wo.t() + bo


def _prenorm_attn(p, pre, z, spec):
    I = spec.num_heads * spec.dim_head
    w = p[pre + "fn.to_qkv.weight"]
    wo = p.get(pre + "fn.to_out.0.weight")
    bo = p.get(pre + "fn.to_out.0.bias")
    zn = _ln(z, p[pre + "norm.weight"], p[pre + "norm.bias"])
    return mha(zn, w[:I], w[I:2 * I], w[2 * I:], spec.num_heads, spec.dim_head ** -0.5, wo, bo)


def _ff(p, pre, z):
    h = F.gelu(z @ p[pre + "net.0.weight"].t() + p[pre + "net.0.bias"])      # exact erf GELU
    return h @ p[pre + "net.3.weight"].t() + p[pre + "net.3.bias"]


def encoder_m2(p: Dict[str, Tensor], spec: ModelSpec, x: Tensor) -> Tensor:
    """RAT_m2.py:219-236,252-259.  x [B,T,N,D]."""
    B, T, N, D = x.shape
    for l in range(spec.depth):
        pre = f"encoder.encoder.{l}."
        u = x.reshape(B * T, N, D)
        u = _prenorm_attn(p, pre + "intra_attention.", u, spec) + u
        v = u.reshape(B, T, N, D).transpose(1, 2).reshape(B * N, T, D)
        v = _prenorm_attn(p, pre + "cross_attention.", v, spec) + v
        v = _ff(p, pre + "mlp.", v) + v                                          # no pre-norm
        x = v.reshape(B, N, T, D).transpose(1, 2)
    return x


def transformer(p: Dict[str, Tensor], prefix: str, spec: ModelSpec, z: Tensor) -> Tensor:
    """RAT_m0.py:193-208."""
    for l in range(spec.depth):
        z = _prenorm_attn(p, f"{prefix}layers.{l}.0.", z, spec) + z
        pre = f"{prefix}layers.{l}.1."
        z = _ff(p, pre + "fn.", _ln(z, p[pre + "norm.weight"], p[pre + "norm.bias"])) + z
    return _ln(z, p[prefix + "norm.weight"], p[prefix + "norm.bias"])


def encoder_m3(p: Dict[str, Tensor], spec: ModelSpec, x: Tensor) -> Tensor:
    """RAT_m3.py:164-242: parallel intra || cross attention sharing W_q; heads halved,
    scale still dim_head**-0.5; x = mlp(mean(out_s,out_t)) + x."""
    B, T, N, D = x.shape
    h = int(spec.num_heads / 2)
    scale = spec.dim_head ** -0.5
    for l in range(spec.depth):
        pre = f"encoder.encoder.{l}."
        u = x.reshape(B * T, N, D)
        un = _ln(u, p[pre + "intra_attention.norm.weight"], p[pre + "intra_attention.norm.bias"])
        out_s = mha(un, p[pre + "W_q.weight"], p[pre + "W_k_s.weight"], p[pre + "W_v_s.weight"], h, scale,
                    p[pre + "intra_attention.fn.to_out.0.weight"], p[pre + "intra_attention.fn.to_out.0.bias"])
        out_s = out_s.reshape(B, T, N, D)
        v = x.transpose(1, 2).reshape(B * N, T, D)
        vn = _ln(v, p[pre + "cross_attention.norm.weight"], p[pre + "cross_attention.norm.bias"])
        out_t = mha(vn, p[pre + "W_q.weight"], p[pre + "W_k_t.weight"], p[pre + "W_v_t.weight"], h, scale,
                    p[pre + "cross_attention.fn.to_out.0.weight"], p[pre + "cross_attention.fn.to_out.0.bias"])
        out_t = out_t.reshape(B, N, T, D).transpose(1, 2)
        out = 0.5 * (out_s + out_t)
        x = _ff(p, pre + "mlp.", out) + x
    return x


def encode(p: Dict[str, Tensor], spec: ModelSpec, x: Tensor) -> Tensor:
    """feature block [B,T,N,D] -> pooled token [B,D] (RAT_m{0,1,2,3}.py forward)."""
    B, T, N, D = x.shape
    if spec.model == "RAT_m2":
        return encoder_m2(p, spec, x)[:, 0, 0]                                   # RAT_m2.py:137-140
    if spec.model == "RAT_m3":
        return encoder_m3(p, spec, x)[:, 0, 0]
    if spec.model == "RAT_m0":
        z = transformer(p, "encoder.", spec, x.reshape(B, T * N, D))             # RAT_m0.py:123-127
        return z.reshape(B, T, N, D)[:, 0, 0]
    if spec.model == "RAT_m1":
        z = transformer(p, "intra_transformer.", spec, x.reshape(B * T, N, D))   # RAT_m1.py:123-129
        z = z[:, 0].reshape(B, T, D)
        return transformer(p, "cross_transformer.", spec, z)[:, 0]
    raise ValueError(spec.model)


# --------------------------------------------------------------------------------------
# a8: head  (deep.py:108-141, shallow.py:36-45, RAT_m2.py:144-150)
# --------------------------------------------------------------------------------------
def dnn_forward(p: Dict[str, Tensor], bufs: Dict[str, Tensor], spec: ModelSpec, h: Tensor,
                training: bool, drop_masks: Optional[List[Tensor]] = None) -> Tensor:
    layers, final = dnn_layout(spec)
    for li, (lin, bn) in enumerate(layers):
        h = h @ p[f"dnn.dnn.{lin}.weight"].t() + p[f"dnn.dnn.{lin}.bias"]
        if bn is not None:
            rm, rv = bufs[f"dnn.dnn.{bn}.running_mean"], bufs[f"dnn.dnn.{bn}.running_var"]
            h = F.batch_norm(h, rm, rv, p[f"dnn.dnn.{bn}.weight"], p[f"dnn.dnn.{bn}.bias"],
                             training=training, momentum=0.1, eps=1e-5)
            if training:
                bufs[f"dnn.dnn.{bn}.num_batches_tracked"] += 1
        h = torch.relu(h)
        if training and spec.net_dropout > 0 and drop_masks is not None:
            h = h * drop_masks[li] / (1.0 - spec.net_dropout)
    return h @ p[f"dnn.dnn.{final}.weight"].t() + p[f"dnn.dnn.{final}.bias"]


def forward(p: Dict[str, Tensor], bufs: Dict[str, Tensor], spec: ModelSpec, X: Tensor, y: Tensor,
            training: bool = False, emb_mask: Optional[Tensor] = None,
            dnn_masks: Optional[List[Tensor]] = None, return_parts: bool = False):
    """X [B,T,L] (float64 wire format or integer), y [B,T].  Returns y_pred [B,1] (probabilities).

    ``emb_mask`` [B,T,N,D] / ``dnn_masks`` are optional 0/1 keep-masks so that train-mode dropout can be
    replayed deterministically (the reference's philox stream is not reproducible)."""
    ids = X.long()                                                              # embedding.py:166
    block = feature_block(p, spec, ids, y)                                      # [B,T,N,D]
    x_emb = block[:, 0, 1:, :]                                                  # RAT_m2.py:120 (no dropout)
    x = block
    if training and spec.emb_dropout > 0 and emb_mask is not None:
        x = x * emb_mask / (1.0 - spec.emb_dropout)                             # RAT_m2.py:135
    pooled = encode(p, spec, x)
    logit = pooled @ p["fc.weight"].t() + p["fc.bias"]                          # RAT_m2.py:144
    if len(spec.dnn_hidden_units) > 0:
        logit = logit + dnn_forward(p, bufs, spec, x_emb.flatten(1), training, dnn_masks)
    if spec.use_wide:
        lr = embed_rows(p, spec, ids[:, 0:1, :], prefix=LR).sum(dim=-2).mean(dim=1)   # shallow.py:37-40
        logit = logit + lr
    y_pred = torch.sigmoid(logit)
    if return_parts:
        return y_pred, dict(block=block, pooled=pooled, logit=logit)
    return y_pred


# --------------------------------------------------------------------------------------
# a9/a10: loss + regularisation ; a11-a13: backward, clip, Adam
# --------------------------------------------------------------------------------------
def bce_mean(y_pred: Tensor, y_true: Tensor) -> Tensor:
    return F.binary_cross_entropy(y_pred, y_true, reduction="mean")              # base_model.py:74-77


def regularization(p: Dict[str, Tensor], spec: ModelSpec) -> Tensor:
    """base_model.py:79-94 with get_regularizer(float) -> [(2, lambda)] (torch_utils.py:65-68)."""
    reg = torch.zeros(())
    for name, w in p.items():
        lam = spec.embedding_regularizer if is_embedding_named(name) else spec.net_regularizer
        if lam:
            reg = reg + (lam / 2.0) * torch.norm(w, 2) ** 2
    return reg


@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, Tensor] = field(default_factory=dict)
    v: Dict[str, Tensor] = field(default_factory=dict)


def total_loss_and_grads(p, bufs, spec, X, y, emb_mask=None, dnn_masks=None):
    """One training forward+backward (base_model.py:221-223). Returns (loss, bce, grads)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    y_pred = forward(leaves, bufs, spec, X, y, training=True, emb_mask=emb_mask, dnn_masks=dnn_masks)
    bce = bce_mean(y_pred, y[:, 0:1].float())
    loss = bce + regularization(leaves, spec)
    loss.backward()
    grads = OrderedDict((k, v.grad) for k, v in leaves.items())                 # None for dead params
    return loss.detach(), bce.detach(), grads


def clip_grad_norm(grads: Dict[str, Optional[Tensor]], max_norm: float) -> Tuple[Tensor, float]:
    """nn.utils.clip_grad_norm_ (base_model.py:224): global L2 over params that have a grad."""
    sq = sum((g.double() ** 2).sum() for g in grads.values() if g is not None)
    total = torch.sqrt(sq).float()
    coef = float(min(1.0, max_norm / (float(total) + 1e-6)))
    return total, coef


def adam_update(p: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], st: AdamState, lr: float,
                coef: float = 1.0, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults (torch_utils.py:41-49): dense, every parameter that has a grad."""
    st.step += 1
    bc1, bc2 = 1.0 - b1 ** st.step, 1.0 - b2 ** st.step
    for k, w in p.items():
        g = grads.get(k)
        if g is None:
            continue
        g = g * coef
        m = st.m.setdefault(k, torch.zeros_like(w))
        v = st.v.setdefault(k, torch.zeros_like(w))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(m, denom, value=-lr / bc1)


def train_step(p, bufs, spec, st: AdamState, X, y, lr=None, emb_mask=None, dnn_masks=None):
    """zero_grad -> loss -> backward -> clip -> Adam (base_model.py:221-225). Mutates p/bufs/st."""
    loss, bce, grads = total_loss_and_grads(p, bufs, spec, X, y, emb_mask, dnn_masks)
    norm, coef = clip_grad_norm(grads, spec.max_gradient_norm)
    adam_update(p, grads, st, spec.learning_rate if lr is None else lr, coef)
    return dict(loss=float(loss), bce=float(bce), grad_norm=float(norm), clip_coef=coef, grads=grads)


# --------------------------------------------------------------------------------------
# a15: metrics  (metrics.py:21-29: sklearn roc_auc_score / log_loss(eps=1e-7))
# --------------------------------------------------------------------------------------
def logloss(y_true: np.ndarray, y_pred: np.ndarray, eps: float = 1e-7) -> float:
    p = np.clip(np.asarray(y_pred, np.float64), eps, 1 - eps)
    t = np.asarray(y_true, np.float64)
    return float(-(t * np.log(p) + (1 - t) * np.log(1 - p)).mean())


def auc(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """Rank-statistic AUC with average ranks for ties == sklearn.roc_auc_score for binary labels."""
    t = np.asarray(y_true, np.float64) > 0.5
    s = np.asarray(y_pred, np.float64)
    order = np.argsort(s, kind="mergesort")
    ss = s[order]
    ranks = np.empty(len(s), np.float64)
    i = 0
    n = len(s)
    # average ranks over tie groups
    boundaries = np.flatnonzero(np.concatenate([[True], ss[1:] != ss[:-1], [True]]))
    for a, b in zip(boundaries[:-1], boundaries[1:]):
        ranks[order[a:b]] = 0.5 * (a + b - 1) + 1.0
    npos = int(t.sum())
    nneg = n - npos
    if npos == 0 or nneg == 0:
        return float("nan")
    return float((ranks[t].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))


# --------------------------------------------------------------------------------------
# synthetic shapes of the three shipped configs (vocab split documented in DESIGN.md)
# --------------------------------------------------------------------------------------
def _split_vocab(total: int, weights: Sequence[float], minimum: int = 3) -> List[int]:
    raw = [max(minimum, int(total * w / sum(weights))) for w in weights]
    raw[int(np.argmax(raw))] += total - sum(raw)
    return raw


def shape_spec(name: str, model: str = "RAT_m2", vocab_scale: float = 1.0, **over) -> ModelSpec:
    """ModelSpec for 'ml' | 'kkbox' | 'tmall' with V_total = 90,239 / 92,247 / 1,529,680 rows
    (the totals that reproduce the logged parameter counts; per-field split is synthetic)."""
    if name == "ml":
        V = _split_vocab(int(90239 * vocab_scale), [17, 24, 49])
        feats = [Feature(n, "categorical", v) for n, v in zip(["user_id", "item_id", "tag_id"], V)]
        kw = dict(embedding_dim=10, num_heads=2, scale_dim=4, dnn_hidden_units=(400, 400, 400), batch_norm=False,
                  use_wide=True, emb_dropout=0.0, net_dropout=0.0, embedding_regularizer=0.03)
    elif name == "kkbox":
        names = ["msno", "song_id", "source_system_tab", "source_screen_name", "source_type", "city", "gender",
                 "registered_via", "language", "genre_ids", "artist_name", "isrc", "bd"]
        wts = [30000, 48000, 10, 22, 14, 23, 4, 7, 12, 170, 11800, 110, 75]
        V = _split_vocab(int(92247 * vocab_scale), wts)
        feats = []
        for n, v in zip(names, V):
            if n in ("genre_ids", "artist_name"):
                feats.append(Feature(n, "sequence", v, max_len=3))
            else:
                feats.append(Feature(n, "categorical", v))
        kw = dict(embedding_dim=40, num_heads=8, scale_dim=2, dnn_hidden_units=(400, 400, 400), batch_norm=True,
                  use_wide=True, emb_dropout=0.1, net_dropout=0.0, embedding_regularizer=0.0005)
    elif name == "tmall":
        names = ["user_id", "item_id", "cat_id", "seller_id", "brand_id", "action_type", "age_range", "gender"]
        wts = [400000, 1100000, 1600, 5000, 8400, 5, 10, 4]
        V = _split_vocab(int(1529680 * vocab_scale), wts)
        feats = [Feature(n, "categorical", v) for n, v in zip(names, V)]
        kw = dict(embedding_dim=10, num_heads=32, scale_dim=2, dnn_hidden_units=(200, 80), batch_norm=True,
                  use_wide=True, emb_dropout=0.1, net_dropout=0.08, embedding_regularizer=0.07)
    else:
        raise ValueError(name)
    kw.setdefault("dim_head", 10)
    kw.setdefault("depth", 4)
    kw.update(over)
    return ModelSpec(features=feats, model=model, **kw)


def synthetic_pool(spec: ModelSpec, n_rows: int, seed: int, pos_ratio: float = 0.5,
                   zipf_a: float = 1.05) -> np.ndarray:
    """[n_rows, L+1] float64 'h5' array: id columns then the label (data_utils.py:46-54 layout).
    ids ~ clipped Zipf in [1, V-1) ; sequence fields get 1..max_len valid ids then the pad id."""
    rng = np.random.default_rng(seed)
    cols = []
    for f in spec.features:
        hi = max(2, f.vocab_size - (1 if f.pad is not None else 0))
        z = rng.zipf(zipf_a, size=(n_rows, f.width)).astype(np.int64)
        ids = 1 + (z - 1) % (hi - 1) if hi > 2 else np.ones_like(z)
        ids = np.minimum(ids, hi - 1)
        if f.type == "sequence":
            nvalid = rng.integers(1, f.width + 1, size=n_rows)
            mask = np.arange(f.width)[None, :] >= nvalid[:, None]
            ids[mask] = f.pad
        cols.append(ids)
    lab = (rng.random(n_rows) < pos_ratio).astype(np.int64)[:, None]
    return np.concatenate(cols + [lab], axis=1).astype(np.float64)


def synthetic_neighbours(n_query: int, n_pool: int, K: int, seed: int, missing: float = 0.02) -> np.ndarray:
    """[Q,K] int64 neighbour indices, tail-padded with -1 like sort_results (data_utils.py:787-794)."""
    rng = np.random.default_rng(seed + 7)
    idx = rng.integers(0, n_pool, size=(n_query, K), dtype=np.int64)
    nmiss = (rng.random(n_query) < missing) * rng.integers(1, K + 1, size=n_query)
    tail = np.arange(K)[None, :] >= (K - nmiss)[:, None]
    idx[tail] = -1
    return idx
