"""Helpers for the -m gpu parity tests (CUDA path called through the C ABI, oracle as checker)."""
import numpy as np
import torch

from oracle import rat_oracle as O
from rat_native.engine import EngineSpec, FeatureSpec, RatEngine


PREC = {"mode": "fp16"}


def tf32():
    return PREC["mode"] in ("tf32", "bf16", "fp16")


def bf16():
    return PREC["mode"] in ("bf16", "fp16")


def ptol(rtol, atol, rt=1e-2, at_scale=40.0):
    """(rtol, atol) for the current precision mode: fp32 values as given; tf32 (mma.sync) and fp16 (tcgen05, the
    benchmarked mode) both round their operands to a 10-bit mantissa (unit roundoff 2^-11) and accumulate in fp32:
    rtol 1e-2 and atol x40 (x80 in fp16: the gradient-domain tiles are additionally scaled to a power of two and the
    LayerNorm / GELU inputs of the backward kernels are recomputed from fp16 tiles)."""
    if bf16():
        return (max(rtol, rt), atol * at_scale * 2)
    return (max(rtol, rt), atol * at_scale) if tf32() else (rtol, atol)


def to_engine_spec(spec: O.ModelSpec, **over) -> EngineSpec:
    feats = [FeatureSpec(f.name, f.type, f.vocab_size, f.max_len, f.padding_idx) for f in spec.features]
    kw = dict(features=feats, model=spec.model, embedding_dim=spec.embedding_dim, num_heads=spec.num_heads,
              dim_head=spec.dim_head, scale_dim=spec.scale_dim, depth=spec.depth,
              dnn_hidden_units=tuple(spec.dnn_hidden_units), batch_norm=spec.batch_norm, use_wide=spec.use_wide,
              emb_dropout=spec.emb_dropout, net_dropout=spec.net_dropout,
              embedding_regularizer=spec.embedding_regularizer, net_regularizer=spec.net_regularizer,
              learning_rate=spec.learning_rate, max_gradient_norm=spec.max_gradient_norm)
    kw.update(over)
    return EngineSpec(**kw)


def make_engine(spec: O.ModelSpec, params, bufs=None, **over) -> RatEngine:
    eng = RatEngine(to_engine_spec(spec, **over), "cuda:0")
    sd = dict(params)
    if bufs:
        sd.update(bufs)
    eng.load_params(sd)
    return eng


def assert_close(name, got, want, rtol, atol):
    got = got.detach().float().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = want.detach().float().cpu().numpy() if torch.is_tensor(want) else np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    if bad.any() or not np.isfinite(got).all():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{name}: {int(bad.sum())}/{bad.size} mismatches, max abs err {err.max():.3e} "
                             f"(at {i}: got {got[i]:.6e} want {want[i]:.6e}), max |want| {np.abs(want).max():.3e}, "
                             f"finite={np.isfinite(got).all()}")


def rand_params_nontrivial(spec: O.ModelSpec, seed=0):
    """oracle params with non-degenerate embeddings / biases / LayerNorm affine (init values are ~0 / 1)."""
    p = O.init_params(spec, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in p.items():
        if k.startswith("query_proj"):
            continue
        if "embedding_layer.embedding_layer" in k:
            pad_zero = (v.abs().sum(1) == 0)
            v.mul_(3000.0 if v.shape[1] > 1 else 1000.0)
            v[pad_zero] = 0
        elif k.endswith(".bias") or "norm.weight" in k or (".dnn." in k and v.ndim == 1):
            v.add_(0.1 * torch.randn(v.shape, generator=g))
    return p


def assert_close_adam(name, got, want, rtol, atol, lr_steps, max_outlier_frac=2e-3):
    """post-Adam weights: Adam normalises every element's step to ~lr, so elements whose gradient is rounding noise
    (|g| ~ eps) can legitimately move by up to lr per step in either direction.  Allow a tiny fraction of such
    outliers, each bounded by the total possible Adam travel `lr_steps` (= n_steps * lr, plus slack)."""
    got = got.detach().float().cpu().numpy()
    want = want.detach().float().cpu().numpy()
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    frac = bad.mean()
    max_outlier_frac = max(max_outlier_frac, 2.0 / bad.size)
    assert frac <= max_outlier_frac, f"{name}: {int(bad.sum())}/{bad.size} elements off (max abs err {err.max():.3e})"
    assert err.max() <= lr_steps, f"{name}: max abs err {err.max():.3e} exceeds the Adam travel bound {lr_steps:.1e}"
