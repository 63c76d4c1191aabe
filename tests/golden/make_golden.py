"""Generate the golden fixtures in tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, read-only).  The reference is imported
with the module stubs documented in SURVEY.md Appendix C (h5py/dgl/matplotlib/thop/tensorflow are
absent here); nothing from it is copied into this repository -- only its numeric outputs on
seeded synthetic inputs are stored.  Re-run:  python tests/golden/make_golden.py

Each fixture holds: the experiment spec (json), the reference state_dict before training
(minus the dead query_proj), the wire-format batch (X f64 [B,1+K,L], y f64 [B,1+K]) produced by the
reference's own Dataset.__getitem__ + default collate (so the -1 wraparound is pinned too), the
eval-mode forward outputs, and two reference training steps (loss, grad-norm, all gradients of
step 1, all parameters + BN buffers after step 2).
"""
import importlib
import importlib.machinery as mm
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)

    def _stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = mm.ModuleSpec(name, None)
        m.__path__ = []
        m.__dict__.update(attrs)
        sys.modules[name] = m

    for n in ["h5py", "dgl", "dgl.function", "dgl.nn", "dgl.nn.functional", "matplotlib", "matplotlib.pyplot",
              "thop", "tensorflow", "tensorflow.keras", "tensorflow.keras.utils"]:
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    sys.modules["dgl.nn.functional"].__dict__.setdefault("edge_softmax", None)
    sys.modules["matplotlib"].__dict__.setdefault("use", lambda *a, **k: None)
    sys.modules["thop"].__dict__.setdefault("profile", None)
    sys.modules["tensorflow.keras.utils"].__dict__.setdefault("pad_sequences", None)
    import fuxictr  # noqa
    from fuxictr.pytorch import models
    from fuxictr.pytorch.data_generator import Dataset
    from fuxictr.features import FeatureMap
    for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
        del sys.modules[k]
    np.Inf = np.inf
    return models, Dataset, FeatureMap


def build_feature_map(FeatureMap, dataset_id, feats):
    fm = FeatureMap(dataset_id, "/tmp/rat_golden")
    for name, typ, vocab, max_len in feats:
        spec = {"source": "", "type": typ, "vocab_size": vocab}
        if typ == "sequence":
            spec.update({"padding_idx": vocab - 1, "max_len": max_len, "encoder": "MaskedSumPooling"})
        fm.feature_specs[name] = spec
    fm.num_fields = len(feats)
    fm.set_feature_index()
    return fm


def synth_array(feats, n_rows, rng):
    cols = []
    for name, typ, vocab, max_len in feats:
        if typ == "sequence":
            ids = rng.integers(0, vocab - 1, size=(n_rows, max_len))
            nvalid = rng.integers(1, max_len + 1, size=n_rows)
            ids[np.arange(max_len)[None, :] >= nvalid[:, None]] = vocab - 1
        else:
            ids = rng.integers(0, vocab, size=(n_rows, 1))
        cols.append(ids)
    lab = rng.integers(0, 2, size=(n_rows, 1))
    return np.concatenate(cols + [lab], axis=1).astype(np.float64)


CASES = {
    # name: (model, feats, hyper-params, B, K)
    "ml_small": ("RAT_m2", [("user_id", "categorical", 50, 1), ("item_id", "categorical", 40, 1),
                            ("tag_id", "categorical", 30, 1)],
                 dict(embedding_dim=10, num_heads=2, dim_head=10, scale_dim=4, depth=4, dnn_hidden_units=[32, 16],
                      batch_norm=False, use_wide=True, embedding_regularizer=0.03, net_regularizer=0), 16, 5),
    "kkbox_small": ("RAT_m2",
                    [("msno", "categorical", 60, 1), ("song_id", "categorical", 70, 1),
                     ("source_system_tab", "categorical", 10, 1), ("source_screen_name", "categorical", 22, 1),
                     ("source_type", "categorical", 14, 1), ("city", "categorical", 23, 1),
                     ("gender", "categorical", 4, 1), ("registered_via", "categorical", 7, 1),
                     ("language", "categorical", 12, 1), ("genre_ids", "sequence", 25, 3),
                     ("artist_name", "sequence", 45, 3), ("isrc", "categorical", 30, 1),
                     ("bd", "categorical", 16, 1)],
                    dict(embedding_dim=40, num_heads=8, dim_head=10, scale_dim=2, depth=2,
                         dnn_hidden_units=[48, 32, 16], batch_norm=True, use_wide=True,
                         embedding_regularizer=0.0005), 8, 5),
    "tmall_small": ("RAT_m2",
                    [("user_id", "categorical", 90, 1), ("item_id", "categorical", 120, 1),
                     ("cat_id", "categorical", 30, 1), ("seller_id", "categorical", 40, 1),
                     ("brand_id", "categorical", 50, 1), ("action_type", "categorical", 5, 1),
                     ("age_range", "categorical", 10, 1), ("gender", "categorical", 4, 1)],
                    dict(embedding_dim=10, num_heads=32, dim_head=10, scale_dim=2, depth=2,
                         dnn_hidden_units=[40, 16], batch_norm=True, use_wide=True,
                         embedding_regularizer=0.07), 12, 3),
}
_VAR_FEATS = [("u", "categorical", 40, 1), ("i", "categorical", 50, 1), ("c", "categorical", 9, 1),
              ("g", "sequence", 12, 3), ("s", "categorical", 6, 1)]
for _m in ("RAT_m0", "RAT_m1", "RAT_m3"):
    CASES[_m.lower() + "_small"] = (_m, _VAR_FEATS,
                                    dict(embedding_dim=20, num_heads=4, dim_head=10, scale_dim=2, depth=2,
                                         dnn_hidden_units=[24, 12], batch_norm=True, use_wide=(_m != "RAT_m1"),
                                         embedding_regularizer=0.001), 6, 4)


# BASELINE configs[3]: the variants on the kkbox SCHEMA (13 fields incl. two sum-pooled sequence fields, D=40, 8 heads of
# width 10, K=5): RAT_m0 attends over one flat sequence of 6*14 = 84 tokens, RAT_m1 over 14 then 6, RAT_m3 uses 4 heads of
# width 20.  Vocabulary, depth and DNN widths are reduced to keep the fixtures small.
for _m in ("RAT_m0", "RAT_m1", "RAT_m3"):
    CASES[_m.lower() + "_kkbox"] = (_m, CASES["kkbox_small"][1],
                                    dict(embedding_dim=40, num_heads=8, dim_head=10, scale_dim=2, depth=2,
                                         dnn_hidden_units=[48, 32, 16], batch_norm=True, use_wide=(_m != "RAT_m1"),
                                         embedding_regularizer=0.0005), 8, 5)


def run_case(name, models, Dataset, FeatureMap):
    model_name, feats, hp, B, K = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    torch.manual_seed(sum(map(ord, name)))
    fm = build_feature_map(FeatureMap, name, feats)
    n_pool, n_q = 64, B
    pool = synth_array(feats, n_pool, rng)
    darray = synth_array(feats, n_q, rng)
    idx = rng.integers(0, n_pool, size=(n_q, K)).astype(np.int64)
    idx[1, K - 1] = -1                     # BM25 found < K matches -> -1 (wraps to the last pool row)
    idx[2, max(0, K - 2):] = -1
    vals = rng.random((n_q, K))
    lens = (idx >= 0).sum(1).astype(np.int64)
    ds = Dataset(darray, feature_map=fm, retr_pool_darray=pool, retr_indices=idx, retr_values=vals, retr_lens=lens)
    loader = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False)
    batch = next(iter(loader))
    X, y, rv, rl = batch
    kwargs = dict(model_id=name, gpu=-1, task="binary_classification", learning_rate=1e-3,
                  dnn_activations="relu", net_dropout=0, emb_dropout=0.0, dropout=0.0,
                  optimizer="adam", loss="binary_crossentropy", model_root="/tmp/rat_golden/", metrics=["AUC", "logloss"],
                  verbose=0, retrieval_augmented=True, retrieval_configs={"topK": K, "label_wise": False})
    kwargs.update(hp)
    model = getattr(models, model_name)(fm, **kwargs)
    # give the zero-initialised biases / unit LayerNorm+BN affine parameters non-trivial values so that
    # their forward use and gradients are actually pinned
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.startswith("query_proj"):
                continue
            if n_.endswith(".bias") or "norm.weight" in n_ or (".dnn." in n_ and p_.ndim == 1):
                p_.add_(0.1 * torch.randn_like(p_))
        # embedding rows are N(0,1e-4): scale up so attention is not degenerate; keep padding rows zero
        for n_, p_ in model.named_parameters():
            if "embedding_layer.embedding_layer" in n_:
                p_.mul_(3000.0 if p_.shape[1] > 1 else 1000.0)
    out = {}
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd0.items():
        if k.startswith("query_proj"):
            continue
        out["sd0/" + k] = v.numpy()
    out["X"], out["y"] = X.numpy(), y.numpy()
    out["retr_indices"], out["darray"], out["pool"] = idx, darray, pool
    # eval forward
    model.eval()
    with torch.no_grad():
        rd = model.forward(batch)
    out["eval/y_pred"] = rd["y_pred"].numpy()
    out["eval/y_true"] = rd["y_true"].numpy()
    # two training steps, exactly base_model.py:221-225
    model.train()
    for step in (1, 2):
        model.optimizer.zero_grad()
        loss = model.get_total_loss(batch)
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        if step == 1:
            for n_, p_ in model.named_parameters():
                if n_.startswith("query_proj"):
                    assert p_.grad is None
                    continue
                # clip_grad_norm_ scaled the grads in place; store the UNclipped gradient
                coef = min(1.0, 10.0 / (float(norm) + 1e-6))
                out["grad1/" + n_] = (p_.grad / coef).numpy()
        model.optimizer.step()
        out[f"train/loss{step}"] = np.float64(loss.item())
        out[f"train/norm{step}"] = np.float64(float(norm))
    for k, v in model.state_dict().items():
        if k.startswith("query_proj"):
            continue
        out["sd2/" + k] = v.detach().numpy()
    out["param_count"] = np.int64(sum(p.numel() for p in model.parameters() if p.requires_grad))
    meta = dict(model=model_name, feats=feats, hp=hp, B=B, K=K)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["train/loss1"], out["train/loss2"], "norm", out["train/norm1"],
          "params", int(out["param_count"]), "y_pred[:3]", out["eval/y_pred"][:3, 0])


def param_counts(models, FeatureMap):
    """Known-answer: the three 'Total number of parameters' lines of exps/RAT_m2/*/*.log."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.rat_oracle import shape_spec
    res = {}
    for shape in ("ml", "kkbox", "tmall"):
        spec = shape_spec(shape)
        feats = [(f.name, f.type, f.vocab_size, f.max_len) for f in spec.features]
        fm = build_feature_map(FeatureMap, shape, feats)
        model = models.RAT_m2(fm, model_id=shape, gpu=-1, learning_rate=1e-3, embedding_dim=spec.embedding_dim,
                              dnn_hidden_units=list(spec.dnn_hidden_units), dnn_activations="relu",
                              num_heads=spec.num_heads, net_dropout=spec.net_dropout, batch_norm=spec.batch_norm,
                              use_wide=True, embedding_regularizer=spec.embedding_regularizer, depth=4, dim_head=10,
                              emb_dropout=spec.emb_dropout, scale_dim=spec.scale_dim, optimizer="adam",
                              loss="binary_crossentropy", model_root="/tmp/rat_golden/", metrics=["AUC"], verbose=0,
                              retrieval_augmented=True, retrieval_configs={"topK": 5, "label_wise": False})
        res[shape] = int(sum(p.numel() for p in model.parameters() if p.requires_grad))
        print(shape, res[shape])
    with open(os.path.join(HERE, "param_counts.json"), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    models, Dataset, FeatureMap = import_reference()
    os.makedirs("/tmp/rat_golden", exist_ok=True)
    only = sys.argv[1:]
    for name in CASES:
        if not only or name in only:
            run_case(name, models, Dataset, FeatureMap)
    if not only:
        param_counts(models, FeatureMap)
