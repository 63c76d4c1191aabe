"""Synthetic dataset shapes of the three shipped RAT_m2 experiments + synthetic data generators.

The reference ships no feature_map.json, so per-field vocabulary sizes are not recoverable; the TOTALS
(90,239 / 92,247 / 1,529,680 rows) are the ones that reproduce the logged parameter counts
(1,337,241 / 4,714,649 / 16,970,282) and the per-field split below is a documented long-tail guess."""
from collections import OrderedDict

import numpy as np

_KKBOX = [("msno", 30000), ("song_id", 48000), ("source_system_tab", 10), ("source_screen_name", 22),
          ("source_type", 14), ("city", 23), ("gender", 4), ("registered_via", 7), ("language", 12),
          ("genre_ids", 170), ("artist_name", 11800), ("isrc", 110), ("bd", 75)]
_TMALL = [("user_id", 400000), ("item_id", 1100000), ("cat_id", 1600), ("seller_id", 5000), ("brand_id", 8400),
          ("action_type", 5), ("age_range", 10), ("gender", 4)]
_ML = [("user_id", 17), ("item_id", 24), ("tag_id", 49)]

SHAPES = {
    "ml": dict(dataset_id="movielenslatest_x1_10fold_retrieval", total_vocab=90239, weights=_ML, sequence=(),
               hp=dict(embedding_dim=10, num_heads=2, scale_dim=4, dnn_hidden_units=[400, 400, 400], batch_norm=False,
                       emb_dropout=0.0, net_dropout=0, embedding_regularizer=0.03, net_regularizer=0),
               pos_ratio=0.33, pool_rows=1404801),
    "kkbox": dict(dataset_id="kkbox_x1_10fold_retrieval", total_vocab=92247, weights=_KKBOX,
                  sequence=("genre_ids", "artist_name"),
                  hp=dict(embedding_dim=40, num_heads=8, scale_dim=2, dnn_hidden_units=[400, 400, 400], batch_norm=True,
                          emb_dropout=0.1, net_dropout=0, embedding_regularizer=0.0005),
                  pos_ratio=0.5, pool_rows=5901932),
    "tmall": dict(dataset_id="tmall_x1_002_retrieval", total_vocab=1529680, weights=_TMALL, sequence=(),
                  hp=dict(embedding_dim=10, num_heads=32, scale_dim=2, dnn_hidden_units=[200, 80], batch_norm=True,
                          emb_dropout=0.1, net_dropout=0.08, embedding_regularizer=0.07),
                  pos_ratio=0.5, pool_rows=20038830),
}
COMMON_HP = dict(task="binary_classification", learning_rate=1e-3, dnn_activations="relu", use_wide=True, depth=4,
                 dim_head=10, dropout=0.0, optimizer="adam", loss="binary_crossentropy", metrics=["AUC", "logloss"],
                 monitor="AUC", monitor_mode="max", patience=2, every_x_epochs=1, save_best_only=True, verbose=0,
                 retrieval_augmented=True, seed=2021, layer_norm=True, use_scale=True, use_residual=True, pool="cls")


def split_vocab(total, weights, minimum=3):
    wsum = float(sum(w for _, w in weights))
    raw = [max(minimum, int(total * w / wsum)) for _, w in weights]
    raw[int(np.argmax(raw))] += total - sum(raw)
    return raw


def feature_specs(shape, vocab_scale=1.0):
    """OrderedDict usable as FeatureMap.feature_specs (fuxictr/features.py:36-57 layout)."""
    cfg = SHAPES[shape]
    sizes = split_vocab(int(cfg["total_vocab"] * vocab_scale), cfg["weights"])
    specs, idx = OrderedDict(), 0
    for (name, _), v in zip(cfg["weights"], sizes):
        if name in cfg["sequence"]:
            specs[name] = {"source": "", "type": "sequence", "vocab_size": v, "padding_idx": v - 1, "max_len": 3,
                           "encoder": "MaskedSumPooling", "index": [idx, idx + 1, idx + 2]}
            idx += 3
        else:
            specs[name] = {"source": "", "type": "categorical", "vocab_size": v, "index": idx}
            idx += 1
    return specs, idx


def make_feature_map(shape, vocab_scale=1.0, data_dir="/tmp/rat_synth"):
    from fuxictr.features import FeatureMap
    cfg = SHAPES[shape]
    fm = FeatureMap(cfg["dataset_id"], data_dir)
    fm.feature_specs, fm.input_length = feature_specs(shape, vocab_scale)
    fm.num_fields = len(fm.feature_specs)
    fm.num_features = sum(s["vocab_size"] for s in fm.feature_specs.values())
    return fm


def model_params(shape, K=5, model_root="/tmp/rat_synth/exps/", **over):
    cfg = SHAPES[shape]
    p = dict(COMMON_HP)
    p.update(cfg["hp"])
    p.update(model="RAT_m2", model_id="RAT_m2_" + cfg["dataset_id"], dataset_id=cfg["dataset_id"], model_root=model_root,
             batch_size=4096, retrieval_configs={"topK": K, "label_wise": False, "pre_retrieval": True,
                                                 "split_type": "10-fold"})
    p.update(over)
    return p


def synthetic_array(specs, n_rows, seed, pos_ratio=0.5, zipf_a=1.05):
    """[n_rows, L+1] float64 h5-style block: id columns then label. ids ~ clipped Zipf in [1, V-1);
    sequence fields get 1..max_len valid ids then the padding id (= vocab_size-1)."""
    rng = np.random.default_rng(seed)
    cols = []
    for name, s in specs.items():
        width = s.get("max_len", 1) if s["type"] == "sequence" else 1
        pad = s.get("padding_idx", None)
        hi = max(2, s["vocab_size"] - (1 if pad is not None else 0))
        z = rng.zipf(zipf_a, size=(n_rows, width)).astype(np.int64)
        ids = 1 + (z - 1) % (hi - 1) if hi > 2 else np.ones_like(z)
        ids = np.minimum(ids, hi - 1)
        if s["type"] == "sequence":
            nvalid = rng.integers(1, width + 1, size=n_rows)
            ids[np.arange(width)[None, :] >= nvalid[:, None]] = pad
        cols.append(ids)
    lab = (rng.random(n_rows) < pos_ratio).astype(np.int64)[:, None]
    return np.concatenate(cols + [lab], axis=1).astype(np.float64)


def synthetic_neighbours(n_query, n_pool, K, seed, missing=0.02):
    """[Q,K] int64 neighbour indices, tail-padded with -1 where BM25 would have found < K matches."""
    rng = np.random.default_rng(seed + 7)
    idx = rng.integers(0, n_pool, size=(n_query, K), dtype=np.int64)
    nmiss = (rng.random(n_query) < missing) * rng.integers(1, K + 1, size=n_query)
    idx[np.arange(K)[None, :] >= (K - nmiss)[:, None]] = -1
    return idx


def host_wire_batch(darray, pool, nbr, rows):
    """the reference wire format for a list of rows (vectorised Dataset.__getitem__ + default collate)."""
    full = np.concatenate([darray[rows][:, None, :], pool[nbr[rows]]], axis=1)
    K = nbr.shape[1]
    return (np.ascontiguousarray(full[..., :-1]), np.ascontiguousarray(full[..., -1]),
            np.zeros((len(rows), K)), (nbr[rows] >= 0).sum(1).astype(np.int64))
