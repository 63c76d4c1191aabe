"""Golden fixtures for the csv -> id preparation, produced by RUNNING the reference.

    python tests/golden/make_golden_encoder.py        # needs /root/reference (build container only)

For each case the synthetic csv files are written from a seed (tests regenerate them with `write_case_csvs`), the
REFERENCE's FeatureEncoder (fuxictr/datasets/kkbox.py / tmall.py subclasses) and build_dataset run on them with `save_hdf5`
replaced by a capture function (h5py is not installed here; module stubs as in make_golden.py / SURVEY.md Appendix C), and
the outputs are stored in tests/golden/encoder_<case>.npz:
    feature_map  json text of feature_map.json
    vocab        json {feature: {token: id}} of every tokenizer
    <block>      every array build_dataset saved ("train", "valid", "test", "retrieval_pool", "train_part_0", ...)
"""
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _rand_tokens(rng, n, vocab, p_nan=0.0, zipf=1.3):
    ids = np.minimum(rng.zipf(zipf, size=n), vocab) - 1
    out = np.array(["v%d" % i for i in ids], dtype=object)
    if p_nan > 0:
        out[rng.random(n) < p_nan] = np.nan
    return out


def _rand_seq(rng, n, vocab, max_tokens, sep, p_empty=0.1):
    out = []
    for _ in range(n):
        if rng.random() < p_empty:
            out.append(np.nan)
            continue
        k = int(rng.integers(1, max_tokens + 1))
        out.append(sep.join("g%d" % (min(int(z), vocab) - 1) for z in rng.zipf(1.4, size=k)))
    return np.array(out, dtype=object)


def kkbox_frame(rng, n):
    df = pd.DataFrame({
        "label": rng.integers(0, 2, size=n),
        "msno": _rand_tokens(rng, n, 60),
        "song_id": _rand_tokens(rng, n, 200, zipf=1.15),
        "city": _rand_tokens(rng, n, 8, p_nan=0.05),
        "genre_ids": _rand_seq(rng, n, 30, 5, " "),
        "artist_name": _rand_seq(rng, n, 50, 4, " ", p_empty=0.02),
        "isrc": np.array([("%s%03d" % (c, i)) for c, i in zip(rng.choice(["US", "TW", "GB", "JP"], size=n), rng.integers(0, 999, size=n))],
                         dtype=object),
        "bd": np.array([str(int(a)) for a in rng.integers(-5, 110, size=n)], dtype=object),
    })
    df.loc[rng.random(n) < 0.07, "isrc"] = np.nan
    return df           # no missing `bd`: the reference fills NaN with "" BEFORE the hook and float("") raises there


def tmall_frame(rng, n):
    months = rng.integers(5, 12, size=n)
    days = rng.integers(1, 29, size=n)
    return pd.DataFrame({
        "label": rng.integers(0, 2, size=n),
        "time_stamp": np.array(["%d%02d" % (m, d) for m, d in zip(months, days)], dtype=object),
        "user_id": _rand_tokens(rng, n, 300, zipf=1.1),
        "item_id": _rand_tokens(rng, n, 500, zipf=1.1),
        "cat_id": _rand_tokens(rng, n, 20),
        "age_range": _rand_tokens(rng, n, 6, p_nan=0.1),
        "weekday": np.array(["x"] * n, dtype=object),
        "weekend": np.array(["x"] * n, dtype=object),
    })


def misc_frame(rng, n):
    """numeric columns (normalizers, explicit na_value), a vocabulary shared by two id columns, a sequence that shares the
    vocabulary of a categorical column (which therefore gets the padding row), a non-default splitter"""
    price = rng.gamma(2.0, 10.0, size=n)
    price[rng.random(n) < 0.05] = np.nan
    return pd.DataFrame({
        "label": rng.integers(0, 2, size=n),
        "price": price,
        "score": rng.normal(3.0, 2.0, size=n),
        "raw": rng.integers(0, 50, size=n).astype(float),
        "user": _rand_tokens(rng, n, 40),
        "friend": _rand_tokens(rng, n, 55, p_nan=0.04),
        "item": _rand_tokens(rng, n, 80, zipf=1.2),
        "hist": _rand_seq(rng, n, 90, 6, "^", p_empty=0.08),
        "tags": _rand_seq(rng, n, 12, 3, "^", p_empty=0.0),
    }).assign(hist=lambda d: d["hist"].str.replace("g", "v"))      # history tokens live in the item vocabulary


MISC_COLS = [
    {"active": True, "dtype": "float", "name": "price", "type": "numeric", "normalizer": "StandardScaler", "na_value": 0},
    {"active": True, "dtype": "float", "name": "score", "type": "numeric", "normalizer": "MinMaxScaler"},
    {"active": True, "dtype": "float", "name": "raw", "type": "numeric"},
    {"active": True, "dtype": "str", "name": "user", "type": "categorical", "embedding_dim": 8},
    {"active": True, "dtype": "str", "name": "friend", "type": "categorical", "share_embedding": "user", "na_value": "nobody"},
    {"active": True, "dtype": "str", "name": "item", "type": "categorical", "min_categr_count": 1},
    {"active": True, "dtype": "str", "name": "hist", "type": "sequence", "share_embedding": "item", "splitter": "^", "max_len": 4,
     "padding": "pre", "encoder": "MaskedAveragePooling"},
    {"active": True, "dtype": "str", "name": "tags", "type": "sequence", "splitter": "^"},
]

KKBOX_COLS = [
    {"active": True, "dtype": "str", "name": ["msno", "song_id", "city"], "type": "categorical"},
    {"active": True, "dtype": "str", "encoder": "MaskedSumPooling", "max_len": 3, "name": "genre_ids", "type": "sequence"},
    {"active": True, "dtype": "str", "encoder": "MaskedSumPooling", "name": "artist_name", "type": "sequence", "padding": "pre"},
    {"active": True, "dtype": "str", "name": "isrc", "preprocess": "extract_country_code", "type": "categorical"},
    {"active": True, "dtype": "str", "name": "bd", "preprocess": "bucketize_age", "type": "categorical", "min_categr_count": 1},
]
TMALL_COLS = [
    {"active": False, "dtype": "str", "name": "time_stamp", "type": "categorical"},
    {"active": True, "dtype": "str", "name": ["user_id", "item_id", "cat_id", "age_range"], "type": "categorical"},
    {"active": True, "dtype": "str", "name": "weekday", "preprocess": "convert_weekday", "type": "categorical"},
    {"active": True, "dtype": "str", "name": "weekend", "preprocess": "convert_weekend", "type": "categorical"},
]
LABEL = {"dtype": "float", "name": "label"}

# case -> (dataset module, frame builder, feature_cols, sizes (train, valid, test, pool), build_dataset kwargs)
CASES = {
    "kkbox_pool_ratio": ("kkbox", kkbox_frame, KKBOX_COLS, (600, 150, 150, 0),
                         dict(min_categr_count=2, retrieval_configs={"split_type": "sequential", "pool_ratio": 0.2})),
    "kkbox_10fold_blocks": ("kkbox", kkbox_frame, KKBOX_COLS, (500, 120, 0, 0),
                            dict(min_categr_count=3, data_block_size=200,
                                 retrieval_configs={"split_type": "10-fold", "pool_ratio": 0.2})),
    "kkbox_split_sizes": ("kkbox", kkbox_frame, KKBOX_COLS, (700, 0, 0, 0),
                          dict(min_categr_count=2, valid_size=0.1, test_size=70)),
    "misc_features": ("kkbox", misc_frame, MISC_COLS, (400, 100, 0, 0), dict(min_categr_count=2)),
    "tmall_pool_file": ("tmall", tmall_frame, TMALL_COLS, (500, 0, 140, 300),
                        dict(min_categr_count=2, retrieval_configs={"split_type": "sequential"})),
}


def write_case_csvs(case, out_dir):
    """deterministic csv files of a case; returns the build_dataset keyword arguments that point at them"""
    module, frame, cols, (n_tr, n_va, n_te, n_pool), kw = CASES[case]
    rng = np.random.default_rng(sum(map(ord, case)))
    os.makedirs(out_dir, exist_ok=True)
    paths = {}
    for tag, n in (("train", n_tr), ("valid", n_va), ("test", n_te), ("retrieval_pool", n_pool)):
        if n:
            paths[tag] = os.path.join(out_dir, tag + ".csv")
            frame(rng, n).to_csv(paths[tag], index=False)
    kw = json.loads(json.dumps(kw))
    args = dict(train_data=paths["train"], valid_data=paths.get("valid"), test_data=paths.get("test"))
    if "retrieval_pool" in paths:
        kw["retrieval_configs"]["retrieval_pool_data"] = paths["retrieval_pool"]
    if case == "tmall_pool_file":
        args["valid_data"] = paths["test"]          # the tmall configs validate on the test file
    args.update(kw)
    return module, json.loads(json.dumps(cols)), args


def run_encoder(datasets_pkg, case, work_dir, capture):
    """FeatureEncoder + build_dataset of `datasets_pkg` (the reference's or ours) on the case; capture(path, array) per block"""
    module, cols, args = write_case_csvs(case, os.path.join(work_dir, "csv"))
    enc = getattr(datasets_pkg, module).FeatureEncoder(feature_cols=cols, label_col=dict(LABEL), dataset_id=case,
                                                       data_root=os.path.join(work_dir, "data"))
    datasets_pkg.build_dataset(enc, **args)
    return enc


def main():
    import tempfile
    from make_golden import import_reference
    import_reference()
    from fuxictr import datasets as ref_datasets
    from fuxictr.datasets import data_utils as ref_du
    for case in CASES:
        blocks = {}
        def capture(data_array, data_path, key="data"):
            blocks[os.path.splitext(os.path.basename(data_path))[0]] = np.asarray(data_array)
        ref_du.save_hdf5 = capture
        ref_datasets.save_hdf5 = capture
        with tempfile.TemporaryDirectory() as tmp:
            np.random.seed(0)
            enc = run_encoder(ref_datasets, case, tmp, capture)
            fm_text = open(enc.json_file).read()
        vocab = {name[:-len("_tokenizer")]: {str(k): int(v) for k, v in tok.vocab.items()}
                 for name, tok in enc.encoders.items() if name.endswith("_tokenizer")}
        out = os.path.join(HERE, "encoder_{}.npz".format(case))
        np.savez_compressed(out, feature_map=np.array(fm_text), vocab=np.array(json.dumps(vocab, sort_keys=True)), **blocks)
        print(case, {k: v.shape for k, v in blocks.items()}, "num_features", enc.feature_map.num_features)


if __name__ == "__main__":
    main()
