"""csv -> id preparation against the reference's own FeatureEncoder / build_dataset.

tests/golden/encoder_<case>.npz hold what the REFERENCE produced on seeded synthetic csv files
(tests/golden/make_golden_encoder.py); here our fuxictr.datasets package runs on the regenerated csv files and must
reproduce feature_map.json, every vocabulary and every saved id block bit for bit.  CPU only."""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_golden_encoder as G  # noqa: E402


def _load_block(path):
    if path.endswith(".npz"):
        with np.load(path) as z:
            return z["data"]
    import h5py
    with h5py.File(path, "r") as hf:
        return hf["data"][()]


@pytest.mark.parametrize("case", sorted(G.CASES))
def test_build_dataset_reproduces_the_reference(case, tmp_path):
    from fuxictr import datasets
    gold = np.load(os.path.join(ROOT, "tests", "golden", "encoder_{}.npz".format(case)))
    enc = G.run_encoder(datasets, case, str(tmp_path), None)
    # feature_map.json: same content AND same key order (the gather's column layout follows it)
    want_fm = json.loads(str(gold["feature_map"]))
    got_fm = json.load(open(enc.json_file))
    assert got_fm == want_fm
    assert list(got_fm["feature_specs"]) == list(want_fm["feature_specs"])
    assert json.dumps(got_fm, indent=4) == str(gold["feature_map"])
    # vocabularies
    want_vocab = json.loads(str(gold["vocab"]))
    got_vocab = {n[:-len("_tokenizer")]: {str(k): int(v) for k, v in t.vocab.items()} for n, t in enc.encoders.items()
                 if n.endswith("_tokenizer")}
    assert got_vocab == want_vocab
    # every block the reference saved, bit for bit, and nothing else
    files = {os.path.splitext(os.path.basename(p))[0]: p for p in glob.glob(os.path.join(enc.data_dir, "*"))
             if p.endswith((".npz", ".h5"))}
    want_blocks = sorted(k for k in gold.files if k not in ("feature_map", "vocab"))
    assert sorted(files) == want_blocks
    # columns of normalised numeric features are float arithmetic (sklearn scalers in the reference): 1e-12 relative
    fcols = [spec["index"] for spec in want_fm["feature_specs"].values() if spec["type"] == "numeric"]
    icols = [c for c in range(want_fm["input_length"] + 1) if c not in fcols]
    for name in want_blocks:
        got = _load_block(files[name])
        assert got.dtype == np.float64 and got.shape == gold[name].shape, name
        assert np.array_equal(got[:, icols], gold[name][:, icols]), name
        np.testing.assert_allclose(got[:, fcols], gold[name][:, fcols], rtol=1e-12, atol=1e-12, err_msg=name)
    # the pickled encoder reloads and transforms identically (run_expid.py's csv branch on a prepared directory)
    enc2 = getattr(datasets, G.CASES[case][0]).FeatureEncoder(feature_cols=enc.feature_cols, label_col=enc.label_col,
                                                               dataset_id=case, data_root=os.path.dirname(enc.data_dir))
    enc2 = enc2.load_pickle()
    assert enc2.feature_map.feature_specs == enc.feature_map.feature_specs


def test_tokenizer_contracts():
    from fuxictr.preprocess import Tokenizer, pad_sequences
    t = Tokenizer(min_freq=2, na_value="")
    t.fit_on_texts(np.array(["b", "a", "b", "a", "c", "", "", "a"], dtype=object))
    assert t.vocab == {"a": 1, "b": 2, "__OOV__": 0} and t.vocab_size == 3          # count desc, token asc; "" and rare "c" out
    assert t.encode_category(["a", "c", np.nan, "b"]).tolist() == [1, 0, 0, 2]
    s = Tokenizer(splitter="^", na_value="", padding="post")
    s.fit_on_texts(np.array(["x^y", "y", "x^y^z^w"], dtype=object), use_padding=True)
    assert s.max_len == 4 and s.vocab["__PAD__"] == s.vocab_size - 1 == 5
    assert s.encode_sequence(np.array(["y^q", "", np.nan], dtype=object)).tolist() == [[1, 0, 5, 5], [5, 5, 5, 5], [5, 5, 5, 5]]
    s.max_len, s.padding = 2, "pre"                                               # keep the LAST two, right aligned
    assert s.encode_sequence(np.array(["x^y^z", "x"], dtype=object)).tolist() == [[1, s.vocab["z"]], [5, s.vocab["x"]]]
    s.padding = "post"                                                            # keep the FIRST two, left aligned
    assert s.encode_sequence(np.array(["x^y^z", "x"], dtype=object)).tolist() == [[s.vocab["x"], 1], [s.vocab["x"], 5]]
    assert pad_sequences([[1, 2, 3], [], [4]], maxlen=2, value=9, padding="post", truncating="post").tolist() == [[1, 2], [9, 9], [4, 9]]
    assert pad_sequences([[1, 2, 3], [4]], maxlen=2, value=9).tolist() == [[2, 3], [9, 4]]


def test_split_train_test_matches_reference_semantics():
    import pandas as pd
    from fuxictr.datasets import split_train_test
    df = pd.DataFrame({"a": np.arange(100)})
    tr, va, te = split_train_test(df, valid_size=0.1, test_size=20)
    assert te["a"].tolist() == list(range(80, 100)) and va["a"].tolist() == list(range(70, 80)) and len(tr) == 70
    pool, tr2, _ = split_train_test(train_ddf=df, valid_size=0.8)                  # build_dataset's pool_ratio = 0.2 split
    assert pool["a"].tolist() == list(range(20)) and tr2["a"].tolist() == list(range(20, 100))
