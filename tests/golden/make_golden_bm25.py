"""Golden vectors for BM25 top-K retrieval, produced by RUNNING the reference's BM25_topk_retrieval_v4
(/root/reference/fuxictr/datasets/data_utils.py:773-1064) in the build container.

    python tests/golden/make_golden_bm25.py        -> tests/golden/bm25_<case>.npz

The reference imports tensorflow only for keras' pad_sequences; the stub below restates its documented semantics
(padding / truncating 'pre' | 'post', default truncating='pre').  Case "selfcheck" is the reference's own __main__
self-check configuration (data_utils.py:1287-1312, same seed, same shapes)."""
import importlib
import importlib.machinery as mm
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    sequences = list(sequences)
    lengths = [len(s) for s in sequences]
    if maxlen is None:
        maxlen = max(lengths) if lengths else 0
    x = np.full((len(sequences), maxlen), value, dtype=dtype)
    for i, s in enumerate(sequences):
        s = np.asarray(s)
        if not len(s):
            continue
        trunc = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        if padding == "post":
            x[i, :len(trunc)] = trunc
        else:
            x[i, -len(trunc):] = trunc
    return x


def import_reference_bm25():
    sys.path.insert(0, REF)

    def _stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = mm.ModuleSpec(name, None)
        m.__path__ = []
        m.__dict__.update(attrs)
        sys.modules[name] = m

    for n in ["h5py", "dgl", "dgl.function", "dgl.nn", "dgl.nn.functional", "matplotlib", "matplotlib.pyplot", "thop",
              "tensorflow", "tensorflow.keras", "tensorflow.keras.utils"]:
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    sys.modules["tensorflow.keras.utils"].__dict__["pad_sequences"] = pad_sequences
    np.Inf = np.inf
    from fuxictr.datasets import data_utils
    return data_utils.BM25_topk_retrieval_v4


CASES = {
    # name: (N, Q, C, id range db, id range qry, exact-match columns, qry_batch_size, db_chunk_size, topK, seed)
    "plain":      (3000, 200, 6, 8, 8, None, 64, None, 10, 1),
    "chunked":    (3000, 200, 6, 8, 8, None, 64, 500, 10, 1),
    "selfcheck":  (2000, 100, 5, 5, 5, [0, 4], 50, 50, 10, 0),
    "exm_small":  (500, 120, 4, 12, 12, [0, 1], 40, None, 10, 2),
    "exm_only":   (400, 60, 2, 6, 7, [0, 1], 20, None, 5, 3),
    "sparse":     (50, 40, 3, 50, 60, None, None, None, 5, 4),
    "kkbox_like": (4000, 150, 13, 40, 40, None, 50, 1000, 5, 5),
}


def make_inputs(name):
    N, Q, C, rd, rq, exm, qbs, chunk, K, seed = CASES[name]
    np.random.seed(seed)
    db = np.random.randint(0, rd, (N, C))
    qry = np.random.randint(0, rq, (Q, C))
    return db, qry, exm, qbs, chunk, K


if __name__ == "__main__":
    fn = import_reference_bm25()
    for name in CASES:
        db, qry, exm, qbs, chunk, K = make_inputs(name)
        res = fn(db, qry, exact_match_col_indices=exm, qry_batch_size=qbs, db_chunk_size=chunk, device="cpu", topK=K)
        np.savez_compressed(os.path.join(HERE, f"bm25_{name}.npz"), db=db.astype(np.int64), qry=qry.astype(np.int64),
                            exm=np.array(exm if exm else [], dtype=np.int64), qbs=np.array(-1 if qbs is None else qbs),
                            topK=np.array(K), values=np.asarray(res.values, dtype=np.float64),
                            indices=np.asarray(res.indices, dtype=np.int64), lens=np.asarray(res.lens, dtype=np.int64))
        print(name, "values", res.values.shape, "lens mean", float(np.mean(res.lens)), "zeros", int((res.values == 0).sum()))
