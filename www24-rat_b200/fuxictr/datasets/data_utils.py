"""BM25 top-K retrieval of the retrieval-augmented data pipeline, on the B200.

Drop-in for the reference's `BM25_topk_retrieval_v4` (fuxictr/datasets/data_utils.py:773-1064; aliased as
`BM25_topk_retrieval` in fuxictr/pytorch/data_generator.py:21-22): same arguments, same `ResultsNameTuple(values, indices,
lens)` of numpy arrays.  The host side keeps what the reference does with pandas on the host (value counts -> IDF tables,
exact-match group sizes per query batch); the Q x N compare-scan and the top-K run in `rat_bm25_topk` (csrc/retrieval.cu).
There is no CPU fallback: without the CUDA library / an sm_100 device the call raises.

Ties: the reference ranks with torch.topk, whose order among equal scores is implementation defined; here equal scores are
ordered by db index, so `values` / `lens` equal the reference's bit for bit while `indices` may name a different row of the
SAME score.  `db_chunk_size`, `device`, `enable_clean` are accepted and ignored (the kernel streams the db once).
"""
from collections import namedtuple

import numpy as np

ResultsNamedTuple = namedtuple("ResultsNameTuple", ["values", "indices", "lens"])


def _idf_tables(db_cols):
    """per scored column: (sorted values, log(N / count))   (reference data_utils.py:879-887)"""
    N = len(db_cols)
    out = []
    for c in range(db_cols.shape[1]):
        vals, cnt = np.unique(db_cols[:, c], return_counts=True)
        out.append((vals, np.log(N / cnt)))
    return out


def _idf_of_queries(tables, qry_cols):
    """IDF of the query's own value in every scored column for ONE query batch, 0 for values the db never holds
    (map_data_to_IDF_v1, reference data_utils.py:842-846).  The reference maps through np.vectorize, which takes its output
    dtype from the first element: when the first query of the batch holds an unseen value the column comes back as integers.
    Reproduced, because the retrieval files of the reference carry those scores."""
    out = np.zeros(qry_cols.shape, dtype=np.float64)
    for c, (vals, idf) in enumerate(tables):
        pos = np.clip(np.searchsorted(vals, qry_cols[:, c]), 0, len(vals) - 1)
        hit = vals[pos] == qry_cols[:, c]
        col = np.where(hit, idf[pos], 0.0)
        if len(qry_cols) and not hit[0]:
            col = np.trunc(col)
        out[:, c] = col
    return out


def _dense_ids(db_cols, qry_cols):
    """int32 copies; ids outside int32 are re-coded per column (equality is all the kernel needs)"""
    if len(db_cols) and (db_cols.min() < -2**31 + 2 or db_cols.max() > 2**31 - 2) or \
            len(qry_cols) and (qry_cols.min() < -2**31 + 2 or qry_cols.max() > 2**31 - 2):
        db2, q2 = np.empty(db_cols.shape, np.int32), np.empty(qry_cols.shape, np.int32)
        for c in range(db_cols.shape[1]):
            vals = np.unique(np.concatenate([db_cols[:, c], qry_cols[:, c]]))
            db2[:, c] = np.searchsorted(vals, db_cols[:, c]); q2[:, c] = np.searchsorted(vals, qry_cols[:, c])
        return db2, q2
    return np.ascontiguousarray(db_cols, dtype=np.int32), np.ascontiguousarray(qry_cols, dtype=np.int32)


def _device_topk(rn, torch, dev, db_d, N, qry, idf, E, F, K, unit, prefer_last):
    Q = len(qry)
    q_d = torch.from_numpy(np.ascontiguousarray(qry)).to(dev)
    w_d = torch.from_numpy(np.ascontiguousarray(idf.reshape(Q, max(F, 0)) if F else np.zeros((Q, 1)))).to(dev)
    vals = torch.zeros(Q, K, dtype=torch.float64, device=dev)
    inds = torch.full((Q, K), -1, dtype=torch.int64, device=dev)
    lens = torch.zeros(Q, dtype=torch.int64, device=dev)
    nb = int(rn.query("rat_bm25_topk_workspace_bytes", N, Q, K))
    ws = torch.empty(max(nb // 8 + 2, 2), dtype=torch.float64, device=dev)
    rn.call("rat_bm25_topk", db_d, N, q_d, w_d, Q, E, F, K, 1 if unit else 0, 1 if prefer_last else 0, vals, inds, lens, ws,
            ws.numel() * 8, rn.current_stream())
    return vals.cpu().numpy(), inds.cpu().numpy(), lens.cpu().numpy()


def BM25_topk_retrieval_v4(db_np_data, qry_np_data, exact_match_col_indices=None, qry_batch_size=None, db_chunk_size=None,
                           device="cuda:0", topK=10, enable_clean=False, **kwargs):
    import torch
    import rat_native as rn
    rn.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    db = np.asarray(db_np_data); qry = np.asarray(qry_np_data)
    Q, C = len(qry), db.shape[1]
    exm = list(exact_match_col_indices) if exact_match_col_indices else []
    rest = [c for c in range(C) if c not in exm]
    E, F = len(exm), len(rest)
    db_o, qry_o = _dense_ids(db[:, exm + rest], qry[:, exm + rest])       # exact-match columns first
    tables = _idf_tables(db[:, rest]) if F else []
    qry_rest = qry[:, rest]
    db_d = torch.from_numpy(db_o).to(dev)
    N = len(db_o)
    values = np.zeros((Q, topK), dtype=float)
    indices = np.full((Q, topK), -1, dtype=int)
    lens = np.zeros(Q, dtype=int)
    if Q == 0:
        return ResultsNamedTuple(values, indices, lens)
    qbs = Q if qry_batch_size is None else qry_batch_size
    if not exm:
        for q0 in range(0, Q, qbs):
            sl = slice(q0, min(Q, q0 + qbs))
            v, i, n = _device_topk(rn, torch, dev, db_d, N, qry_o[sl], _idf_of_queries(tables, qry_rest[sl]), 0, F, topK,
                                   False, False)
            values[sl], indices[sl], lens[sl] = v, i, n
        return ResultsNamedTuple(values, indices, lens)
    # exact-match mode: group sizes first (unit scores, K = 1 is enough to know "has a match"; sizes from a unit top-K of
    # K entries would cap at K, so count on the host like the reference's groupby does)
    keys_db = np.ascontiguousarray(db_o[:, :E])
    keys_q = np.ascontiguousarray(qry_o[:, :E])
    uniq, inv, cnt = np.unique(keys_db, axis=0, return_inverse=True, return_counts=True)
    lut = {tuple(k): c for k, c in zip(map(tuple, uniq), cnt)}
    sizes = np.array([lut.get(tuple(k), 0) for k in keys_q])
    for q0 in range(0, Q, qbs):
        sl = slice(q0, min(Q, q0 + qbs))
        matched = sizes[sl] > 0
        if not matched.any():
            continue
        unit = sizes[sl][matched].max() <= topK or F == 0                 # reference data_utils.py:912-917 / :1033-1038
        rows = np.nonzero(matched)[0] + q0
        idf = _idf_of_queries(tables, qry_rest[rows]) if F else np.zeros((len(rows), 0))
        v, i, n = _device_topk(rn, torch, dev, db_d, N, qry_o[rows], idf, E, F, topK, unit, unit)
        if unit:                                                            # the last K members of the group, ascending
            for r in range(len(rows)):
                i[r, :n[r]] = i[r, :n[r]][::-1]
        values[rows], indices[rows], lens[rows] = v, i, n
    return ResultsNamedTuple(values, indices, lens)


BM25_topk_retrieval = BM25_topk_retrieval_v4
