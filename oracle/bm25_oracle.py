"""CPU oracle for BM25 top-K retrieval -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's BM25_topk_retrieval_v4 (fuxictr/datasets/data_utils.py:773-1064, paths relative to
/root/reference).  Only tests/ may import it.  Pinned by tests/golden/bm25_*.npz, which tests/golden/make_golden_bm25.py
produced by RUNNING the reference function (tests/test_oracle_golden.py::test_bm25_oracle_matches_reference_fixtures).

What is pinned exactly: `values` and `lens` bit for bit.  `indices` cannot be pinned bit for bit: the reference ranks with
torch.topk / torch.sort, whose order among EQUAL scores is implementation defined (and BM25 scores over categorical ids tie all
the time), so the fixture test checks that every returned index is a db row whose score is the returned value and that no
row with a strictly larger score was left out.  This oracle (and the CUDA kernel) break ties towards the smaller db index.
"""
import numpy as np


def idf_tables(db):
    """per column: value -> log(N / count)   (data_utils.py:879-887; N / count in float64, np.log)"""
    N = len(db)
    out = []
    for c in range(db.shape[1]):
        vals, cnt = np.unique(db[:, c], return_counts=True)
        out.append((vals, np.log(N / cnt)))
    return out


def map_idf(qry, tables):
    """IDF of the query's own value per column, 0 when the db never holds it (map_data_to_IDF_v1, data_utils.py:842-846).
    Called once per QUERY BATCH, like the reference.  Reference quirk reproduced on purpose: the mapping goes through
    np.vectorize, which takes its output dtype from the FIRST element it evaluates -- when the first query of the batch holds
    a value the db never has, `.get(x, 0)` returns the int 0 and the whole column of that batch is cast to integers
    (every IDF truncated towards zero)."""
    out = np.zeros(qry.shape, dtype=np.float64)
    for c, (vals, idf) in enumerate(tables):
        pos = np.searchsorted(vals, qry[:, c])
        pos = np.clip(pos, 0, len(vals) - 1)
        hit = vals[pos] == qry[:, c]
        col = np.where(hit, idf[pos], 0.0)
        if len(qry) and not hit[0]:
            col = np.trunc(col)
        out[:, c] = col
    return out


def _scores(q, w, db):
    """sum_f (q == db) * w in the association torch's float64 .sum(-1) uses for fewer than 20 contiguous elements (found by
    enumerating all summation trees against torch on this container, F = 1..19): four interleaved accumulators over the
    full groups of four, the remaining < 4 elements summed in order and added to accumulator 0, then ((a0 + a1) + a2) + a3."""
    F = db.shape[1]
    nfull = (F // 4) * 4
    acc = [np.zeros(len(db), dtype=np.float64) for _ in range(4)]
    term = lambda f: np.where(db[:, f] == q[f], w[f], 0.0)
    for f in range(nfull):
        acc[f % 4] = acc[f % 4] + term(f)
    if nfull < F:
        tail = term(nfull)
        for f in range(nfull + 1, F):
            tail = tail + term(f)
        acc[0] = acc[0] + tail
    return ((acc[0] + acc[1]) + acc[2]) + acc[3]


def _topk(score, K, prefer_last=False):
    """K best (score desc, then db index asc -- desc with prefer_last); zero scores are no matches"""
    n = len(score)
    idx = np.arange(n)
    order = np.lexsort((-idx if prefer_last else idx, -score))
    order = order[:K]
    order = order[score[order] > 0]
    vals = np.zeros(K); inds = np.full(K, -1, dtype=np.int64)
    vals[:len(order)] = score[order]; inds[:len(order)] = order
    return vals, inds, len(order)


def bm25_topk(db, qry, exact_match_col_indices=None, qry_batch_size=None, topK=10):
    """(values [Q,K] f64, indices [Q,K] i64, lens [Q] i64) -- same contract as data_utils.py:1064."""
    db = np.asarray(db); qry = np.asarray(qry)
    Q = len(qry)
    values = np.zeros((Q, topK)); indices = np.full((Q, topK), -1, dtype=np.int64); lens = np.zeros(Q, dtype=np.int64)
    exm = list(exact_match_col_indices) if exact_match_col_indices else []
    rest = [c for c in range(db.shape[1]) if c not in exm]
    db_r, qry_r = db[:, rest], qry[:, rest]
    tables = idf_tables(db_r) if rest else []
    w = np.zeros((Q, len(rest)))
    qbs = Q if qry_batch_size is None else qry_batch_size
    for q0 in range(0, Q, qbs):
        rows = range(q0, min(Q, q0 + qbs))
        cand = {}
        for b in rows:
            cand[b] = np.all(db[:, exm] == qry[b, exm], axis=1) if exm else np.ones(len(db), dtype=bool)
        if exm:
            sizes = [int(cand[b].sum()) for b in rows if cand[b].any()]
            if not sizes:
                continue
            unit = max(sizes) <= topK or not rest          # data_utils.py:912-917 / :1033-1038
        else:
            unit = False
        live = [b for b in rows if (not exm) or cand[b].any()]     # the batch the reference maps to IDFs (:920-925)
        if rest and live:
            w[live] = map_idf(qry_r[live], tables)
        for b in rows:
            if exm and not cand[b].any():
                continue
            if unit:
                s = cand[b].astype(np.float64)
                # pad_sequences(truncating="pre") keeps the LAST K members of a group; they stay in ascending index order
                v, i, n = _topk(s, topK, prefer_last=True)
                i[:n] = i[:n][::-1]
            else:
                s = _scores(qry_r[b], w[b], db_r)
                if exm:
                    s = (s + 1.0) * cand[b]
                v, i, n = _topk(s, topK)
            values[b], indices[b], lens[b] = v, i, n
    return values, indices, lens


def all_scores(db, qry, exact_match_col_indices=None, qry_batch_size=None, topK=10):
    """dense [Q, N] score matrix with the same batch-wise quirks (fixture property checks: every index the reference returned
    must be a row with exactly the returned score, and nothing with a larger score may be missing)"""
    db = np.asarray(db); qry = np.asarray(qry)
    Q = len(qry)
    exm = list(exact_match_col_indices) if exact_match_col_indices else []
    rest = [c for c in range(db.shape[1]) if c not in exm]
    db_r, qry_r = db[:, rest], qry[:, rest]
    tables = idf_tables(db_r) if rest else []
    out = np.zeros((Q, len(db)))
    qbs = Q if qry_batch_size is None else qry_batch_size
    for q0 in range(0, Q, qbs):
        rows = list(range(q0, min(Q, q0 + qbs)))
        cand = {b: (np.all(db[:, exm] == qry[b, exm], axis=1) if exm else np.ones(len(db), dtype=bool)) for b in rows}
        live = [b for b in rows if cand[b].any()]
        if not live:
            continue
        unit = bool(exm) and (max(int(cand[b].sum()) for b in live) <= topK or not rest)
        w = map_idf(qry_r[live], tables) if rest else None
        for j, b in enumerate(live):
            if unit:
                out[b] = cand[b].astype(np.float64)
            else:
                s = _scores(qry_r[b], w[j], db_r)
                out[b] = (s + 1.0) * cand[b] if exm else s
    return out
