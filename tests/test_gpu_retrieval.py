"""-m gpu parity tests of BM25 top-K retrieval (rat_bm25_topk through the drop-in fuxictr.datasets.data_utils
.BM25_topk_retrieval_v4) against the CPU oracle (bit for bit, including the indices: both break ties towards the smaller db
index) and against the vectors the reference itself produced (values / lens bit for bit, indices by score)."""
import numpy as np
import pytest

from tests.test_oracle_golden import BM25_CASES, check_bm25_against_reference, load_bm25_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bm25():
    import rat_native
    rat_native.require_device()
    from fuxictr.datasets.data_utils import BM25_topk_retrieval_v4
    return BM25_topk_retrieval_v4


@pytest.mark.parametrize("name", BM25_CASES)
def test_bm25_matches_reference_fixtures_and_oracle(bm25, name):
    from oracle import bm25_oracle as B
    c = load_bm25_case(name)
    res = bm25(c["db"], c["qry"], exact_match_col_indices=c["exm"], qry_batch_size=c["qbs"], db_chunk_size=50, topK=c["topK"])
    assert res.values.dtype == np.float64 and res.indices.shape == c["indices"].shape
    check_bm25_against_reference(c, res.values, res.indices, res.lens)
    v, i, n = B.bm25_topk(c["db"], c["qry"], c["exm"], c["qbs"], c["topK"])
    assert np.array_equal(res.values, v) and np.array_equal(res.indices, i) and np.array_equal(res.lens, n)


@pytest.mark.parametrize("N,Q,C,rng_db,exm,K", [(20000, 37, 13, 30, None, 5),      # few queries: the db range is split over CTAs
                                                 (70000, 300, 6, 10, [1, 3], 10),   # exact-match groups larger than K
                                                 (5, 9, 3, 4, None, 8),             # fewer db rows than K
                                                 (1, 1, 1, 2, None, 1),
                                                 (3000, 1000, 19, 6, None, 32)])    # widest supported: 19 scored columns, K = 32
def test_bm25_random_shapes_match_oracle(bm25, N, Q, C, rng_db, exm, K):
    from oracle import bm25_oracle as B
    g = np.random.default_rng(N + Q)
    db = g.integers(0, rng_db, (N, C))
    qry = g.integers(0, rng_db + 2, (Q, C))
    res = bm25(db, qry, exact_match_col_indices=exm, qry_batch_size=128, topK=K)
    v, i, n = B.bm25_topk(db, qry, exm, 128, K)
    assert np.array_equal(res.lens, n)
    assert np.array_equal(res.values, v)
    assert np.array_equal(res.indices, i)


def test_bm25_is_deterministic_and_independent_of_batching(bm25):
    g = np.random.default_rng(7)
    db = g.integers(0, 9, (30000, 8)); qry = g.integers(0, 9, (500, 8))
    a = bm25(db, qry, topK=10)
    b = bm25(db, qry, topK=10)
    assert np.array_equal(a.values, b.values) and np.array_equal(a.indices, b.indices)
    # every query value occurs in the db here, so the per-batch integer-IDF quirk cannot fire: any batching gives the same rows
    c = bm25(db, qry, qry_batch_size=64, db_chunk_size=1000, topK=10)
    assert np.array_equal(a.values, c.values) and np.array_equal(a.indices, c.indices) and np.array_equal(a.lens, c.lens)


def test_bm25_rejects_unsupported_shapes(bm25):
    import rat_native
    db = np.zeros((10, 30), dtype=np.int64)
    with pytest.raises(rat_native.RatError):
        bm25(db, db[:2], topK=5)
    with pytest.raises(rat_native.RatError):
        bm25(db[:, :4], db[:2, :4], topK=40)


@pytest.mark.parametrize("label_wise", [False, True])
@pytest.mark.parametrize("pool", ["self", "external"])
def test_pre_retrieval_driver_matches_the_reference_recipe(tmp_path, label_wise, pool):
    """the retrieval file a DataGenerator produces when none is cached: 3-fold self retrieval / external pool, label-wise or not
    (reference driver fuxictr/pytorch/data_generator.py:115-212, restated here around the oracle)"""
    import rat_native
    rat_native.require_device()
    from oracle import bm25_oracle as B
    from fuxictr.pytorch.data_generator import DataGenerator, _load_array
    g = np.random.default_rng(3)
    data = np.concatenate([g.integers(0, 7, (301, 5)), g.integers(0, 2, (301, 1))], axis=1).astype(np.float64)
    pool_arr = np.concatenate([g.integers(0, 7, (200, 5)), g.integers(0, 2, (200, 1))], axis=1).astype(np.float64)
    np.savez(tmp_path / "train.npz", data=data)
    np.savez(tmp_path / "pool.npz", data=pool_arr)
    cols, K = [0, 2, 3], 4
    cfg = dict(pre_retrieval=True, split_type="3-fold" if pool == "self" else "sequential", used_col_indices=cols, label_wise=label_wise,
               topK=K, qry_batch_size=64, db_chunk_size=100, device="cuda:0")
    gen = DataGenerator(str(tmp_path / "train.h5"), batch_size=32, retrieval_configs=dict(cfg),
                        retrieval_pool_fname="self" if pool == "self" else str(tmp_path / "pool.h5"), retrieval_augmented=True,
                        num_workers=0)
    got = [_load_array(str(tmp_path / f"retrieval_{K}_train.h5"), k) for k in ("indices", "values", "lens")]

    def topk(db, qry):
        return B.bm25_topk(db, qry, None, 64, K)

    def by_label(db, labels, qry, remap):
        outs = []
        for sel in (np.nonzero(labels)[0], np.nonzero(1 - labels)[0]):
            v, i, n = topk(db[sel], qry)
            outs.append((remap(sel[i]), v, n))
        return (np.concatenate([outs[0][0], outs[1][0]], -1), np.concatenate([outs[0][1], outs[1][1]], -1),
                np.stack([outs[0][2], outs[1][2]], -1))
    arr, lab = data[:, cols].astype(int), data[:, -1].astype(int)
    if pool == "self":
        fs = int(np.ceil(len(arr) / 3))
        parts = []
        for fi in range(3):
            lo, hi = fi * fs, (fi + 1) * fs
            db = np.concatenate([arr[:lo], arr[hi:]]); dbi = np.concatenate([np.arange(lo), np.arange(hi, len(arr))])
            if label_wise:
                parts.append(by_label(db, np.concatenate([lab[:lo], lab[hi:]]), arr[lo:hi], lambda x: dbi[x]))
            else:
                v, i, n = topk(db, arr[lo:hi])
                parts.append((dbi[i], v, n))
        want = [np.concatenate([p[j] for p in parts]) for j in range(3)]
    else:
        db = pool_arr[:, cols].astype(int)
        if label_wise:
            want = list(by_label(db, pool_arr[:, -1].astype(int), arr, lambda x: x))
        else:
            v, i, n = topk(db, arr)
            want = [i, v, n]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    X, y, vals, lens = next(iter(gen))              # the wire batch is served from the freshly computed neighbours
    assert X.shape[1] == 1 + (2 * K if label_wise else K) and vals.shape[1] == X.shape[1] - 1
