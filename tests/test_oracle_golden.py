"""Pin the CPU oracle (oracle/rat_oracle.py) against fixtures produced by running the reference
itself (tests/golden/make_golden.py).  CPU-only; runs in the build container and on the GPU box."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rat_oracle as O
from tests.helpers import (CASES_M2, CASES_VAR, CASES_VAR_KKBOX, GOLDEN, add_dead_params, load_case, noise_grad_param,
                           split_state)

FWD_TOL = dict(rtol=2e-5, atol=2e-6)     # fp32 vs fp32, different op order
GRAD_TOL = dict(rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("name", CASES_M2 + CASES_VAR + CASES_VAR_KKBOX)
def test_assembly_matches_reference_dataset(name):
    """a1: numpy fancy-index assembly incl. the -1 -> last-pool-row wraparound (data_generator.py:66-78)."""
    c = load_case(name)
    z = c["z"]
    rows = np.arange(z["darray"].shape[0])
    X, y = O.assemble_batch(z["darray"], z["pool"], z["retr_indices"], rows)
    assert (z["retr_indices"] == -1).any()
    np.testing.assert_array_equal(X, z["X"])
    np.testing.assert_array_equal(y, z["y"])
    # the wrapped rows really are the last pool row
    b, k = np.argwhere(z["retr_indices"] == -1)[0]
    np.testing.assert_array_equal(X[b, 1 + k], z["pool"][-1, :-1])


@pytest.mark.parametrize("name", CASES_M2 + CASES_VAR + CASES_VAR_KKBOX)
def test_eval_forward_matches_reference(name):
    c = load_case(name)
    params, bufs = split_state(c["sd0"])
    with torch.no_grad():
        y_pred = O.forward(params, bufs, c["spec"], c["X"], c["y"], training=False)
    np.testing.assert_allclose(y_pred.numpy(), c["z"]["eval/y_pred"], **FWD_TOL)
    np.testing.assert_array_equal(c["y"][:, 0:1].float().numpy(), c["z"]["eval/y_true"])


@pytest.mark.parametrize("name", CASES_M2 + CASES_VAR + CASES_VAR_KKBOX)
def test_two_train_steps_match_reference(name):
    c = load_case(name)
    spec = c["spec"]
    params, bufs = split_state(c["sd0"])
    st = O.AdamState()
    r1 = O.train_step(params, bufs, spec, st, c["X"], c["y"])
    assert r1["loss"] == pytest.approx(float(c["z"]["train/loss1"]), rel=2e-5)
    assert r1["grad_norm"] == pytest.approx(float(c["z"]["train/norm1"]), rel=1e-4)
    for k, g in c["grad1"].items():
        if ".fn.W_" in k:
            continue
        np.testing.assert_allclose(r1["grads"][k].numpy(), g.numpy(), err_msg=k, **GRAD_TOL)
    r2 = O.train_step(params, bufs, spec, st, c["X"], c["y"])
    assert r2["loss"] == pytest.approx(float(c["z"]["train/loss2"]), rel=5e-5)
    assert r2["grad_norm"] == pytest.approx(float(c["z"]["train/norm2"]), rel=2e-4)
    ref_p, ref_b = split_state(c["sd2"])
    for k, w in ref_p.items():
        # Adam normalises the step to ~lr, so compare with an absolute tolerance of a few % of lr
        atol = 5e-5
        if noise_grad_param(k, spec):
            atol = 4e-3      # true gradient is 0 (bias before BatchNorm): Adam turns rounding noise into +-lr steps
        np.testing.assert_allclose(params[k].numpy(), w.numpy(), rtol=1e-4, atol=atol, err_msg=k)
    for k, w in ref_b.items():
        np.testing.assert_allclose(bufs[k].numpy(), w.numpy(), rtol=1e-4, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", CASES_M2 + CASES_VAR + CASES_VAR_KKBOX)
def test_param_count_small(name):
    c = load_case(name)
    params, _ = split_state(c["sd0"])
    add_dead_params(params, c["spec"])
    assert O.count_parameters(params) == int(c["z"]["param_count"])
    # init_params builds exactly the same key set / shapes as the reference state_dict
    fresh = O.init_params(c["spec"], seed=1)
    assert {k: tuple(v.shape) for k, v in fresh.items()} == {k: tuple(v.shape) for k, v in params.items()}


def test_param_counts_known_answer():
    """'Total number of parameters' of exps/RAT_m2/*/*.log (ml :81, kkbox :80, tmall :82)."""
    want = {"ml": 1337241, "kkbox": 4714649, "tmall": 16970282}
    with open(os.path.join(GOLDEN, "param_counts.json")) as f:
        assert json.load(f) == want            # what the reference itself instantiates to
    for shape, n in want.items():
        spec = O.shape_spec(shape)
        shapes = O.init_params(spec.__class__(**{**spec.__dict__, "features": spec.features}), seed=0) \
            if shape != "tmall" else None
        if shapes is not None:
            assert O.count_parameters(shapes) == n
        else:                                   # avoid materialising 17M randoms twice: count analytically
            D, H, dh, F_ = spec.embedding_dim, spec.num_heads, spec.dim_head, spec.num_fields
            I, M = H * dh, D * spec.scale_dim
            enc = spec.depth * (2 * (2 * D + 3 * I * D + D * I + D) + (M * D + M + D * M + D))
            units = [F_ * D] + list(spec.dnn_hidden_units)
            dnn = sum(a * b + b + 2 * b for a, b in zip(units[:-1], units[1:])) + units[-1] + 1
            tot = spec.total_vocab * (D + 1) + 3 * D + (F_ * D) ** 2 + F_ * D + enc + dnn + D + 1
            assert tot == n


def test_metrics_match_sklearn():
    from sklearn.metrics import log_loss, roc_auc_score
    rng = np.random.default_rng(0)
    y = (rng.random(5000) < 0.3).astype(np.float64)
    p = np.round(rng.random(5000), 2)          # many ties
    assert O.auc(y, p) == pytest.approx(roc_auc_score(y, p), abs=1e-12)
    assert O.logloss(y, p) == pytest.approx(log_loss(y, np.clip(p, 1e-7, 1 - 1e-7)), abs=1e-12)


# ------------------------------------------------------------------------------------------ BM25 top-K retrieval
BM25_CASES = ["plain", "chunked", "selfcheck", "exm_small", "exm_only", "sparse", "kkbox_like"]


def load_bm25_case(name):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"bm25_{name}.npz"))
    qbs = int(z["qbs"])
    return dict(db=z["db"], qry=z["qry"], exm=[int(c) for c in z["exm"]] or None, qbs=None if qbs < 0 else qbs,
                topK=int(z["topK"]), values=z["values"], indices=z["indices"], lens=z["lens"])


def check_bm25_against_reference(c, values, indices, lens):
    """values / lens bit for bit; indices: the reference's torch.topk orders equal scores arbitrarily, so every index
    (ours and the reference's) must be a db row that scores exactly the value reported for it, without repeats, and no
    row with a strictly larger score than the smallest kept one may be missing."""
    from oracle import bm25_oracle as B
    assert np.array_equal(values, c["values"]), "values differ from the reference"
    assert np.array_equal(lens, c["lens"])
    S = B.all_scores(c["db"], c["qry"], c["exm"], c["qbs"], c["topK"])
    for who, ind in (("ours", indices), ("reference", c["indices"])):
        for b in range(len(c["qry"])):
            n = int(c["lens"][b])
            row = ind[b]
            assert (row[:n] >= 0).all() and (row[n:] == -1).all(), who
            assert len(set(row[:n].tolist())) == n, who
            assert np.array_equal(S[b, row[:n]], c["values"][b, :n]), who
            assert (c["values"][b, n:] == 0).all()
            if n:
                assert int((S[b] > c["values"][b, n - 1]).sum()) <= n - 1 or n == c["topK"] and \
                    int((S[b] > c["values"][b, n - 1]).sum()) < n, who
            if n < c["topK"]:
                assert int((S[b] > 0).sum()) == n, who


@pytest.mark.parametrize("name", BM25_CASES)
def test_bm25_oracle_matches_reference_fixtures(name):
    """oracle/bm25_oracle.py against vectors the reference's BM25_topk_retrieval_v4 produced (tests/golden/make_golden_bm25.py):
    plain / chunked / the reference's own self-check configuration (exact-match columns [0, 4]) / small exact-match groups
    (unit scores) / exact matching only (last-K truncation) / unseen query values (-1 padding, integer-IDF quirk) / 13 columns."""
    from oracle import bm25_oracle as B
    c = load_bm25_case(name)
    v, i, n = B.bm25_topk(c["db"], c["qry"], c["exm"], c["qbs"], c["topK"])
    check_bm25_against_reference(c, v, i, n)
    if name == "selfcheck":     # the reference's own assertion (data_utils.py:1318-1324)
        for b in range(len(c["qry"])):
            got = c["db"][i[b]][:, c["exm"]]
            assert int((got == c["qry"][b][c["exm"]]).all(-1).sum()) >= n[b]
