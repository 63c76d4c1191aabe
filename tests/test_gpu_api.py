"""-m gpu tests of the drop-in FuxiCTR API (the calls run_expid.py makes) on the B200 engine."""
import os

import numpy as np
import pytest
import torch

from oracle import rat_oracle as O
from tests.helpers import CASES_M2, CASES_VAR, CASES_VAR_KKBOX, load_case, split_state

pytestmark = pytest.mark.gpu


def _model_from_case(c, tmp_path, **over):
    from fuxictr.features import FeatureMap
    from fuxictr.pytorch import models
    meta = c["meta"]
    fm = FeatureMap(c["name"], str(tmp_path))
    for n, t, v, ml in meta["feats"]:
        s = {"source": "", "type": t, "vocab_size": v}
        if t == "sequence":
            s.update({"padding_idx": v - 1, "max_len": ml, "encoder": "MaskedSumPooling"})
        fm.feature_specs[n] = s
    fm.num_fields = len(meta["feats"])
    fm.set_feature_index()
    kw = dict(model_id=c["name"], gpu=0, task="binary_classification", learning_rate=1e-3, dnn_activations="relu",
              net_dropout=0, emb_dropout=0.0, dropout=0.0, optimizer="adam", loss="binary_crossentropy",
              model_root=str(tmp_path / "exps"), metrics=["AUC", "logloss"], verbose=0, retrieval_augmented=True,
              retrieval_configs={"topK": meta["K"], "label_wise": False})
    kw.update(meta["hp"])
    kw.update(over)
    os.makedirs(os.path.join(kw["model_root"], c["name"]), exist_ok=True)
    return getattr(models, meta["model"])(fm, **kw), fm


@pytest.mark.parametrize("name", CASES_M2 + CASES_VAR + CASES_VAR_KKBOX)
def test_state_dict_keys_and_forward_match_reference(name, tmp_path):
    """state_dict key set/shapes == the reference's (checkpoint compatibility), and after load_state_dict of the
    reference weights, forward() returns the reference's y_pred (fixture), as [B,1] tensors like RAT_m2.py:151."""
    from rat_native.engine import set_precision
    set_precision("fp32")           # strict parity anchor; the default TF32 mode is covered in test_gpu_backward.py
    c = load_case(name)
    model, fm = _model_from_case(c, tmp_path)
    sd = model.state_dict()
    ref_keys = {k: tuple(v.shape) for k, v in c["sd0"].items()}
    got_keys = {k: tuple(v.shape) for k, v in sd.items() if not k.startswith("query_proj")}
    assert got_keys == ref_keys
    full = dict(c["sd0"])
    FD = fm.num_fields * c["spec"].embedding_dim
    full["query_proj.weight"], full["query_proj.bias"] = torch.zeros(FD, FD), torch.zeros(FD)
    model.load_state_dict(full)
    assert model.count_parameters() == int(c["z"]["param_count"])
    model.eval()
    batch = (c["X"], c["y"], torch.zeros(c["X"].shape[0], c["meta"]["K"], dtype=torch.float64),
             torch.zeros(c["X"].shape[0], dtype=torch.int64))
    rd = model.forward(batch)
    assert rd["y_pred"].shape == (c["X"].shape[0], 1) and rd["y_true"].shape == (c["X"].shape[0], 1)
    np.testing.assert_allclose(rd["y_pred"].cpu().numpy(), c["z"]["eval/y_pred"], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(rd["y_true"].cpu().numpy(), c["z"]["eval/y_true"])
    # two train_step calls == two reference optimisation steps
    model.train()
    l1 = float(model.train_step(batch))
    l2 = float(model.train_step(batch))
    assert l1 == pytest.approx(float(c["z"]["train/loss1"]), rel=1e-4)
    assert l2 == pytest.approx(float(c["z"]["train/loss2"]), rel=2e-4)
    # checkpoint round trip through save_weights / load_weights
    model.save_weights(model.checkpoint)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model.train_step(batch)
    model.load_weights(model.checkpoint)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    set_precision("fp16")     # back to the library default


def _write_dataset(tmp_path, fm, n_train=1500, n_valid=400, K=5, seed=0):
    from rat_native import shapes
    d = tmp_path / "data" / fm.dataset_id
    os.makedirs(d, exist_ok=True)
    train = shapes.synthetic_array(fm.feature_specs, n_train, seed)
    valid = shapes.synthetic_array(fm.feature_specs, n_valid, seed + 1)
    # learnable labels: depend on the first feature's id parity
    train[:, -1] = (train[:, 0] % 2 == 0).astype(np.float64)
    valid[:, -1] = (valid[:, 0] % 2 == 0).astype(np.float64)
    np.savez(d / "train.npz", data=train)
    np.savez(d / "valid.npz", data=valid)
    np.savez(d / "test.npz", data=valid)
    for nm, q in (("train", n_train), ("valid", n_valid), ("test", n_valid)):
        idx = shapes.synthetic_neighbours(q, n_train, K, seed + 3, missing=0.1)
        np.savez(d / f"retrieval_{K}_{nm}.npz", indices=idx, values=np.random.default_rng(1).random((q, K)),
                 lens=(idx >= 0).sum(1))
    fm.save(str(d / "feature_map.json"))
    return d


@pytest.mark.parametrize("device_resident", [False, True])
def test_run_expid_call_sequence(tmp_path, device_resident):
    """the exact call sequence of run_expid.py:46-102 (h5 branch): FeatureMap.load -> h5_generator(train/test) ->
    model_class(feature_map, **params) -> count_parameters -> fit_generator -> load_weights -> evaluate_generator."""
    from fuxictr import datasets
    from fuxictr.features import FeatureMap
    from fuxictr.pytorch import models
    from fuxictr.pytorch.torch_utils import seed_everything
    from rat_native import shapes
    fm0 = shapes.make_feature_map("ml", vocab_scale=0.01, data_dir=str(tmp_path))
    d = _write_dataset(tmp_path, fm0)
    params = shapes.model_params("ml", K=5, gpu=0, model_root=str(tmp_path / "exps"), data_root=str(tmp_path / "data"),
                                 data_format="h5", batch_size=256, epochs=2, shuffle=True, num_workers=0,
                                 dnn_hidden_units=[32, 16], embedding_regularizer=1e-6, device_resident=device_resident,
                                 train_data=str(d / "train.h5"), valid_data=str(d / "valid.h5"),
                                 test_data=str(d / "test.h5"), version="pytorch")
    seed_everything(seed=params["seed"])
    feature_map = FeatureMap(params["dataset_id"], str(d), params["version"])
    feature_map.load(str(d / "feature_map.json"))
    train_gen, valid_gen = datasets.h5_generator(feature_map, stage="train", **params)
    test_gen = datasets.h5_generator(feature_map, stage="test", **params)
    model = getattr(models, params["model"])(feature_map, **params)
    model.count_parameters()
    before = model.evaluate_generator(valid_gen)
    model.fit_generator(train_gen, validation_data=valid_gen, **params)
    assert os.path.exists(model.checkpoint)
    model.load_weights(model.checkpoint)
    valid_result = model.evaluate_generator(valid_gen)
    test_result = model.evaluate_generator(test_gen)
    assert set(valid_result) == {"AUC", "logloss"}
    assert valid_result["AUC"] > max(0.8, before["AUC"]), (before, valid_result)     # it learned the parity rule
    assert test_result["AUC"] == pytest.approx(valid_result["AUC"], abs=1e-9)        # same file
    preds = model.predict_generator(test_gen)
    assert preds.shape == (400,) and preds.dtype == np.float64


def test_host_and_device_generators_agree(tmp_path):
    """wire-format DataGenerator (reference semantics, host numpy assembly) and the HBM-resident generator feed the
    model identical batches: identical predictions, bit for bit."""
    from fuxictr import datasets
    from fuxictr.pytorch import models
    from rat_native import shapes
    fm = shapes.make_feature_map("kkbox", vocab_scale=0.01, data_dir=str(tmp_path))
    d = _write_dataset(tmp_path, fm, n_train=600, n_valid=300)
    base = shapes.model_params("kkbox", K=5, gpu=0, model_root=str(tmp_path / "exps"), batch_size=128, num_workers=0,
                               dnn_hidden_units=[32], train_data=str(d / "train.h5"), valid_data=str(d / "valid.h5"),
                               test_data=str(d / "test.h5"))
    os.makedirs(os.path.join(base["model_root"], fm.dataset_id), exist_ok=True)
    model = models.RAT_m2(fm, **base)
    with torch.no_grad():
        for k, v in model._engine.p.items():
            if "embedding_layer.embedding_layer" in k:
                v.mul_(3000.0)
    preds = []
    for dr in (False, True):
        p = dict(base, device_resident=dr)
        gen = datasets.h5_generator(fm, stage="test", **p)
        preds.append(model.predict_generator(gen))
    assert np.array_equal(preds[0], preds[1])


def test_metrics_agree_with_oracle(tmp_path):
    from fuxictr.metrics import evaluate_metrics
    rng = np.random.default_rng(0)
    y = (rng.random(3000) < 0.4).astype(np.float64)
    p = np.clip(rng.random(3000), 0, 1)
    r = evaluate_metrics(y, p, ["AUC", "logloss"])
    assert r["AUC"] == pytest.approx(O.auc(y, p), abs=1e-12)
    assert r["logloss"] == pytest.approx(O.logloss(y, p), abs=1e-12)


@pytest.mark.parametrize("n,ties", [(1, False), (1000, False), (5000, True), (300_000, True)])
def test_device_auc_logloss_match_sklearn(n, ties):
    """rat_auc_logloss vs sklearn (the reference's evaluate_metrics): exact tie-aware AUC and float64 logloss; 1e-9."""
    from sklearn.metrics import log_loss, roc_auc_score
    from rat_native.engine import EngineSpec, FeatureSpec, RatEngine
    eng = RatEngine(EngineSpec(features=[FeatureSpec("a", "categorical", 8)]), "cuda:0")
    g = torch.Generator().manual_seed(n)
    p = torch.rand(n, generator=g)
    if ties:
        p = (p * 50).round() / 50                          # heavy ties incl. exact 0.0 and 1.0
    y = (torch.rand(n, generator=g) < 0.3 + 0.4 * p).float()
    if n == 1:
        auc, ll = eng.auc_logloss(p.cuda(), y.cuda())
        assert np.isnan(auc)                               # one class only (sklearn raises)
        return
    auc, ll = eng.auc_logloss(p.cuda(), y.cuda())
    pn, yn = p.double().numpy(), y.double().numpy()
    assert auc == pytest.approx(roc_auc_score(yn, pn), abs=1e-9)
    assert ll == pytest.approx(log_loss(yn, np.clip(pn, 1e-7, 1 - 1e-7)), abs=1e-9)
    auc2, ll2 = eng.auc_logloss(p.cuda(), y.cuda())
    assert auc2 == auc and ll2 == ll                       # deterministic


def test_dropout_training_is_statistically_sane(tmp_path):
    """emb_dropout / net_dropout > 0 (kkbox, tmall configs): the reference philox stream cannot be reproduced
    (SURVEY H4); check that training with the device-side masks still decreases the loss and stays finite."""
    from fuxictr.pytorch import models
    from rat_native import shapes
    fm = shapes.make_feature_map("tmall", vocab_scale=0.001, data_dir=str(tmp_path))
    params = shapes.model_params("tmall", K=5, gpu=0, model_root=str(tmp_path / "exps"), dnn_hidden_units=[64, 32],
                                 embedding_regularizer=1e-6)
    assert params["emb_dropout"] == 0.1 and params["net_dropout"] == 0.08
    os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
    model = models.RAT_m2(fm, **params)
    pool = shapes.synthetic_array(fm.feature_specs, 3000, seed=3)
    pool[:, -1] = (pool[:, 2] % 2 == 0).astype(np.float64)
    nbr = shapes.synthetic_neighbours(3000, 3000, 5, seed=3)
    model.train()
    losses = []
    for i in range(30):
        rows = np.arange(i * 100, (i + 1) * 100) % 3000
        batch = tuple(torch.from_numpy(t) for t in shapes.host_wire_batch(pool, pool, nbr, rows))
        losses.append(float(model.train_step(batch)))
    assert np.isfinite(losses).all()
    assert np.mean(losses[-5:]) < np.mean(losses[:5])


def test_csv_branch_end_to_end(tmp_path):
    """run_expid.py's csv branch (:53-72) on synthetic kkbox-like csv files: datasets.kkbox.FeatureEncoder -> build_dataset
    (feature_map.json, train / valid / test / retrieval_pool id blocks) -> h5_generator (BM25 pre-retrieval on the GPU, no cached
    retrieval files) -> RAT_m2(feature_map) -> fit_generator -> evaluate_generator."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_encoder as G
    from fuxictr import datasets
    from fuxictr.pytorch import models
    from fuxictr.pytorch.torch_utils import seed_everything
    from rat_native import shapes
    case = "kkbox_pool_ratio"
    module, cols, args = G.write_case_csvs(case, str(tmp_path / "csv"))
    used = ["msno", "song_id", "city", "isrc", "bd"]
    params = shapes.model_params("kkbox", K=3, gpu=0, model_root=str(tmp_path / "exps"), data_root=str(tmp_path / "data"),
                                 data_format="csv", batch_size=64, epochs=1, shuffle=True, num_workers=0,
                                 dnn_hidden_units=[32, 16], embedding_regularizer=1e-6, version="pytorch",
                                 feature_cols=cols, label_col=dict(G.LABEL), **args)
    params["dataset_id"] = case
    params["retrieval_configs"] = dict(params["retrieval_configs"], used_cols=used, exact_match_cols=[], label_wise=False,
                                       pre_retrieval=True, enable_clean=False, qry_batch_size=128, db_chunk_size=1000,
                                       device="cuda:0", topK=3)
    seed_everything(seed=params["seed"])
    enc = getattr(datasets, module).FeatureEncoder(**params)
    assert not os.path.exists(enc.json_file)
    datasets.build_dataset(enc, **params)
    data_dir = os.path.join(params["data_root"], case)
    params["train_data"] = os.path.join(data_dir, "train*.h5")
    params["valid_data"] = os.path.join(data_dir, "valid*.h5")
    params["test_data"] = os.path.join(data_dir, "test*.h5")
    params["retrieval_configs"]["retrieval_pool_data"] = os.path.join(data_dir, "retrieval_pool.h5")
    feature_map = enc.feature_map
    assert feature_map.num_fields == 7 and feature_map.input_length == 3 + 3 + feature_map.feature_specs["artist_name"]["max_len"] + 2
    train_gen, valid_gen = datasets.h5_generator(feature_map, stage="train", **params)
    test_gen = datasets.h5_generator(feature_map, stage="test", **params)
    assert train_gen.num_samples == 480 and valid_gen.num_samples == 150 and test_gen.num_samples == 150
    model = getattr(models, params["model"])(feature_map, **params)
    model.fit_generator(train_gen, validation_data=valid_gen, **params)
    res = model.evaluate_generator(test_gen)
    assert set(res) == {"AUC", "logloss"} and np.isfinite(res["logloss"]) and 0.0 <= res["AUC"] <= 1.0
    model._engine.check_errors()
