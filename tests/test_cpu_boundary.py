"""CPU-only tests of the boundary: the C-ABI library loads and exports every symbol include/rat_b200.h declares,
the product refuses to run without a GPU (no CPU fallback), the drop-in API mirrors the reference's signatures,
config / FeatureMap / data generators behave like the reference, and the data-parallel math holds (gloo, 2 ranks)."""
import inspect
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import rat_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    import ctypes
    import rat_native
    protos = rat_native.parse_header()
    assert len(protos) >= 35
    cdll = ctypes.CDLL(rat_native.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/rat_b200.h but not exported"
    L = rat_native.lib()
    assert L.fn["rat_abi_version"]() == 1
    # every entry point documents the reference code it replaces
    src = open(rat_native.HEADER).read()
    for token in ("data_generator.py:66-78", "embedding.py", "RAT_m2.py", "base_model.py:224", "torch_utils.py:41-49"):
        assert token in src


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import rat_native
    from rat_native import shapes
    from fuxictr.pytorch import models
    with pytest.raises(rat_native.RatError):
        rat_native.require_device()
    fm = shapes.make_feature_map("ml", vocab_scale=0.01)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        models.RAT_m2(fm, **shapes.model_params("ml", gpu=-1))
    with pytest.raises(RuntimeError):
        models.RAT_m2(fm, **shapes.model_params("ml", gpu=0))
    # a compute entry point without a device fails with an error code, never silently
    rc = rat_native.lib().fn["rat_layernorm_fwd"](None, None, None, None, 4, 4, None)
    assert rc != 0 and rat_native.lib().last_error()


def test_product_code_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "www24-rat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if "rat_oracle" in txt or "from oracle" in txt or "import oracle" in txt:
                    bad.append(os.path.join(base, f))
    assert not bad, bad


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_api_signatures_mirror_reference():
    """same constructor / method names and argument names as the reference classes (read with ast, not imported)."""
    import ast

    def sigs(path, cls):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef) and node.name == cls:
                return {f.name: [a.arg for a in f.args.args] + ([f.args.kwarg.arg] if f.args.kwarg else [])
                        for f in node.body if isinstance(f, ast.FunctionDef)}
        raise KeyError(cls)

    from fuxictr.pytorch import models
    ref_base = sigs(f"{REF}/fuxictr/pytorch/models/base_model.py", "BaseModel")
    for meth, args in ref_base.items():
        assert hasattr(models.BaseModel, meth), f"BaseModel.{meth} missing"
        got = list(inspect.signature(getattr(models.BaseModel, meth)).parameters)
        assert got == args, f"BaseModel.{meth}: {got} != {args}"
    for m in ("RAT_m0", "RAT_m1", "RAT_m2", "RAT_m3"):
        ref_init = sigs(f"{REF}/fuxictr/pytorch/models/{m}.py", m)["__init__"]
        got = list(inspect.signature(getattr(models, m).__init__).parameters)
        assert got == ref_init, f"{m}.__init__: {got} != {ref_init}"
    import fuxictr
    assert fuxictr.__version__.startswith("1.2")           # run_expid.py:15
    from fuxictr import datasets
    from fuxictr.features import FeatureMap
    ref_fm = sigs(f"{REF}/fuxictr/features.py", "FeatureMap")
    for meth, args in ref_fm.items():
        assert list(inspect.signature(getattr(FeatureMap, meth)).parameters) == args
    assert callable(datasets.h5_generator) and callable(datasets.build_dataset)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_load_config_on_the_shipped_reference_configs():
    from fuxictr.utils import load_config
    for d, expid, ds in [("kkbox_x1", "RAT_m2_kkbox_x1_10fold_retrieval", "kkbox_x1_10fold_retrieval"),
                         ("movielenslatest_x1", "RAT_m2_movielenslatest_x1_10fold_retrieval",
                          "movielenslatest_x1_10fold_retrieval"),
                         ("tmall_x1_002", "RAT_m2_tmall_x1_002_retrieval", "tmall_x1_002_retrieval")]:
        p = load_config(f"{REF}/configs/RAT_m2/{d}", expid)
        assert p["model"] == "RAT_m2" and p["dataset_id"] == ds and p["model_id"] == expid
        assert p["retrieval_configs"]["topK"] == 5 and p["batch_size"] == 4096


def test_feature_map_json_roundtrip_and_shapes(tmp_path):
    from fuxictr.features import FeatureMap
    from rat_native import shapes
    from rat_native.engine import EngineSpec, FeatureSpec, param_shapes
    for shape, want in (("ml", 1337241), ("kkbox", 4714649), ("tmall", 16970282)):
        fm = shapes.make_feature_map(shape, data_dir=str(tmp_path))
        p = str(tmp_path / f"{shape}.json")
        fm.save(p)
        fm2 = FeatureMap(fm.dataset_id, str(tmp_path))
        fm2.load(p)
        assert json.dumps(fm2.feature_specs) == json.dumps(fm.feature_specs)
        assert fm2.input_length == fm.input_length and fm2.num_fields == fm.num_fields
        # product-side parameter inventory reproduces the reference's logged parameter counts
        hp = shapes.SHAPES[shape]["hp"]
        feats = [FeatureSpec(n, s["type"], s["vocab_size"], s.get("max_len", 1), s.get("padding_idx"))
                 for n, s in fm.feature_specs.items()]
        es = EngineSpec(features=feats, embedding_dim=hp["embedding_dim"], num_heads=hp["num_heads"], dim_head=10,
                        scale_dim=hp["scale_dim"], depth=4, dnn_hidden_units=tuple(hp["dnn_hidden_units"]),
                        batch_norm=hp["batch_norm"], use_wide=True, net_dropout=hp["net_dropout"])
        net, emb = param_shapes(es)
        total = sum(int(np.prod(s)) for s in list(net.values()) + list(emb.values()))
        assert total == want
        # and the oracle's synthetic shape agrees with the product's
        ospec = O.shape_spec(shape)
        assert [f.vocab_size for f in ospec.features] == [s["vocab_size"] for s in fm.feature_specs.values()]


def test_host_data_generator_wire_format(tmp_path):
    """DataGenerator over npz mirrors yields the reference wire format, -1 neighbours wrap to the last pool row."""
    from fuxictr import datasets
    from rat_native import shapes
    fm = shapes.make_feature_map("kkbox", vocab_scale=0.01, data_dir=str(tmp_path))
    d = tmp_path / fm.dataset_id
    os.makedirs(d)
    train = shapes.synthetic_array(fm.feature_specs, 300, 0)
    valid = shapes.synthetic_array(fm.feature_specs, 100, 1)
    np.savez(d / "train.npz", data=train)
    np.savez(d / "valid.npz", data=valid)
    K = 5
    for nm, q in (("train", 300), ("valid", 100)):
        idx = shapes.synthetic_neighbours(q, 300, K, 3, missing=0.3)
        np.savez(d / f"retrieval_{K}_{nm}.npz", indices=idx, values=np.ones((q, K)), lens=(idx >= 0).sum(1))
    rc = {"topK": K, "label_wise": False, "pre_retrieval": True, "split_type": "10-fold",
          "used_cols": ["msno", "song_id"]}
    tr, va = datasets.h5_generator(fm, stage="train", train_data=str(d / "train.h5"), valid_data=str(d / "valid.h5"),
                                   batch_size=64, shuffle=False, retrieval_configs=rc, retrieval_augmented=True,
                                   num_workers=0, device_resident=False, gpu=-1)
    assert len(tr) == 5 and len(va) == 2 and tr.num_samples == 300
    X, y, vals, lens = next(iter(va))
    assert X.dtype == torch.float64 and tuple(X.shape) == (64, K + 1, fm.input_length) and tuple(y.shape) == (64, K + 1)
    idx = np.load(d / f"retrieval_{K}_valid.npz")["indices"]
    wantX, wanty = O.assemble_batch(valid, train, idx, np.arange(64))      # valid retrieves from train block 0
    np.testing.assert_array_equal(X.numpy(), wantX)
    np.testing.assert_array_equal(y.numpy(), wanty)
    assert (idx[:64] == -1).any()
    assert rc["used_col_indices"] == [0, 1]


def test_device_generator_sharding_logic_on_cpu():
    """DeviceDataGenerator is device-agnostic torch code: check len(), shuffling and rank sharding on CPU tensors."""
    from fuxictr.pytorch.data_generator import DeviceDataGenerator
    arr = np.concatenate([np.arange(1000)[:, None].repeat(3, 1), np.zeros((1000, 1))], 1).astype(np.float64)
    nbr = np.zeros((1000, 2), np.int64)
    seen = []
    for rank in range(2):
        g = DeviceDataGenerator(arr, arr, nbr, batch_size=128, shuffle=True, device="cpu", seed=7, rank=rank, world=2)
        assert len(g) == 8
        rows = [b.rows for b in g]
        assert all(r.numel() == 64 for r in rows[:-1])
        seen.append(torch.cat(rows))
    both = torch.cat(seen)
    assert both.unique().numel() == 1000                       # the two ranks partition every global batch
    g = DeviceDataGenerator(arr, arr, nbr, batch_size=128, shuffle=False, device="cpu")
    bs = list(g)
    assert [b.row0 for b in bs][:3] == [0, 128, 256] and bs[-1].size == 1000 - 7 * 128


@pytest.mark.parametrize("shape,count", [("ml", 1337241), ("kkbox", 4714649), ("tmall", 16970282)])
def test_flat_parameter_layout_reproduces_the_reference_counts(shape, count):
    """the product's flat parameter buffer (ParamStore, built on CPU here): same parameter count as the reference logs
    (SURVEY 8c known answers), the reference's state_dict key set (query_proj included), 16-byte aligned net segments
    and table regions, the per-field tables contiguous as emb_W / lr_W, and [net | embedding-named] split at
    reg_boundary."""
    sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
    from rat_native import shapes
    from rat_native.engine import EMB, LRP, EngineSpec, FeatureSpec, ParamStore
    fm = shapes.make_feature_map(shape)
    hp = shapes.SHAPES[shape]["hp"]
    feats = [FeatureSpec(n, sp["type"], sp["vocab_size"], sp.get("max_len", 1), sp.get("padding_idx")) for n, sp in
             fm.feature_specs.items()]
    spec = EngineSpec(features=feats, model="RAT_m2", embedding_dim=hp["embedding_dim"], num_heads=hp["num_heads"],
                      dim_head=10, scale_dim=hp["scale_dim"], depth=4, dnn_hidden_units=tuple(hp["dnn_hidden_units"]),
                      batch_norm=hp["batch_norm"], use_wide=True, emb_dropout=hp["emb_dropout"],
                      net_dropout=hp["net_dropout"])           # net_dropout shifts the nn.Sequential indices (deep.py:126-137)
    st = ParamStore(spec, "cpu")
    assert st.numel_params() == count
    ospec = O.shape_spec(shape)
    want = O.init_params(ospec, 0)
    assert set(st.offsets) == set(want), set(st.offsets) ^ set(want)
    for k, (off, shp) in st.offsets.items():
        assert tuple(want[k].shape) == tuple(shp), k
        if off < st.emb_off:                          # net parameters and the label table start on 16-byte boundaries
            assert off % 4 == 0, k
    assert st.emb_off % 4 == 0 and st.lr_off % 4 == 0 and st.net_end % 4 == 0
    V, D = spec.V, spec.embedding_dim
    assert st.emb_W.shape == (V, D) and st.lr_W.shape == (V,)
    row = 0
    for f in feats:                                   # field tables are consecutive row ranges of emb_W / lr_W
        assert st.offsets[EMB + f.name + ".weight"][0] == st.emb_off + row * D
        assert st.offsets[LRP + f.name + ".weight"][0] == st.lr_off + row
        row += f.vocab_size
    assert row == V == shapes.SHAPES[shape]["total_vocab"]
    emb_named = [k for k in st.offsets if "embedding_layer" in k]
    assert all(st.offsets[k][0] >= st.net_end for k in emb_named)
    assert all(st.offsets[k][0] < st.net_end for k in st.offsets if k not in emb_named)


def test_bench_algorithmic_bytes_match_the_survey():
    """bench.py's gather bytes per sample = SURVEY 8(d)'s figures (1,810 / 30,282 / 4,350 B) + the X_emb write."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    for (K, L, F, D), survey in (((5, 3, 3, 10), 1810), ((5, 17, 13, 40), 30282), ((5, 8, 8, 10), 4350)):
        assert bench.gather_bytes_per_sample(K, L, F, D) == survey + F * D * 4
    assert bench.encoder_flops_per_sample(5, 13, 40, 8, 10, 2) == pytest.approx(23.7e6, rel=0.01)


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    spec = O.shape_spec("kkbox", vocab_scale=0.003, dnn_hidden_units=(32, 16), depth=1, emb_dropout=0.0)
    params = O.init_params(spec, 0)
    for k, v in params.items():
        if "embedding_layer.embedding_layer" in k:
            v.mul_(3000.0)
    pool = O.synthetic_pool(spec, 500, seed=1)
    nbr = O.synthetic_neighbours(32, 500, 5, seed=1)
    X, y = O.assemble_batch(pool[:32], pool, nbr, np.arange(32))
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    lo, hi = rank * 16, (rank + 1) * 16

    # data-parallel recipe used by the engine: local sum-of-BCE / B_global, BN batch statistics from all-reduced
    # raw sums (SyncBN-equivalent), gradients all-reduced with SUM, regulariser added once afterwards.
    class SyncBN(torch.autograd.Function):
        @staticmethod
        def forward(ctx, h, gamma, beta):
            n = torch.tensor(float(h.shape[0]))
            s = torch.cat([h.sum(0), (h * h).sum(0), n.view(1)])
            dist.all_reduce(s)
            C = h.shape[1]
            mean, var = s[:C] / s[-1], s[C:2 * C] / s[-1] - (s[:C] / s[-1]) ** 2
            rstd = 1.0 / torch.sqrt(var + 1e-5)
            xh = (h - mean) * rstd
            ctx.save_for_backward(xh, gamma, rstd)
            ctx.count = float(s[-1])
            return xh * gamma + beta
        @staticmethod
        def backward(ctx, dy):
            xh, gamma, rstd = ctx.saved_tensors
            s = torch.cat([dy.sum(0), (dy * xh).sum(0)])
            dist.all_reduce(s)
            C = dy.shape[1]
            db, dg = s[:C], s[C:]
            dx = gamma * rstd * (dy - db / ctx.count - xh * dg / ctx.count)
            return dx, (dy * xh).sum(0), dy.sum(0)

    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ids = X[lo:hi].long()
    block = O.feature_block(leaves, spec, ids, y[lo:hi])
    pooled = O.encode(leaves, spec, block)
    logit = pooled @ leaves["fc.weight"].t() + leaves["fc.bias"]
    h = block[:, 0, 1:, :].flatten(1)
    layers, final = O.dnn_layout(spec)
    for lin, bn in layers:
        h = h @ leaves[f"dnn.dnn.{lin}.weight"].t() + leaves[f"dnn.dnn.{lin}.bias"]
        h = torch.relu(SyncBN.apply(h, leaves[f"dnn.dnn.{bn}.weight"], leaves[f"dnn.dnn.{bn}.bias"]))
    logit = logit + h @ leaves[f"dnn.dnn.{final}.weight"].t() + leaves[f"dnn.dnn.{final}.bias"]
    logit = logit + O.embed_rows(leaves, spec, ids[:, 0:1, :], prefix=O.LR).sum(dim=-2).mean(dim=1)
    yp = torch.sigmoid(logit)
    loss = torch.nn.functional.binary_cross_entropy(yp, y[lo:hi, 0:1].float(), reduction="sum") / 32.0
    loss.backward()
    flat = torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in leaves.values()])
    dist.all_reduce(flat)
    if rank == 0:
        _, _, ref = O.total_loss_and_grads(params, O.init_buffers(spec), O.ModelSpec(**{**spec.__dict__, "embedding_regularizer": 0.0}), X, y)
        want = torch.cat([(g if g is not None else torch.zeros_like(params[k])).reshape(-1) for k, g in ref.items()])
        q.put(float((flat - want).abs().max() / want.abs().max()))
    dist.destroy_process_group()


def test_data_parallel_recipe_equals_single_device_gloo():
    """2 gloo ranks, each with half of the batch: (local BCE sum / B_global) + SyncBN raw-sum all-reduce + SUM
    all-reduce of gradients reproduces the single-process gradient at the global batch (what engine.py does on NCCL)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _q
    rel = None
    for _ in range(240):
        try:
            rel = q.get(timeout=1)
            break
        except _q.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rel is not None
    assert rel < 2e-5, rel
