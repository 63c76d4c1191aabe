"""-m gpu parity tests of the forward kernels (called through the C ABI) against the CPU oracle and the
golden fixtures produced by the reference.  Integer / index work and pure gathers are checked BIT-EXACT;
floating-point kernels within the fp32 tolerances written next to each assert."""
import numpy as np
import pytest
import torch

from oracle import rat_oracle as O
from tests.helpers import CASES_M2, load_case, split_state

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def rn():
    import rat_native
    rat_native.require_device()
    return rat_native


# ------------------------------------------------------------------------------------------ K0
@pytest.mark.parametrize("name", CASES_M2)
def test_convert_wire_bit_exact(rn, name):
    c = load_case(name)
    X, y = c["X"].cuda(), c["y"].cuda()
    B, T, L = X.shape
    ids = torch.empty(B, T, L, dtype=torch.int32, device=_dev())
    labels = torch.empty(B, T, dtype=torch.int32, device=_dev())
    yt = torch.empty(B, device=_dev())
    rn.call("rat_convert_wire_f64", X, y, ids, labels, yt, B, T, L, rn.current_stream())
    want_lab = c["y"].long().clone()
    want_lab[:, 0] = 2
    assert torch.equal(ids.cpu().long(), c["X"].long())
    assert torch.equal(labels.cpu().long(), want_lab)
    assert torch.equal(yt.cpu(), c["y"][:, 0].float())


@pytest.mark.parametrize("name", CASES_M2)
def test_assemble_ids_bit_exact_with_wraparound(rn, name):
    """device-side pool[retr_indices] incl. -1 -> last pool row == reference Dataset.__getitem__ output."""
    c = load_case(name)
    z = c["z"]
    darray, pool, idx = z["darray"], z["pool"], z["retr_indices"]
    assert (idx == -1).any()
    B, K = idx.shape
    T, L = K + 1, darray.shape[1] - 1
    d = _dev()
    q_ids = torch.from_numpy(darray[:, :-1].astype(np.int32)).to(d)
    q_lab = torch.from_numpy(darray[:, -1].astype(np.uint8)).to(d)
    p_ids = torch.from_numpy(pool[:, :-1].astype(np.int32)).to(d)
    p_lab = torch.from_numpy(pool[:, -1].astype(np.uint8)).to(d)
    nbr = torch.from_numpy(idx).to(d)
    ids = torch.empty(B, T, L, dtype=torch.int32, device=d)
    labels = torch.empty(B, T, dtype=torch.int32, device=d)
    yt = torch.empty(B, device=d)
    err = torch.zeros(1, dtype=torch.int32, device=d)
    # explicit row list (shuffled batch) and the implicit contiguous form
    rows = torch.arange(B, device=d, dtype=torch.int64)
    for r in (rows, None):
        rn.call("rat_assemble_ids", q_ids, q_lab, r, 0, p_ids, p_lab, nbr, pool.shape[0], ids, labels, yt, B, T, L,
                err, rn.current_stream())
        want_lab = c["y"].long().clone()
        want_lab[:, 0] = 2
        assert torch.equal(ids.cpu().long(), c["X"].long())
        assert torch.equal(labels.cpu().long(), want_lab)
        assert torch.equal(yt.cpu(), c["y"][:, 0].float())
        assert int(err.item()) == 0
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    rn.call("rat_assemble_ids", q_ids, q_lab, perm.to(d), 0, p_ids, p_lab, nbr, pool.shape[0], ids, labels, yt, B, T,
            L, err, rn.current_stream())
    assert torch.equal(ids.cpu().long(), c["X"].long()[perm])
    # out-of-range neighbour index is flagged, not dereferenced
    bad = nbr.clone()
    bad[0, 0] = pool.shape[0] + 5
    rn.call("rat_assemble_ids", q_ids, q_lab, None, 0, p_ids, p_lab, bad, pool.shape[0], ids, labels, yt, B, T, L,
            err, rn.current_stream())
    assert int(err.item()) & 2


# ------------------------------------------------------------------------------------------ K1
def _gather_case(rn, spec, params, X, y, drop_p=0.0):
    from tests.gpu_util import make_engine
    eng = make_engine(spec, params)
    ws = eng.load_wire(X.cuda(), y.cuda(), training=False)
    B, T, L = X.shape
    F, D = spec.num_fields, spec.embedding_dim
    block = torch.empty(B, T, F + 1, D, device=_dev())
    rn.call("rat_gather_fwd", eng.store.emb_W, eng.store.lr_W, eng.p["label_embedding_layer.weight"], ws["ids"],
            ws["labels"], eng.col_off, eng.col_vocab, eng.field_col0, eng.field_width, block, ws["x_emb"],
            ws["lr_out"], B, T, L, F, D, float(drop_p), 1234, 7, eng.err_flag, rn.current_stream())
    return eng, ws, block


@pytest.mark.parametrize("name", CASES_M2)
def test_gather_bit_exact_golden(rn, name):
    c = load_case(name)
    params, _ = split_state(c["sd0"])
    spec = c["spec"]
    eng, ws, block = _gather_case(rn, spec, params, c["X"], c["y"])
    want = O.feature_block(params, spec, c["X"].long(), c["y"])
    assert torch.equal(block.cpu(), want), "gather is pure data movement (+ left-to-right 3-term sum-pool): bit-exact"
    assert torch.equal(ws["x_emb"].cpu(), want[:, 0, 1:, :].reshape(want.shape[0], -1))
    lr = O.embed_rows(params, spec, c["X"].long()[:, 0:1, :], prefix=O.LR).sum(dim=-2).mean(dim=1)[:, 0]
    # the LR logit is an F-term fp32 reduction (torch's vectorised sum order is not left-to-right): 1e-6, not bitwise
    torch.testing.assert_close(ws["lr_out"].cpu(), lr, rtol=1e-6, atol=1e-7)
    assert int(eng.err_flag.item()) == 0


@pytest.mark.parametrize("shape,B,K", [("ml", 300, 5), ("kkbox", 257, 5), ("tmall", 128, 8), ("kkbox", 16, 64)])
def test_gather_bit_exact_synthetic(rn, shape, B, K):
    spec = O.shape_spec(shape, vocab_scale=0.02)
    from tests.gpu_util import rand_params_nontrivial
    params = rand_params_nontrivial(spec, seed=3)
    pool = O.synthetic_pool(spec, 2000, seed=5)
    nbr = O.synthetic_neighbours(B, 2000, K, seed=5, missing=0.2)
    X, y = O.assemble_batch(pool[:B], pool, nbr, np.arange(B))
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    eng, ws, block = _gather_case(rn, spec, params, X, y)
    want = O.feature_block(params, spec, X.long(), y)
    assert torch.equal(block.cpu(), want)
    assert torch.equal(ws["x_emb"].cpu(), want[:, 0, 1:, :].reshape(B, -1))
    lr = O.embed_rows(params, spec, X.long()[:, 0:1, :], prefix=O.LR).sum(dim=-2).mean(dim=1)[:, 0]
    torch.testing.assert_close(ws["lr_out"].cpu(), lr, rtol=1e-6, atol=1e-7)


def test_gather_flags_bad_ids_and_dropout_stats(rn):
    spec = O.shape_spec("kkbox", vocab_scale=0.02)
    from tests.gpu_util import rand_params_nontrivial
    params = rand_params_nontrivial(spec, seed=3)
    pool = O.synthetic_pool(spec, 500, seed=5)
    nbr = O.synthetic_neighbours(64, 500, 5, seed=5)
    X, y = O.assemble_batch(pool[:64], pool, nbr, np.arange(64))
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    Xbad = X.clone()
    Xbad[3, 2, 1] = 10 ** 7
    eng, ws, block = _gather_case(rn, spec, params, Xbad, y)
    assert int(eng.err_flag.item()) & 1
    # dropout: kept elements are scaled by 1/(1-p), dropped are 0, keep-rate ~ 1-p, x_emb untouched
    eng, ws, clean = _gather_case(rn, spec, params, X, y, 0.0)
    x_emb_clean = ws["x_emb"].clone()
    eng, ws, dropped = _gather_case(rn, spec, params, X, y, 0.25)
    nz = clean != 0
    kept = (dropped != 0) & nz
    rate = kept.sum().item() / nz.sum().item()
    assert abs(rate - 0.75) < 0.01, rate
    torch.testing.assert_close(dropped[kept], clean[kept] / 0.75, rtol=1e-6, atol=0)
    assert torch.equal(ws["x_emb"], x_emb_clean)


# ------------------------------------------------------------------------------------------ K2
ATTN_SHAPES = [  # B, T, N, D, H, dh
    (5, 6, 4, 10, 2, 10),      # movielens
    (9, 6, 14, 40, 8, 10),     # kkbox
    (7, 6, 9, 10, 32, 10),     # tmall
    (3, 65, 14, 40, 8, 10),    # K=64 sweep: cross sequences of 65 > 32 lanes
    (2, 1, 84, 40, 8, 10),     # RAT_m0: one flat sequence of T*N tokens
    (4, 3, 5, 20, 2, 20),      # RAT_m3 head width 2*dim_head
    (300, 2, 3, 16, 4, 8),     # many tiles / ragged last tile
    (420, 6, 14, 40, 8, 10),   # kkbox shape, several tasks per warp (double-buffered bulk row staging)
]


@pytest.mark.parametrize("B,T,N,D,H,dh", ATTN_SHAPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_attn_fwd_matches_oracle(rn, precision, B, T, N, D, H, dh, mode):
    """PreNorm + MHA + out-proj + residual. fp32 vs fp32-CPU: rtol 2e-4, atol 2e-5 (online softmax, fma order)."""
    from tests.gpu_util import assert_close, ptol
    g = torch.Generator().manual_seed(B * 1000 + T * 100 + N * 10 + mode)
    I = H * dh
    x = torch.randn(B, T, N, D, generator=g)
    lnw, lnb = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    wqkv = torch.randn(3 * I, D, generator=g) * (2.0 / (D + 3 * I)) ** 0.5 * 3
    wo = torch.randn(D, I, generator=g) * (2.0 / (D + I)) ** 0.5
    bo = 0.1 * torch.randn(D, generator=g)
    scale = 10 ** -0.5
    z = x.reshape(B * T, N, D) if mode == 0 else x.transpose(1, 2).reshape(B * N, T, D)
    zn = torch.nn.functional.layer_norm(z, (D,), lnw, lnb, 1e-5)
    o = O.mha(zn, wqkv[:I], wqkv[I:2 * I], wqkv[2 * I:], H, scale, wo, bo)
    want = z + o
    want = want.reshape(B, T, N, D) if mode == 0 else want.reshape(B, N, T, D).transpose(1, 2)
    d = _dev()
    xd, out = x.to(d), torch.empty(B, T, N, D, device=d)
    wq = wqkv.to(d)
    rn.call("rat_attn_fwd", xd, xd, out, lnw.to(d), lnb.to(d), wq[:I], wq[I:2 * I], wq[2 * I:], wo.to(d), bo.to(d),
            B, T, N, D, H, dh, scale, 1.0, mode, rn.current_stream())
    assert_close("attn_fwd", out, want, *ptol(2e-4, 2e-5 * float(want.abs().max())))
    # res=None, alpha=0.5 variant (RAT_m3)
    rn.call("rat_attn_fwd", xd, None, out, lnw.to(d), lnb.to(d), wq[:I], wq[I:2 * I], wq[2 * I:], wo.to(d), bo.to(d),
            B, T, N, D, H, dh, scale, 0.5, mode, rn.current_stream())
    o4 = o.reshape(B, T, N, D) if mode == 0 else o.reshape(B, N, T, D).transpose(1, 2)
    assert_close("attn_fwd(alpha=.5,res=None)", out, 0.5 * o4, *ptol(2e-4, 2e-5 * float(want.abs().max())))


@pytest.mark.parametrize("rows,D,M,prenorm", [(120, 10, 40, False), (5000, 40, 80, False), (333, 10, 20, False),
                                              (777, 20, 40, True), (4097, 40, 80, True),
                                              (40011, 40, 80, False)])     # several tiles per persistent CTA, ragged tail
def test_ff_fwd_matches_oracle(rn, precision, rows, D, M, prenorm):
    """x + W2 gelu_erf(W1 [LN]x + b1) + b2 ; rtol 1e-4 atol 1e-5."""
    from tests.gpu_util import assert_close, ptol
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, D, generator=g)
    w1, b1 = torch.randn(M, D, generator=g) * 0.3, 0.1 * torch.randn(M, generator=g)
    w2, b2 = torch.randn(D, M, generator=g) * 0.3, 0.1 * torch.randn(D, generator=g)
    lnw, lnb = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    u = torch.nn.functional.layer_norm(x, (D,), lnw, lnb, 1e-5) if prenorm else x
    want = x + torch.nn.functional.gelu(u @ w1.t() + b1) @ w2.t() + b2
    d = _dev()
    xd, out = x.to(d), torch.empty(rows, D, device=d)
    rn.call("rat_ff_fwd", xd, xd, out, lnw.to(d) if prenorm else None, lnb.to(d) if prenorm else None, w1.to(d),
            b1.to(d), w2.to(d), b2.to(d), rows, D, M, rn.current_stream())
    assert_close("ff_fwd", out, want, *ptol(1e-4, 1e-5 * float(want.abs().max())))


def test_layernorm_fwd(rn):
    from tests.gpu_util import assert_close
    for rows, D in [(1000, 40), (77, 10), (5, 20)]:
        x = torch.randn(rows, D)
        w, b = torch.randn(D), torch.randn(D)
        out = torch.empty(rows, D, device=_dev())
        rn.call("rat_layernorm_fwd", x.cuda(), out, w.cuda(), b.cuda(), rows, D, rn.current_stream())
        assert_close("ln", out, torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5), 1e-5, 1e-5)


# ------------------------------------------------------------------------------------------ K3
@pytest.mark.parametrize("M,N,K", [(4096, 400, 520), (100, 33, 30), (400, 520, 4096), (64, 1, 400), (4096, 1, 400)])
def test_sgemm_all_layouts(rn, M, N, K):
    """fp32 SIMT GEMM (precision mode fp32) vs float64 reference: rtol 1e-5*sqrt(K)."""
    from tests.gpu_util import assert_close
    from rat_native.engine import set_precision
    set_precision("fp32")
    g = torch.Generator().manual_seed(M + N + K)
    d = _dev()
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    want = (A.double() @ Bm.double().t() + bias.double()).float()
    nbytes = int(rn.query("rat_sgemm_workspace_bytes", M, N, K))
    ws = torch.empty(max(nbytes // 4, 4), device=d)
    C = torch.empty(M, N, device=d)
    tol = 2e-6 * K ** 0.5
    for ta in (0, 1):
        for tb in (0, 1):
            Ad = (A.t().contiguous() if ta else A).to(d)
            Bd = (Bm.t().contiguous() if tb else Bm).to(d)
            lda = M if ta else K
            ldb = N if tb else K
            for use_ws in (True, False):
                C.fill_(float("nan"))
                rn.call("rat_sgemm", Ad, Bd, C, bias.to(d), M, N, K, lda, ldb, N, ta, tb, ws if use_ws else None,
                        ws.numel() * 4 if use_ws else 0, rn.current_stream())
                assert_close(f"sgemm ta={ta} tb={tb} ws={use_ws}", C, want, 1e-5, tol * 4)
    set_precision("fp16")     # back to the library default


@pytest.mark.parametrize("M,N,K", [(4096, 400, 520), (4096, 520, 400), (400, 520, 4096), (200, 48, 100), (130, 16, 64)])
def test_gemm_tcgen05_all_layouts(rn, M, N, K):
    """tcgen05 GEMM (fp16 operands, fp32 TMEM accumulate) incl. MN-major (reduction-index-major) operands and split-K
    vs float64: error of a K-term dot product of fp16-rounded O(1) operands ~ 2^-11 sqrt(K) => atol 0.004 sqrt(K).
    rat_sgemm_scaled with tiny A (gradient operand, 1e-6 scale) keeps the same RELATIVE accuracy (dynamic power-of-two
    lift from max|A|); without the lift the operand would sit in the fp16 subnormal range."""
    from tests.gpu_util import assert_close
    from rat_native.engine import set_precision
    g = torch.Generator().manual_seed(M + N + K)
    d = _dev()
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    want = (A.double() @ Bm.double().t() + bias.double()).float()
    nbytes = int(rn.query("rat_sgemm_workspace_bytes", M, N, K))
    ws = torch.empty(max(nbytes // 4, 4), device=d)
    C = torch.empty(M, N, device=d)
    set_precision("fp16")
    try:
        for ta in (0, 1):
            for tb in (0, 1):
                Ad = (A.t().contiguous() if ta else A).to(d)
                Bd = (Bm.t().contiguous() if tb else Bm).to(d)
                lda = M if ta else K
                ldb = N if tb else K
                for use_ws in (True, False):
                    C.fill_(float("nan"))
                    rn.call("rat_sgemm", Ad, Bd, C, bias.to(d), M, N, K, lda, ldb, N, ta, tb, ws if use_ws else None,
                            ws.numel() * 4 if use_ws else 0, rn.current_stream())
                    assert_close(f"gemm_tc ta={ta} tb={tb} ws={use_ws}", C, want, 2e-2, 0.004 * K ** 0.5)
                am = torch.zeros(1, device=d)
                As = Ad * 1e-6
                rn.call("rat_absmax", As, As.shape[0], As.shape[1], As.shape[1], am, rn.current_stream())
                rn.call("rat_sgemm_scaled", As, Bd, C, None, M, N, K, lda, ldb, N, ta, tb, am, ws, ws.numel() * 4,
                        rn.current_stream())
                assert_close(f"gemm_tc scaled ta={ta} tb={tb}", C * 1e6, want - bias, 2e-2, 0.004 * K ** 0.5)
    finally:
        set_precision("fp16")     # back to the library default


def test_bn_act_forward(rn):
    from tests.gpu_util import assert_close
    d = _dev()
    rows, C = 4096, 400
    z = torch.randn(rows, C) * 2 + 0.5
    gamma, beta = torch.randn(C), torch.randn(C)
    rm, rv = torch.zeros(C), torch.ones(C)
    want = torch.relu(torch.nn.functional.batch_norm(z, rm, rv, gamma, beta, True, 0.1, 1e-5))
    sums = torch.empty(2 * C, dtype=torch.float64, device=d)
    mean, rstd = torch.empty(C, device=d), torch.empty(C, device=d)
    rmd, rvd = torch.zeros(C, device=d), torch.ones(C, device=d)
    out = torch.empty(rows, C, device=d)
    st = rn.current_stream()
    rn.call("rat_bn_sums", z.cuda(), rows, C, sums, st)
    rn.call("rat_bn_finalize", sums, float(rows), C, mean, rstd, rmd, rvd, 0.1, 1e-5, st)
    rn.call("rat_bn_act_fwd", z.cuda(), mean, rstd, gamma.cuda(), beta.cuda(), out, rows, C, 0.0, 0, 0, st)
    assert_close("bn_relu", out, want, 1e-5, 1e-5)
    assert_close("running_mean", rmd, rm, 1e-5, 1e-6)
    assert_close("running_var", rvd, rv, 1e-5, 1e-6)
    # eval mode
    want_eval = torch.relu(torch.nn.functional.batch_norm(z, rm, rv, gamma, beta, False, 0.1, 1e-5))
    rn.call("rat_bn_eval_stats", rmd, rvd, C, mean, rstd, 1e-5, st)
    rn.call("rat_bn_act_fwd", z.cuda(), mean, rstd, gamma.cuda(), beta.cuda(), out, rows, C, 0.0, 0, 0, st)
    assert_close("bn_relu_eval", out, want_eval, 1e-5, 1e-5)


# ------------------------------------------------------------------------------------------ whole forward
@pytest.mark.parametrize("name", CASES_M2 + ["rat_m3_small", "rat_m0_small", "rat_m1_small", "rat_m0_kkbox", "rat_m1_kkbox",
                                  "rat_m3_kkbox"])
def test_eval_forward_matches_reference_golden(rn, precision, name):
    """engine forward (eval) vs the y_pred the REFERENCE produced (fixture). fp32: rtol 1e-4 atol 1e-5; tf32 / fp16
    (the benchmarked mode): 3e-3 absolute on the probability."""
    from tests.gpu_util import assert_close, make_engine, ptol
    c = load_case(name)
    params, bufs = split_state(c["sd0"])
    eng = make_engine(c["spec"], params, bufs)
    X, y = c["X"].cuda(), c["y"].cuda()
    ws = eng.load_wire(X, y, training=False)
    y_pred = eng.forward_ids(ws, X.shape[0], X.shape[1], training=False, with_loss=True)
    eng.check_errors()
    tol = (1e-4, 1e-5) if precision == "fp32" else (0.0, 3e-3)
    assert_close("y_pred", y_pred, torch.from_numpy(c["z"]["eval/y_pred"][:, 0]), *tol)
    want_loss = O.bce_mean(torch.from_numpy(c["z"]["eval/y_pred"]), c["y"][:, 0:1].float())
    assert_close("bce", ws["loss"][1:2], want_loss.reshape(1), *((1e-4, 1e-6) if precision == "fp32" else (5e-3, 1e-4)))


@pytest.mark.parametrize("shape,B,K", [("ml", 512, 5), ("kkbox", 192, 5), ("tmall", 160, 5), ("kkbox", 24, 16)])
def test_eval_forward_full_width_vs_oracle(rn, precision, shape, B, K):
    """full-width configs (reduced vocabulary), oracle as checker."""
    from tests.gpu_util import assert_close, make_engine, ptol, rand_params_nontrivial
    spec = O.shape_spec(shape, vocab_scale=0.02)
    params = rand_params_nontrivial(spec, seed=11)
    bufs = O.init_buffers(spec)
    for k in bufs:
        if k.endswith("running_mean"):
            bufs[k] += 0.1 * torch.randn(bufs[k].shape)
        if k.endswith("running_var"):
            bufs[k] *= 1.5
    pool = O.synthetic_pool(spec, 3000, seed=9)
    nbr = O.synthetic_neighbours(B, 3000, K, seed=9)
    X, y = O.assemble_batch(pool[:B], pool, nbr, np.arange(B))
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    with torch.no_grad():
        want, parts = O.forward(params, bufs, spec, X, y, training=False, return_parts=True)
    eng = make_engine(spec, params, bufs)
    ws = eng.load_wire(X.cuda(), y.cuda(), training=False)
    y_pred = eng.forward_ids(ws, B, K + 1, training=False)
    eng.check_errors()
    enc = ws["enc_out"]                 # [B,T,N,D], or [B,T,D] (field token 0 only) from RAT_m2's last block
    assert_close("pooled", enc[:, 0, 0, :] if enc.ndim == 4 else enc[:, 0, :], parts["pooled"], *ptol(3e-4, 3e-5 * float(parts["pooled"].abs().max()), rt=2e-2, at_scale=150.0))
    assert_close("y_pred", y_pred, want[:, 0], *((2e-4, 2e-5) if precision == "fp32" else (0.0, 3e-3)))


@pytest.mark.parametrize("B,D,stride", [(4096, 40, 40 * 14 * 6), (300, 10, 10), (7, 128, 130)])
def test_head_kernel_matches_torch(rn, B, D, stride):
    """rat_head (one warp per sample): sigmoid(fc(pooled) + dnn + lr), BCE with the -100 log clamp, dlogit, denc and the
    stored max|denc| (= max|dlogit| * max|fc.weight|, exact) vs float64 torch."""
    from tests.gpu_util import assert_close
    torch.manual_seed(B + D)
    d, st = _dev(), rn.current_stream()
    enc = torch.randn(B, stride, device=d)
    w, b = torch.randn(D, device=d) * 0.3, torch.randn(1, device=d)
    dnn, lr = torch.randn(B, device=d), torch.randn(B, device=d) * 0.1
    y = (torch.rand(B, device=d) < 0.4).float()
    y_pred, dlogit = torch.empty(B, device=d), torch.empty(B, device=d)
    denc = torch.zeros(B, stride, device=d)
    nb = int(rn.query("rat_head_blocks", B))
    part = torch.empty(2 * nb, dtype=torch.float64, device=d)
    loss = torch.empty(2, device=d)
    amax = torch.full((1,), -1.0, device=d)
    rn.call("rat_head", enc, stride, w, b, dnn, lr, y, B, D, y_pred, dlogit, denc, 1.0 / B, part, loss[0:1], loss[1:2], amax, st)
    logit = (enc[:, :D].double() @ w.double()) + b.double() + dnn.double() + lr.double()
    p = torch.sigmoid(logit)
    assert_close("y_pred", y_pred, p.float(), 1e-5, 1e-6)
    bce = -(y.double() * torch.log(p).clamp_min(-100) + (1 - y.double()) * torch.log(1 - p).clamp_min(-100))
    assert_close("loss sum", loss[0:1], bce.sum().float().view(1), 1e-5, 1e-5)
    assert_close("loss mean", loss[1:2], bce.mean().float().view(1), 1e-5, 1e-6)
    assert_close("dlogit", dlogit, ((p - y.double()) / B).float(), 1e-4, 1e-9)
    assert torch.equal(denc[:, :D], dlogit[:, None] * w[None, :])
    assert float(denc[:, D:].abs().max()) == 0.0 if stride > D else True
    assert float(amax) == float(denc.abs().max())
    # eval form: no labels, no gradient outputs
    y2 = torch.empty(B, device=d)
    rn.call("rat_head", enc, stride, w, b, None, None, None, B, D, y2, None, None, 1.0, None, None, None, None, st)
    assert_close("y_pred (no dnn / lr)", y2, torch.sigmoid(enc[:, :D].double() @ w.double() + b.double()).float(), 1e-5, 1e-6)
