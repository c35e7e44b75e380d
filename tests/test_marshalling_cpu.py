"""What ``fit.Engine`` hands to the C ABI, checked on CPU against a recording stand-in for libpmf: array layouts (the
reference's column-major K x M / K x N / M x N arrays as [M][K] / [N][K] / [N][M] buffers), 0-based half-open ranges, the
frozen masks, mixture weights, closure regularisers -- and the ROW-SHARD logic of the sample-sharded multi-GPU path
(``rows=``): a rank's X columns, its rows of the data and of the batch index, the condition groups clipped to its block,
the refusal of a sample-coupling penalty.  The GPU suite exercises the same code against the real library; this file
makes the N > 1 marshalling fail on CPU when it breaks."""
import ctypes as C

import numpy as np
import pytest

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib, fit as F


class RecordingLib:
    """Every entry point returns 0 and records (name, args); pointer arguments are kept as ctypes pointers and decoded
    by the test with the sizes the header documents."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("pmf_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            if name == "pmf_last_error":
                return b""
            if name == "pmf_default_fit_opts":
                return None
            return 0
        return fn

    def of(self, name):
        return [a for n, a in self.calls if n == name]


def arr(ptr, n, dtype=np.float32):
    if ptr is None:
        return None
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype).copy()


@pytest.fixture
def lib(monkeypatch):
    rec = RecordingLib()
    monkeypatch.setattr(_lib, "load", lambda: rec)
    monkeypatch.setattr(F._lib, "load", lambda: rec)
    return rec


def _model(M=20, K=3, **kw):
    rng = np.random.default_rng(0)
    views = ["methylation"] * 5 + ["mrnaseq"] * 4 + ["mutation"] * 3
    dists = ["normal"] * 9 + ["bernoulli"] * 3
    D = rng.standard_normal((M, 12)).astype(np.float32)
    D[3, 4] = np.nan
    cond = ["c1"] * 8 + ["c2"] * 7 + ["c3"] * (M - 15)
    batch = {"methylation": [f"b{(i * 7) % 3}" for i in range(M)], "mrnaseq": [f"p{i % 2}" for i in range(M)]}
    model = P.PathMatFacModel(D, K=K, feature_views=views, feature_distributions=dists, sample_conditions=cond,
                              batch_dict=batch, **kw)
    mf = model.matfac
    mf.X[...] = rng.standard_normal(mf.X.shape)
    mf.Y[...] = rng.standard_normal(mf.Y.shape)
    ct = mf.col_transform
    ct.layers[0].logsigma[...] = rng.standard_normal(12)
    ct.layers[2].mu[...] = rng.standard_normal(12)
    for v in ct.layers[1].logdelta.values + ct.layers[3].theta.values:
        v[...] = rng.standard_normal(v.shape)
    return model


def test_full_model_layouts_ranges_and_masks(lib):
    model = _model(lambda_X_l2=0.5)
    M, N, K = 20, 12, 3
    P.freeze_layer(model.matfac.col_transform, [2, 4])
    eng = P.Engine(model)
    dims = lib.of("pmf_create")[0][0]._obj
    assert (dims.M, dims.N, dims.K, dims.device) == (M, N, K, 0)
    # data: the reference's column-major M x N = [N][M]; NaN stays NaN
    (h, p), = lib.of("pmf_set_data")
    A = arr(p, N * M).reshape(N, M)
    assert np.array_equal(A, np.asarray(model.data).T, equal_nan=True) and np.isnan(A).sum() == 1
    # factors: K x M / K x N column-major = [M][K] / [N][K]
    (h, px, py), = lib.of("pmf_set_factors")
    assert np.array_equal(arr(px, M * K).reshape(M, K), model.matfac.X.T)
    assert np.array_equal(arr(py, N * K).reshape(N, K), model.matfac.Y.T)
    # noise ranges: sorted (distribution, view) blocks, 0-based half-open; one code per range
    h, nr, cs, ce, dc, th, w = lib.of("pmf_set_noise")[0]
    assert nr == 2 and list(arr(cs, 2, np.int32)) == [0, 3] and list(arr(ce, 2, np.int32)) == [3, 12]
    assert list(arr(dc, 2, np.int32)) == [1, 0] and np.array_equal(arr(w, N), np.ones(N, np.float32))   # bernoulli, normal
    assert model.feature_distributions[:3] == ["bernoulli"] * 3                    # the constructor sorted the columns
    # batch layout: two batched views after the bernoulli block, batch ordinals in first-appearance order
    h, nv, bcs, bce, nb, bos = lib.of("pmf_set_batch_layout")[0]
    assert nv == 2 and list(arr(bcs, 2, np.int32)) == [3, 8] and list(arr(bce, 2, np.int32)) == [8, 12]
    assert list(arr(nb, 2, np.int32)) == [3, 2]
    B = arr(bos, 2 * M, np.int32).reshape(2, M)
    ld = model.matfac.col_transform.unwrapped(1).logdelta
    assert np.array_equal(B[0], ld.batch_index[0]) and np.array_equal(B[1], ld.batch_index[1])
    assert list(B[0][:4]) == [0, 1, 2, 0] and list(B[1][:4]) == [0, 1, 0, 1]
    # batch values: n_b x N_v column-major = [N_v][n_b]
    vals = lib.of("pmf_set_batch_values")
    assert [a[1] for a in vals] == [0, 1]
    assert np.array_equal(arr(vals[0][2], 15).reshape(5, 3), ld.values[0].T)
    th_ba = model.matfac.col_transform.unwrapped(3).theta
    assert np.array_equal(arr(vals[1][3], 8).reshape(4, 2), th_ba.values[1].T)
    # frozen masks: bit s-1 per slot s; layers 2 and 4 frozen
    h, fl, fr = lib.of("pmf_set_frozen")[0]
    assert fl == 0b1010 and fr == 0
    # X penalty: with conditions AND lambda_X_l2 the constructor mixes L2 and the condition groups half and half
    l2 = [a for a in lib.of("pmf_set_reg_l2") if a[1] == 0]
    grp = [a for a in lib.of("pmf_set_reg_group") if a[1] == 0]
    assert len(l2) >= 1 and len(grp) >= 1
    assert l2[0][3] == pytest.approx(0.5) and grp[0][6] == pytest.approx(0.5)      # mixture weights (regularizers.jl:655-689)
    assert np.allclose(arr(l2[0][2], K), 0.5)
    ng = grp[0][2]
    assert ng == 3 and list(arr(grp[0][3], 3, np.int32)) == [0, 8, 15] and list(arr(grp[0][4], 3, np.int32)) == [8, 15, 20]
    eng.close()


def test_row_shard_marshalling(lib):
    """rank 1 of 3 on 20 samples owns rows 6..12: its X columns, its rows of the data and of the batch index, the
    condition groups clipped to the block (c1 = 0..8 -> [0, 2), c2 = 8..15 -> [2, 7) in local coordinates)."""
    from pathmatfac_b200.dist import shard_rows
    model = _model()
    rows = shard_rows(20, 1, 3)
    assert (rows.start, rows.stop) == (6, 13)
    eng = P.Engine(model, rows=rows)
    Ms, N, K = 7, 12, 3
    dims = lib.of("pmf_create")[0][0]._obj
    assert (dims.M, dims.N, dims.K) == (Ms, N, K)
    A = arr(lib.of("pmf_set_data")[0][1], N * Ms).reshape(N, Ms)
    assert np.array_equal(A, np.asarray(model.data)[6:13].T, equal_nan=True)
    (h, px, py), = lib.of("pmf_set_factors")
    assert np.array_equal(arr(px, Ms * K).reshape(Ms, K), model.matfac.X[:, 6:13].T)
    assert np.array_equal(arr(py, N * K).reshape(N, K), model.matfac.Y.T)                 # Y is replicated
    h, nv, bcs, bce, nb, bos = lib.of("pmf_set_batch_layout")[0]
    B = arr(bos, 2 * Ms, np.int32).reshape(2, Ms)
    ld = model.matfac.col_transform.unwrapped(1).logdelta
    assert np.array_equal(B[0], ld.batch_index[0][6:13]) and np.array_equal(B[1], ld.batch_index[1][6:13])
    assert list(arr(nb, 2, np.int32)) == [3, 2]                                           # batch tables stay whole
    grp = [a for a in lib.of("pmf_set_reg_group") if a[1] == 0][0]
    assert grp[2] == 2 and list(arr(grp[3], 2, np.int32)) == [0, 2] and list(arr(grp[4], 2, np.int32)) == [2, 7]
    # downloads land in the rank's columns only
    lib.calls.clear()
    before = model.matfac.X.copy()
    eng.pull_params()
    (h, px, py), = lib.of("pmf_get_factors")
    assert np.array_equal(model.matfac.X[:, :6], before[:, :6]) and np.array_equal(model.matfac.X[:, 13:], before[:, 13:])
    eng.close()


def test_sample_coupling_penalty_is_refused_on_a_shard(lib):
    """a NetworkRegularizer on X (sample_graphs) couples samples -- replicas only."""
    rng = np.random.default_rng(1)
    M, N, K = 12, 6, 2
    sids = [f"s{i}" for i in range(M)]
    graphs = [[[sids[i], sids[i + 1], 1.0] for i in range(M - 1)] for _ in range(K)]
    model = P.PathMatFacModel(rng.standard_normal((M, N)).astype(np.float32), sample_ids=sids, sample_graphs=graphs,
                              lambda_X_graph=1.0)
    P.Engine(model).close()                                   # the whole problem on one handle: fine
    assert any(a[1] == 0 for a in lib.of("pmf_set_reg_network"))
    with pytest.raises(_lib.PmfError, match="not shardable"):
        P.Engine(model, rows=range(0, 6))


def test_closure_and_composite_regularisers(lib):
    model = _model()
    K = 3
    model.matfac.X_reg = lambda X: 0.5 * np.float32(0.3) * float((X * X).sum())        # src/fit.jl:686-style closure
    model.matfac.Y_reg = lambda Y: 0.0                                                   # the zero closures
    eng = P.Engine(model)
    l2 = lib.of("pmf_set_reg_l2")
    assert [a[1] for a in l2] == [0] and np.allclose(arr(l2[0][2], K), 0.3, rtol=1e-5) and l2[0][3] == 1.0
    assert [a[1] for a in lib.of("pmf_clear_reg")] == [0, 1]
    model.matfac.X_reg = lambda X: float(np.abs(X).sum())                                # not one of the reference's
    with pytest.raises(_lib.PmfError, match="unsupported regulariser closure"):
        eng.push_regs()
    eng.close()


class ScriptedLib(RecordingLib):
    """RecordingLib whose pmf_fit plays a script of (term_code, last epoch) results."""

    def __init__(self, script):
        super().__init__()
        self.script = list(script)
        self.fits = []

    def __getattr__(self, name):
        base = super().__getattr__(name)
        if name != "pmf_fit":
            return base

        def fit(h, opts, hist):
            o, hh = opts._obj, hist._obj
            term, last = self.script.pop(0)
            self.fits.append(dict(epoch=o.epoch, max_epochs=o.max_epochs, lr=o.lr, update_X=o.update_X, update_Y=o.update_Y,
                                  update_col_layers=o.update_col_layers, update_noise_models=o.update_noise_models,
                                  rel_tol=o.rel_tol))
            n = min(last - o.epoch + 1, hh.capacity)
            for i in range(n):
                hh.loss_total[i] = 100.0 - (o.epoch + i)
            hh.term_code, hh.epochs, hh.n_recorded, hh.kernel_launches = term, last, n, 2 * n
            return base(h, opts, hist)
        return fit


def test_restart_policy_of_mf_fit_adapt_lr(monkeypatch):
    """src/fit.jl:46-75 in the mirror, on CPU: "loss_increase" halves eta and resumes AT the epoch that raised it with
    the SAME optimiser (accumulators reset once, when the optimiser first meets the handle); any other code ends the
    loop; so does eta < min_lr.  The GPU suite checks the same against the real library (test_lr_halving_restart_policy)."""
    LOSS_INCREASE, MAX_EPOCHS, REL_TOL = 3, 0, 2
    rec = ScriptedLib([(LOSS_INCREASE, 7), (LOSS_INCREASE, 12), (REL_TOL, 30)])
    monkeypatch.setattr(_lib, "load", lambda: rec)
    model = _model()
    hs = P.mf_fit_adapt_lr(model, lr=0.8, min_lr=0.01, max_epochs=50, update_X=True, update_Y=True, verbosity=0, rel_tol=1e-7)
    assert [h["term_code"] for h in hs] == ["loss_increase", "loss_increase", "rel_tol"]
    assert [f["epoch"] for f in rec.fits] == [1, 7, 12] and all(f["max_epochs"] == 50 for f in rec.fits)
    assert np.allclose([f["lr"] for f in rec.fits], [0.8, 0.4, 0.2])
    assert all(f["update_X"] == 1 and f["update_Y"] == 1 and f["update_col_layers"] == 0 for f in rec.fits)
    assert all(f["update_noise_models"] == 1 and f["rel_tol"] == 1e-7 for f in rec.fits)       # src/fit.jl:14 default
    assert len(rec.of("pmf_reset_opt_state")) == 1                  # AdaGrad state survives the restarts (:55-64)
    assert [h["name"] for h in hs] == ["mf_fit_lr=0.8", "mf_fit_lr=0.4", "mf_fit_lr=0.2"] and hs[0]["epochs"] == 7
    assert hs[1]["loss"] == [100.0 - e for e in range(7, 13)]
    assert model._engine is None                                    # the residency taken for the call was released
    # eta below min_lr ends the loop even though the last call asked for a restart
    rec2 = ScriptedLib([(LOSS_INCREASE, 3), (LOSS_INCREASE, 4), (LOSS_INCREASE, 5)])
    monkeypatch.setattr(_lib, "load", lambda: rec2)
    hs = P.mf_fit_adapt_lr(_model(), lr=0.1, min_lr=0.04, max_epochs=50, update_Y=True, verbosity=0)
    assert len(hs) == 2 and np.allclose([f["lr"] for f in rec2.fits], [0.1, 0.05])
    # a history list handed in is appended to (src/fit.jl:61)
    rec3 = ScriptedLib([(MAX_EPOCHS, 9)])
    monkeypatch.setattr(_lib, "load", lambda: rec3)
    hist = [{"name": "earlier"}]
    P.mf_fit_adapt_lr(_model(), lr=1.0, max_epochs=9, update_X=True, verbosity=0, history=hist)
    assert [h["name"] for h in hist] == ["earlier", "mf_fit_lr=1.0"] and hist[1]["term_code"] == "max_epochs"


def test_transform_host_logic(monkeypatch):
    """src/transform.jl:6-106 in the mirror, on CPU: new columns are matched to the training features by id (unmatched
    training columns become NaN = missing, unknown new columns are dropped), batch layers and factor penalties are dropped,
    the column layers and their penalties frozen, X starts at zero and is the only thing fitted; the training model keeps
    its data.  (GPU: test_transform_new_samples.)"""
    rec = ScriptedLib([(2, 17)])
    monkeypatch.setattr(_lib, "load", lambda: rec)
    model = _model(lambda_X_l2=1.0)
    N, K = 12, 3
    ids = [f"f{j}" for j in range(N)]
    model.feature_ids = list(ids)
    rng = np.random.default_rng(5)
    new_ids = [f"f{j}" for j in range(5, 15)]                       # f5..f11 known, f12..f14 unknown
    D_new = rng.standard_normal((6, 10)).astype(np.float32)
    data_before = np.asarray(model.data).copy()
    res = P.transform(model, D_new, feature_ids=new_ids, verbosity=0, lr=0.05, max_epochs=40)
    assert np.array_equal(np.asarray(model.data), data_before, equal_nan=True) and model._engine is None
    assert res.matfac.X.shape == (K, 6) and res.matfac.Y.shape == (K, N) and res.sample_ids == [1, 2, 3, 4, 5, 6]
    dims = rec.of("pmf_create")[0][0]._obj
    assert (dims.M, dims.N, dims.K) == (6, N, K)
    A = arr(rec.of("pmf_set_data")[0][1], N * 6).reshape(N, 6).T          # back to samples x features
    assert np.isnan(A[:, :5]).all() and np.array_equal(A[:, 5:], D_new[:, :7])
    px = rec.of("pmf_set_factors")[0][1]
    assert not arr(px, 6 * K).any()                                        # X = 0
    assert rec.of("pmf_set_batch_layout")[0][1] == 0                       # batch effects are ignored on new data
    assert not rec.of("pmf_set_reg_l2") and not rec.of("pmf_set_reg_group")        # X_reg, Y_reg dropped
    masks = rec.of("pmf_set_frozen")[-1]
    # column layers and their penalties frozen (slots 2 and 4 are identities here: nothing to freeze on the device)
    assert masks[1] & 0b0101 == 0b0101 and masks[2] & 0b0101 == 0b0101
    (f,) = rec.fits
    assert (f["update_X"], f["update_Y"], f["update_col_layers"]) == (1, 0, 0) and f["max_epochs"] == 40
    assert f["lr"] == pytest.approx(0.05)
    with pytest.raises(AssertionError, match="Provide `feature_ids`"):
        P.transform(model, D_new, verbosity=0)                             # 10 columns against 12, no ids


def test_graph_selective_l1_and_fsard_marshalling(lib):
    """The Y-side regularisers with structure: the K per-factor Laplacians as concatenated CSR blocks AA / AB / BB (each
    block's row pointer starts at 0), the selective-L1 mask as [N][K] bytes, the FSARD beta as [N][K] -- rebuilt here
    from what the ABI received and compared with the objects (src/regularizers.jl:187-240, 106-163;
    src/featureset_ard.jl:19-65)."""
    import scipy.sparse as sp
    from tests.helpers import random_graphs
    rng = np.random.default_rng(3)
    M, N, K = 10, 14, 3
    graphs = random_graphs(N, K, rng, n_virtual=2, n_edges=12)
    D = rng.standard_normal((M, N)).astype(np.float32)
    model = P.PathMatFacModel(D, feature_graphs=graphs, lambda_Y_graph=0.7, lambda_Y_selective_l1=0.2, lambda_Y_l2=None)
    regs = {type(r).__name__: (r, p) for r, p in zip(model.matfac.Y_reg.regularizers, model.matfac.Y_reg.mixture_p)}
    net, p_net = regs["NetworkRegularizer"]
    sel, p_sel = regs["SelectiveL1Reg"]
    eng = P.Engine(model)
    a = [c for c in lib.of("pmf_set_reg_network") if c[1] == 1][0]
    nv = arr(a[2], K, np.int32)
    assert list(nv) == [b.shape[0] for b in net.BB] and a[13] == pytest.approx(float(p_net))

    def blocks(rp_ptr, ci_ptr, va_ptr, rows, cols):
        out, rp_off, nz_off = [], 0, 0
        for k in range(K):
            rp = arr(rp_ptr, rp_off + rows[k] + 1, np.int32)[rp_off:]
            assert rp[0] == 0
            nnz = int(rp[-1])
            ci = arr(ci_ptr, nz_off + nnz, np.int32)[nz_off:] if nnz else np.zeros(0, np.int32)
            va = arr(va_ptr, nz_off + nnz)[nz_off:] if nnz else np.zeros(0, np.float32)
            out.append(sp.csr_matrix((va, ci, rp), shape=(rows[k], cols[k])))
            rp_off += rows[k] + 1
            nz_off += nnz
        return out
    for got, ref in ((blocks(a[3], a[4], a[5], [N] * K, [N] * K), net.AA), (blocks(a[6], a[7], a[8], [N] * K, list(nv)), net.AB),
                     (blocks(a[9], a[10], a[11], list(nv), list(nv)), net.BB)):
        for g, r in zip(got, ref):
            assert np.allclose(g.toarray(), np.asarray(r.todense(), dtype=np.float32))
    s = [c for c in lib.of("pmf_set_reg_sel_l1") if c[1] == 1][0]
    mask = np.ctypeslib.as_array(s[2], shape=(N * K,)).reshape(N, K)
    assert np.array_equal(mask.T.astype(bool), np.asarray(sel.l1_idx, dtype=bool)) and s[4] == pytest.approx(float(p_sel))
    assert mask.any() and not mask.all()
    eng.close()
    # feature-set ARD: alpha (N), beta K x N column-major = [N][K]
    fids = [f"g{j}" for j in range(N)]
    views = ["a"] * 6 + ["b"] * 8
    fsets = {"a": [fids[0:3], fids[2:6]], "b": [fids[6:10], fids[9:14]]}
    model = P.PathMatFacModel(D.copy(), K=K, feature_ids=fids, feature_views=views, feature_sets_dict=fsets, Y_fsard=True)
    reg = model.matfac.Y_reg
    reg.beta[...] = 0.5 + rng.random(reg.beta.shape).astype(np.float32)
    reg.alpha[...] = 1.0 + rng.random(N).astype(np.float32)
    lib.calls.clear()
    P.Engine(model).close()
    f = [c for c in lib.of("pmf_set_reg_fsard") if c[1] == 1][0]
    assert np.array_equal(arr(f[2], N), reg.alpha) and np.array_equal(arr(f[3], N * K).reshape(N, K), reg.beta.T)
