"""The stage functions above mf_fit! (pathmatfac_b200/staging.py; src/fit.jl:82-1018), executed end to end on the
CPU through tests/oracle_backend.py: regulariser swapping, freezing, call order and the bookkeeping between the
calls are the product's code, the numerical work of every call is the oracle's."""
import numpy as np
import pytest

import pathmatfac_b200 as P
from pathmatfac_b200 import staging as S
from pathmatfac_b200.layers import FrozenLayer
from pathmatfac_b200.regularizers import FrozenRegularizer
from oracle import pmf_oracle as O
from tests import oracle_backend as OB
from tests.helpers import make_pair, random_graphs


def _names():
    return [c[0] for c in OB.CALLS]


def test_basic_fit_without_batch_layers():
    views = {"mutation": ("bernoulli", 14), "methylation": ("normal", 22), "counts": ("poisson", 10)}
    model, om, D = make_pair(70, views, K=3, seed=31, missing=0.2, lambda_X_l2=1.0)
    OB.CALLS.clear()
    hist = []
    S.basic_fit(model, fit_mu=True, fit_logsigma=True, reweight_losses=True, init_factors_=True, fit_factors=True,
                whiten_=True, svd_rotate=True, lr=0.2, max_epochs=25, history=hist, backend=OB.BACKEND, rel_tol=1e-9)
    assert _names() == ["init_mu", "init_logsigma", "reweight_col_losses", "mf_fit_adapt_lr", "mf_fit_adapt_lr"]
    fits = [c[1] for c in OB.CALLS if c[0] == "mf_fit_adapt_lr"]
    assert all(f["flags"] == ("update_X", "update_Y") and f["min_lr"] == 0.05 and f["max_epochs"] == 25 for f in fits)
    assert not any(isinstance(l, FrozenLayer) for l in model.matfac.col_transform.layers)
    assert np.allclose(np.sqrt(np.mean(model.matfac.X ** 2, axis=1)), np.linalg.norm(model.matfac.X, axis=1) / np.sqrt(70))
    gram = model.matfac.Y @ model.matfac.Y.T                      # rotate_by_svd came last: rows of Y orthogonal
    assert np.allclose(gram - np.diag(np.diag(gram)), 0, atol=1e-3 * np.abs(gram).max())
    runs = [h for h in hist if "loss" in h and len(h["loss"]) > 1]
    assert len(runs) >= 2 and all(h["loss"][-1] < h["loss"][0] for h in runs)          # every fit stage went downhill
    assert [h["name"] for h in hist if "name" in h and h["name"] in ("init_factors",)] == ["init_factors"]


def test_init_batch_effects_and_basic_fit_with_batch_layers():
    views = {"methylation": ("normal", 24), "mrnaseq": ("normal", 18)}
    model, om, D = make_pair(90, views, K=4, seed=32, batch_views=["methylation", "mrnaseq"], n_batches=3,
                             n_conditions=3, missing=0.15, lambda_X_l2=1.0)
    X0, Y0, xreg0 = model.matfac.X.copy(), model.matfac.Y.copy(), model.matfac.X_reg
    ct = model.matfac.col_transform
    for v in ct.layers[3].theta.values:
        v[...] = 0
    for v in ct.layers[1].logdelta.values:
        v[...] = 0
    OB.CALLS.clear()
    S.basic_fit(model, fit_batch=True, batch_method="EM", max_epochs=40, lr_regress=0.3, lr_theta=0.5,
                backend=OB.BACKEND)
    names = _names()
    assert names == ["init_mu", "mf_fit_adapt_lr", "mf_fit_adapt_lr", "link_col_sqerr", "batch_stats", "theta_delta_em"]
    calls = dict((n, d) for n, d in OB.CALLS if n != "mf_fit_adapt_lr")
    assert calls["init_mu"]["K"] == 3 and calls["theta_delta_em"]["update_priors"] is True       # K = #conditions in the stand-in
    regress, theta_fit = [d for n, d in OB.CALLS if n == "mf_fit_adapt_lr"]
    assert regress["flags"] == ("update_Y",) and regress["X_reg"] == regress["Y_reg"] == "ZeroReg"
    assert theta_fit["flags"] == ("update_col_layers",) and theta_fit["frozen"] == (True, True, True, False)
    # the model got its own factorisation back, untouched, and batch parameters that are no longer zero
    assert model.matfac.X_reg is xreg0 and np.array_equal(model.matfac.X, X0) and np.array_equal(model.matfac.Y, Y0)
    assert all(np.all(np.isfinite(v)) for v in ct.layers[3].theta.values)
    assert any(np.abs(v).max() > 1e-3 for v in ct.layers[3].theta.values)
    assert all(np.all(np.isfinite(v)) for v in ct.layers[1].logdelta.values)
    assert np.all(np.isfinite(ct.layers[0].logsigma))
    assert not any(isinstance(l, FrozenLayer) for l in ct.layers)
    # the column shift came from the stand-in's M-estimates: the column means where 40 AdaGrad epochs of 0.1 can
    # reach them (methylation, means near 0), on the way there for the mrnaseq columns (means near 10)
    with np.errstate(invalid="ignore"):
        means = np.nanmean(D, axis=0)
    assert np.allclose(ct.layers[2].mu[:24], means[:24], atol=0.05)
    assert np.all(ct.layers[2].mu[24:] > 0.5) and np.all(ct.layers[2].mu[24:] < means[24:])


def test_fit_master_procedure_empirical_bayes():
    """fit! on a model with graph / selective-L1 / group penalties: EB weighting path, post-processing at the end."""
    rng = np.random.default_rng(33)
    views = {"methylation": ("normal", 26), "mrnaseq": ("normal", 20)}
    K, N = 3, 46
    g = random_graphs(N, K, rng, n_virtual=2, n_edges=15)
    model, om, D = make_pair(60, views, K=K, seed=33, n_conditions=2, missing=0.1, lambda_X_l2=1.0, feature_graphs=g,
                             lambda_Y_selective_l1=0.2, lambda_Y_graph=0.5)
    mf = model.matfac
    xreg0, yreg0 = mf.X_reg, mf.Y_reg
    w0 = mf.X_reg.regularizers[0].weights.copy()
    OB.CALLS.clear()
    hist = S.fit(model, lr=0.2, max_epochs=15, keep_history=True, backend=OB.BACKEND, rel_tol=1e-9, abs_tol=1e-9)
    names = _names()
    assert names == ["init_mu", "init_logsigma", "reweight_col_losses", "mf_fit_adapt_lr",       # pre-fit (init_factors)
                     "reweight_col_losses", "mf_fit_adapt_lr",                                      # re-fit, full penalties
                     "reweight_col_losses"]                                                         # post-processing
    pre, full = [d for n, d in OB.CALLS if n == "mf_fit_adapt_lr"]
    assert pre["X_reg"] == "L2Regularizer" and pre["Y_reg"] == "GroupRegularizer"                  # minimal penalties
    assert full["X_reg"] == full["Y_reg"] == "CompositeRegularizer"
    assert mf.X_reg is xreg0 and mf.Y_reg is yreg0
    assert not np.array_equal(mf.X_reg.regularizers[0].weights, w0)                                # reweight_eb! ran
    assert not any(isinstance(r, FrozenRegularizer) for r in mf.col_transform_reg.regs)
    assert np.allclose(np.sqrt(np.mean(mf.X ** 2, axis=1)), 1, rtol=1e-4)                          # whiten!
    assert np.all(np.diff(np.sum(mf.Y ** 2, axis=1)) <= 1e-6)                                      # reorder_by_importance!
    assert [h["name"] for h in hist if h.get("name") in ("start", "reweight_eb", "reorder_factors", "finish")] == \
        ["start", "reweight_eb", "reorder_factors", "finish"]


def test_fit_ard_and_featureset_ard_paths():
    views = {"methylation": ("normal", 20), "mrnaseq": ("normal", 16)}
    model, om, D = make_pair(50, views, K=3, seed=34, missing=0.1, Y_ard=True)
    OB.CALLS.clear()
    S.fit(model, lr=0.2, max_epochs=10, backend=OB.BACKEND)
    names = _names()
    assert names == ["init_mu", "init_logsigma", "reweight_col_losses", "mf_fit_adapt_lr", "reweight_col_losses",
                     "mf_fit_adapt_lr", "reweight_col_losses"]
    pre, ard = [d for n, d in OB.CALLS if n == "mf_fit_adapt_lr"]
    assert pre["X_reg"] == "ZeroReg" and pre["Y_reg"] == "GroupRegularizer" and pre["min_lr"] == 0.05
    assert ard["Y_reg"] == "ARDRegularizer" and ard["min_lr"] == 0.01
    assert model.matfac.Y_reg.alpha == [0.001, 0.001]                                              # reweight_eb!(::ARDRegularizer)

    fsets = {"methylation": [list(range(1, 8)), list(range(5, 15))], "mrnaseq": [list(range(21, 30))]}
    model, om, D = make_pair(50, views, K=3, seed=35, missing=0.1, feature_sets=fsets)
    fs_reg = model.matfac.Y_reg
    OB.CALLS.clear()
    S.fit(model, lr=0.2, max_epochs=8, fsard_max_iter=2, fsard_max_A_iter=30, backend=OB.BACKEND)
    names = _names()
    assert names.count("update_A") == 2 and names[-1] == "reweight_col_losses"
    assert model.matfac.Y_reg is fs_reg
    first_update = names.index("update_A")
    assert [d["Y_reg"] for n, d in OB.CALLS[:first_update] if n == "mf_fit_adapt_lr"] == ["GroupRegularizer", "ARDRegularizer"]
    assert [d["Y_reg"] for n, d in OB.CALLS[first_update:] if n == "mf_fit_adapt_lr"] == ["FeatureSetARDReg"]


def test_reference_fit_tests_featureset_ard_fit_cpu():
    """The reference's own integration test, the one ``fit_tests`` still runs (test/runtests.jl:1134-1171, 1321-1346:
    "Featureset ARD fit CPU"), transcribed: 40 x 60, K = 4, two views x four row batches, five feature sets per view,
    ``Y_fsard`` with ``fsard_v0 = 0.5``, then ``fit!(model; lr=0.05, max_epochs=1000, rel_tol=1e-5, abs_tol=1e-5, fsard_term_rtol=1e-3,
    fsard_max_iter=10, fsard_max_A_iter=500)`` and its three assertions: X changed, Y changed, the batch scales changed.
    Every device-touching call is served by the oracle (tests/oracle_backend.py), the staging is the product's."""
    M, N, K = 40, 60, 4
    rng = np.random.default_rng(1134)
    Z = (rng.standard_normal((K, M)).T @ rng.standard_normal((K, N))).astype(np.float32)
    sample_ids = [f"sample_{i}" for i in range(1, M + 1)]
    sample_conditions = ["condition_1"] * (M // 2) + ["condition_2"] * (M // 2)
    feature_ids = [f"x_{i}" for i in range(1, N + 1)]
    feature_views = [1] * (N // 2) + [2] * (N // 2)
    batch_dict = {j: [f"rowbatch{i}" for i in range(1, 5) for _ in range(M // 4)] for j in (1, 2)}
    blocks = [[range(1, 6), range(6, 11), range(11, 16), range(16, 21), range(21, 31)],
              [range(31, 36), range(36, 41), range(41, 46), range(46, 51), range(51, 61)]]
    feature_sets = {v + 1: [{f"x_{i}" for i in s} for s in blk] for v, blk in enumerate(blocks)}
    model = P.PathMatFacModel(Z, K=4, sample_ids=sample_ids, sample_conditions=sample_conditions, feature_views=feature_views,
                              feature_ids=feature_ids, batch_dict=batch_dict, feature_sets_dict=feature_sets, Y_fsard=True,
                              fsard_v0=0.5, rng=np.random.default_rng(7))
    assert isinstance(model.matfac.Y_reg, P.FeatureSetARDReg) and model.matfac.Y_reg.v0 == np.float32(0.5)
    assert [S_.shape for S_ in model.matfac.Y_reg.S] == [(5, 30), (5, 30)]
    X_start, Y_start = model.matfac.X.copy(), model.matfac.Y.copy()
    logdelta_start = [v.copy() for v in model.matfac.col_transform.layers[1].logdelta.values]
    OB.CALLS.clear()
    S.fit(model, lr=0.05, max_epochs=1000, rel_tol=1e-5, abs_tol=1e-5, fsard_term_rtol=1e-3, fsard_max_iter=10,
          fsard_max_A_iter=500, backend=OB.BACKEND)
    assert not np.allclose(model.matfac.X, X_start)                                            # :1340
    assert not np.allclose(model.matfac.Y, Y_start)                                            # :1341
    assert not all(np.allclose(a, b) for a, b in zip(logdelta_start, model.matfac.col_transform.layers[1].logdelta.values))
    names = _names()
    assert "theta_delta_em" in names or "init_batch_effects" in names or any("batch" in n for n in names), names
    assert names.count("update_A") >= 1 and all(np.isfinite(model.matfac.Y).ravel())
    assert np.all(model.matfac.Y_reg.beta > 0)
