"""GPU parity tests added in round 2: boundary behaviour around repeated fits on a resident model, closures as
regularisers, frozen layers, and the options that make the MatFac.jl unknowns explicit."""
import numpy as np
import pytest

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from oracle import pmf_oracle as O
from tests.helpers import make_pair, random_graphs as _graphs, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_closure_regulariser_is_recognised_as_l2():
    """src/fit.jl:686 installs `X -> 0.5f0*sum(X.*X)` as X_reg in the default fit! path: the device must apply it
    (an L2 penalty of weight 1), not drop it."""
    model, om, D = make_pair(90, {"methylation": ("normal", 70)}, K=5, seed=201, lambda_X_l2=1.0)
    model.matfac.X_reg = lambda X: np.float32(0.5) * np.sum(X * X)
    om.X_reg = O.L2Regularizer(5, 1.0)
    eng = P.Engine(model)
    try:
        got = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert abs(got["components"]["X_reg"] - ref["components"]["X_reg"]) <= TOL * ref["components"]["X_reg"]
    assert relerr(got["dX"], ref["dX"]) < TOL
    model.matfac.X_reg = lambda X: np.sum(np.abs(X))          # not one of the reference's closures
    with pytest.raises(_lib.PmfError):
        P.Engine(model)


def test_resident_engine_keeps_batch_parameters_and_optimiser_state():
    """gpu(model); non-zero theta / logdelta; reweight_col_losses (which re-sends the structure); mf_fit with the layers
    frozen: the batch parameters must come back unchanged (they used to be zeroed), and a second fit
    with the same optimiser object must continue from the first one's AdaGrad state."""
    views = {"methylation": ("normal", 60), "mrnaseq": ("normal", 40)}
    model, om, D = make_pair(140, views, K=4, seed=202, batch_views=["methylation", "mrnaseq"], n_batches=3, missing=0.2,
                             lambda_X_l2=1.0)
    th0 = [v.copy() for v in model.matfac.col_transform.unwrapped(3).theta.values]
    ld0 = [v.copy() for v in model.matfac.col_transform.unwrapped(1).logdelta.values]
    assert all(np.abs(v).max() > 0 for v in th0)
    P.gpu(model)
    try:
        P.reweight_col_losses(model)
        P.mf_fit(model, lr=0.1, max_epochs=3, update_X=True, update_Y=True, update_col_layers=False, verbosity=0,
                 kernel=_lib.KERNEL_FFMA)
        for v in range(2):
            assert np.array_equal(model.matfac.col_transform.unwrapped(3).theta.values[v], th0[v])
            assert np.array_equal(model.matfac.col_transform.unwrapped(1).logdelta.values[v], ld0[v])
        # one optimiser object across two calls == one call of twice the epochs (AdaGrad state kept, src/fit.jl:55-64)
        opt = P.AdaGrad(0.2)
        h1 = P.mf_fit(model, opt=opt, max_epochs=3, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0,
                      verbosity=0, kernel=_lib.KERNEL_FFMA)
        h2 = P.mf_fit(model, opt=opt, max_epochs=6, epoch=4, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0,
                      abs_tol=0, verbosity=0, kernel=_lib.KERNEL_FFMA)
    finally:
        P.cpu(model)
    model2, _, _ = make_pair(140, views, K=4, seed=202, batch_views=["methylation", "mrnaseq"], n_batches=3, missing=0.2,
                             lambda_X_l2=1.0)
    P.gpu(model2)
    try:
        P.reweight_col_losses(model2)
        P.mf_fit(model2, lr=0.1, max_epochs=3, update_X=True, update_Y=True, update_col_layers=False, verbosity=0,
                 kernel=_lib.KERNEL_FFMA)
        h = P.mf_fit(model2, opt=P.AdaGrad(0.2), max_epochs=6, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0,
                     abs_tol=0, verbosity=0, kernel=_lib.KERNEL_FFMA)
    finally:
        P.cpu(model2)
    assert np.allclose(h1["loss"] + h2["loss"], h["loss"], rtol=2e-5)
    assert relerr(model.matfac.col_transform.unwrapped(3).theta.values[0],
                  model2.matfac.col_transform.unwrapped(3).theta.values[0]) < 1e-4


def test_frozen_layer_drops_its_penalty():
    """ColParamReg / BatchArrayReg of a FrozenLayer evaluate to 0 (src/regularizers.jl:508-510, :887): freezing the
    LAYERS only (init_theta!, the joint fit) must remove their penalties from the loss history too."""
    model, om, D = make_pair(64, {"methylation": ("normal", 50)}, K=4, seed=203, batch_views=["methylation"], n_batches=3,
                             lambda_X_l2=1.0)
    P.freeze_layer(model.matfac.col_transform, [1, 2])
    om.frozen = [True, True, False, False]
    href = O.mf_fit(om, D, O.AdaGrad(0.25), max_epochs=5, update_col_layers=True, rel_tol=0, abs_tol=0)
    h = P.mf_fit(model, lr=0.25, max_epochs=5, update_col_layers=True, rel_tol=0, abs_tol=0, kernel=_lib.KERNEL_FFMA,
                 verbosity=0)
    assert relerr(h["loss"], href["loss"]) < TOL
    assert relerr(h["layer_reg"], [c["layer_reg"] for c in href["components"]]) < TOL


# ---- K > 64 (BASELINE configs[3], configs[4]: K = 256 / K = 128 pathway factors, src/model.jl:122) ----------------
def _wide_views():
    # (distribution, view) sorted: bernoulli < normal < ordinal3 < poisson
    return {"mutation": ("bernoulli", 90), "methylation": ("normal", 170), "cna": ("ordinal3", 50), "counts": ("poisson", 75)}


@pytest.mark.parametrize("K", [72, 128, 256])
def test_wide_k_ffma_against_oracle(K):
    """The FP32 kernel at K > 64 (the on-device reference of the tensor-core path and the kernel of batch-layer models
    there) against the oracle."""
    model, om, D = make_pair(310, _wide_views(), K=K, seed=300 + K, missing=0.3, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        got = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert abs(got["loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    for k in ("dX", "dY", "dmu", "dlogsigma"):
        assert relerr(got[k], ref[k]) < TOL, (k, relerr(got[k], ref[k]))


@pytest.mark.parametrize("K,M", [(72, 310), (128, 523), (200, 257), (256, 300)])
def test_wide_k_tensor_core_against_oracle(K, M):
    """wide_tc.cu (Z + link kernel, the two gradient contractions over G') against the oracle: ragged M / N against the
    256-sample and 128-feature tiles, four noise models, 30% missing.  Loss and column gradients to FP32 level
    (TF32 + BF16 first-order corrections of Z); dX / dY are single-pass TF32 contractions: 3e-4 at this size."""
    model, om, D = make_pair(M, _wide_views(), K=K, seed=310 + K, missing=0.3, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        got = eng.loss_grad(include_reg=True)
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        again = eng.loss_grad(include_reg=True)          # second pass over the same handle (scratch reuse)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert got["dX"].shape == (K, M) and got["dY"].shape == (K, 385)
    assert abs(got["loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"]), (got["loss"], ref["loss"])
    assert relerr(got["dmu"], ref["dmu"]) < TOL and relerr(got["dlogsigma"], ref["dlogsigma"]) < TOL
    assert relerr(got["dY"], ref["dY"]) < 3e-4 and relerr(got["dX"], ref["dX"]) < 3e-4, (relerr(got["dY"], ref["dY"]), relerr(got["dX"], ref["dX"]))
    assert relerr(again["dY"], got["dY"]) < 1e-5 and abs(again["loss"] - got["loss"]) <= 1e-9 * abs(got["loss"])


@pytest.mark.parametrize("K", [128, 256])
def test_wide_k_tensor_core_midsize_and_fit(K):
    """2 000 x 3 000: the tensor-core path against the FP32 kernel on the same handle (every CTA range crosses feature
    tiles and item boundaries), then a short fit through AUTO against the oracle's loss curve."""
    views = {"mutation": ("bernoulli", 700), "methylation": ("normal", 1500), "counts": ("poisson", 800)}
    model, om, D = make_pair(2000, views, K=K, seed=330 + K, missing=0.3, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        ref = eng.loss_grad(include_reg=True)
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        got = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    assert abs(got["loss"] - ref["loss"]) <= 2e-6 * abs(ref["loss"])
    assert relerr(got["dmu"], ref["dmu"]) < 2e-5 and relerr(got["dlogsigma"], ref["dlogsigma"]) < 2e-5
    assert relerr(got["dY"], ref["dY"]) < 2e-4 and relerr(got["dX"], ref["dX"]) < 2e-4, (relerr(got["dY"], ref["dY"]), relerr(got["dX"], ref["dX"]))
    h = P.mf_fit(model, lr=0.05, max_epochs=4, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0,
                 verbosity=0, kernel=_lib.KERNEL_AUTO)
    href = O.mf_fit(om, D, O.AdaGrad(0.05), max_epochs=4, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0,
                    abs_tol=0)
    assert np.max(np.abs(np.array(h["loss"]) / np.array(href["loss"]) - 1)) < 1e-4


def test_wide_k_with_batch_layers_runs_fp32_kernel():
    """K > 64 with batch layers is not served by the tensor-core kernels: AUTO runs the FP32 kernel (parity with the
    oracle), an explicit PMF_KERNEL_TC is an error, never a silent fallback."""
    model, om, D = make_pair(200, {"methylation": ("normal", 130)}, K=72, seed=350, batch_views=["methylation"], n_batches=3,
                             missing=0.2, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_AUTO, 0)
        got = eng.loss_grad(include_reg=True)
        ref = O.total_loss_grads(om, D)
        assert relerr(got["dY"], ref["dY"]) < TOL and relerr(got["dtheta"][0], ref["dtheta"][0]) < TOL
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        with pytest.raises(_lib.PmfError):
            eng.loss_grad(include_reg=False)
    finally:
        eng.close()


# ---- MatFac.jl unknowns as explicit options ---------------------------------------------
ORD_VIEWS = {"mutation": ("bernoulli_sq_hinge", 40), "cna": ("ordinal3", 60), "methylation": ("ordinal_sq_hinge3", 45)}


@pytest.mark.parametrize("kernel,M,K", [(_lib.KERNEL_FFMA, 150, 6), (_lib.KERNEL_TC, 700, 16), (_lib.KERNEL_TC, 300, 72)])
def test_ordinal_threshold_gradients(kernel, M, K):
    """d loss / d(t1, t2) of the ordinal noise models (update_noise_models, src/fit.jl:14) on every kernel against the
    oracle (whose values are finite-difference checked on the CPU, tests/test_oracle_golden.py)."""
    model, om, D = make_pair(M, ORD_VIEWS, K=K, seed=400 + K, missing=0.15, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(kernel, 0)
        got = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert got["dthresholds"].shape == (3, 2)
    assert np.all(got["dthresholds"][0] == 0)                    # bernoulli_sq_hinge has no thresholds
    assert relerr(got["dthresholds"][1:], ref["dthresholds"][1:]) < TOL
    assert abs(got["loss"] - ref["loss"]) <= TOL * abs(ref["loss"])


def test_fit_trains_ordinal_thresholds():
    model, om, D = make_pair(160, ORD_VIEWS, K=5, seed=410, missing=0.1, lambda_X_l2=1.0)
    th0 = [None if n.ext_thresholds is None else n.ext_thresholds.copy() for n in model.matfac.noise_model.noises]
    kw = dict(max_epochs=12, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0)
    # thresholds fixed: nothing moves
    import copy
    m_fixed = copy.deepcopy(model)
    h_fixed = P.mf_fit(m_fixed, lr=0.25, update_noise_models=False, verbosity=0, kernel=_lib.KERNEL_FFMA, **kw)
    for n, t in zip(m_fixed.matfac.noise_model.noises, th0):
        assert t is None or np.array_equal(n.ext_thresholds, t)
    om_fixed = copy.deepcopy(om)
    href_fixed = O.mf_fit(om_fixed, D, O.AdaGrad(0.25), update_noise_models=False, **kw)
    assert relerr(h_fixed["loss"], href_fixed["loss"]) < TOL
    # thresholds trained (the reference's default)
    h = P.mf_fit(model, lr=0.25, verbosity=0, kernel=_lib.KERNEL_FFMA, **kw)
    href = O.mf_fit(om, D, O.AdaGrad(0.25), update_noise_models=True, **kw)
    assert relerr(h["loss"], href["loss"]) < TOL
    assert h["loss"][-1] < h_fixed["loss"][-1]
    for n, t, t_ref in zip(model.matfac.noise_model.noises, th0, om.noise.thresholds):
        if t is None:
            continue
        assert np.isneginf(n.ext_thresholds[0]) and np.isposinf(n.ext_thresholds[3])
        assert not np.array_equal(n.ext_thresholds[1:3], t[1:3])
        assert np.allclose(n.ext_thresholds[1:3], t_ref[1:3], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("kernel,M,N", [(_lib.KERNEL_FFMA, 130, 90), (_lib.KERNEL_TC, 900, 600)])
def test_alternating_epochs(kernel, M, N):
    """`alternating` = column-side step, then the row-side step from a second pass at the new Y.  Same
    loss curve and parameters as the oracle's alternating loop; a different curve from the simultaneous step."""
    views = {"mutation": ("bernoulli", N // 3), "methylation": ("normal", N - N // 3)}
    model, om, D = make_pair(M, views, K=8, seed=420, missing=0.2, lambda_X_l2=1.0, batch_views=["methylation"], n_batches=3)
    import copy
    model_s = copy.deepcopy(model)
    kw = dict(max_epochs=8, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0)
    h = P.mf_fit(model, lr=0.25, alternating=True, verbosity=0, kernel=kernel, **kw)
    href = O.mf_fit(om, D, O.AdaGrad(0.25), alternating=True, **kw)
    tol = TOL if kernel == _lib.KERNEL_FFMA else 3 * TOL
    assert np.max(np.abs(np.array(h["loss"]) / np.array(href["loss"]) - 1)) < tol
    assert relerr(model.matfac.X, om.X) < 30 * tol and relerr(model.matfac.Y, om.Y) < 30 * tol
    h_s = P.mf_fit(model_s, lr=0.25, alternating=False, verbosity=0, kernel=kernel, **kw)
    assert abs(h_s["loss"][3] / h["loss"][3] - 1) > 1e-3
    assert h["kernel_launches"] > h_s["kernel_launches"]


def test_network_regulariser_at_a_3000_feature_cut():
    """NetworkRegularizer (src/regularizers.jl:169-306) on a 3 000-feature cut of the C4 shape: K = 24 per-factor graphs
    with ~800 signed edges and 80 virtual nodes each (virtual-node CG in shared memory), selective L1 beside it; value
    and pullback against the oracle, then a warm-started second evaluation (x_virtual carried over, reference quirk vi)."""
    rng = np.random.default_rng(77)
    N, K = 3000, 24
    graphs = _graphs(N, K, rng, n_virtual=80, n_edges=800)
    model, om, D = make_pair(120, {"mrnaseq": ("normal", N)}, K=K, seed=430, feature_graphs=graphs,
                             lambda_Y_selective_l1=0.5, lambda_Y_graph=1.0)
    for r in model.matfac.Y_reg.regularizers:
        if isinstance(r, P.NetworkRegularizer):
            r.cg_rtol, r.cg_atol = 1e-7, 1e-10
    eng = P.Engine(model)
    try:
        got = eng.loss_grad(include_reg=True)
        again = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert abs(got["components"]["Y_reg"] - ref["components"]["Y_reg"]) <= TOL * abs(ref["components"]["Y_reg"])
    assert relerr(got["dY"], ref["dY"]) < TOL
    assert abs(again["components"]["Y_reg"] - got["components"]["Y_reg"]) <= 1e-5 * abs(got["components"]["Y_reg"])
    assert relerr(again["dY"], got["dY"]) < 1e-5


def test_guard_zones_stay_intact():
    """Own memcheck (compute-sanitizer is closed on this pool): with PMF_GUARD=1 every device buffer of the library has
    1 KiB guard zones; after loss / gradient passes of every kernel family at ragged shapes, statistics passes and fit
    epochs no guard byte may have changed (scripts/sanitize_case.py)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PMF_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "sanitize_case.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if "guards" in l]
    assert len(lines) == 5 and all("corrupt bytes 0" in l for l in lines), r.stdout
    assert all(" buffers, " in l and int(l.split("guards ")[1].split(" buffers")[0]) > 20 for l in lines)


def test_c3_full_size_five_batched_views_tc_against_fp32_kernel():
    """BASELINE configs[2] at full size: 10 000 x 30 000, K = 64, FIVE batched assays x 40 batches (batch ids iid per
    sample and view: five sample orders, ten passes' worth of A_tc), 20 sample conditions.  The tcgen05 batch path
    against the FP32 kernel on the same handle: batch gradients and column gradients to 1e-4, loss to 1e-6."""
    from pathmatfac_b200.simulate import simulate_problem
    blocks = (("mutation", "bernoulli", 5000), ("cna", "normal", 5000), ("methylation", "normal", 7500),
              ("mrnaseq", "normal", 7500), ("counts", "poisson", 5000))
    model = simulate_problem(10000, blocks=blocks, K=64, seed=3, missing=0.3, batch_views=[b[0] for b in blocks], n_batches=40,
                             n_conditions=20)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        ref = eng.loss_grad(include_reg=True)
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        got = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    assert len(got["dtheta"]) == 5
    assert abs(got["loss"] - ref["loss"]) <= 1e-6 * abs(ref["loss"])
    for v in range(5):
        assert relerr(got["dtheta"][v], ref["dtheta"][v]) < TOL, (v, relerr(got["dtheta"][v], ref["dtheta"][v]))
        assert relerr(got["dlogdelta"][v], ref["dlogdelta"][v]) < TOL, (v, relerr(got["dlogdelta"][v], ref["dlogdelta"][v]))
    assert relerr(got["dmu"], ref["dmu"]) < TOL and relerr(got["dlogsigma"], ref["dlogsigma"]) < TOL
    assert relerr(got["dY"], ref["dY"]) < TOL and relerr(got["dX"], ref["dX"]) < TOL


@pytest.mark.parametrize("M,N,K", [(2000, 3000, 64), (4000, 6000, 128)])
def test_auto_eligible_sizes_meet_the_gradient_tolerance(M, N, K):
    """north_star: loss and gradients within 1e-4 of the reference's FP32 fit.  At the smallest sizes PMF_KERNEL_AUTO
    hands to the tensor-core kernels (M N >= 6e6 (K/64)^2, min(M, N) >= 2000) every gradient is within 1e-4 of the
    ORACLE; one notch below the threshold AUTO stays on the FP32 kernel."""
    nb, nn = N // 6, N // 3
    views = {"mutation": ("bernoulli", nb), "methylation": ("normal", nn), "mrnaseq": ("normal", nn), "counts": ("poisson", N - nb - 2 * nn)}
    model, om, D = make_pair(M, views, K=K, seed=440 + K, missing=0.3, lambda_X_l2=1.0)
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_AUTO, 0)
        got = eng.loss_grad(include_reg=True)
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        ffma = eng.loss_grad(include_reg=True)
    finally:
        eng.close()
    ref = O.total_loss_grads(om, D)
    assert abs(got["loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    for k in ("dX", "dY", "dmu", "dlogsigma"):
        assert relerr(got[k], ref[k]) < TOL, (k, relerr(got[k], ref[k]))
    assert relerr(got["dY"], ffma["dY"]) > 1e-5          # AUTO did run the TF32 contractions here ...
    small, _, _ = make_pair(1990, views, K=K, seed=441 + K, missing=0.3, lambda_X_l2=1.0)
    eng = P.Engine(small)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_AUTO, 0)
        a = eng.loss_grad(include_reg=False)
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        b = eng.loss_grad(include_reg=False)
    finally:
        eng.close()
    assert relerr(a["dY"], b["dY"]) < 2e-6               # ... and the FP32 kernel below the threshold
