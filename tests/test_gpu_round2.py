"""GPU parity tests added in round 2: boundary behaviour around repeated fits on a resident model, closures as
regularisers, frozen layers, and the options that make the MatFac.jl unknowns explicit."""
import numpy as np
import pytest

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from oracle import pmf_oracle as O
from tests.helpers import make_pair, relerr

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
