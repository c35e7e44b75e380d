"""The stage functions above ``mf_fit!`` (src/fit.jl:82-123, 248-296, 376-500, 560-1018).

They contain no arithmetic of their own on M x N data: they install regularisers, freeze layers, pick learning
rates and call the hot loop (``mf_fit_adapt_lr``) and the passes around it 5-15 times.  In a Julia deployment
this logic stays in the reference and calls the C ABI; it is mirrored here so that a host
language without the reference has the same entry points (``fit``, ``basic_fit``, ...).

Every device-touching call goes through a small backend namespace (default: this package's CUDA-backed
functions) so that the orchestration itself can be executed on a CPU-only machine against a stand-in backend
(tests/oracle_backend.py drives it with the NumPy oracle).  The backend never computes anything here."""
from __future__ import annotations

import copy
import time
from types import SimpleNamespace
from typing import List, Optional

import numpy as np

from . import featureset_ard as _fsard
from . import fit as _fit
from .layers import BatchScale, BatchShift, freeze_layer, unfreeze_layer
from .postfit import (init_ordinal_thresholds, reorder_by_importance, reweight_eb, rotate_by_svd, whiten)
from .regularizers import (ARDRegularizer, FeatureSetARDReg, GroupRegularizer, L2Regularizer, SequenceReg, ZeroReg,
                           freeze_reg, unfreeze_reg)
from .util import ids_to_ind_mat

f32 = np.float32


def _device_link_col_sqerr(model):
    """(sum of squared link-space errors, number of finite entries) per column at the model's current parameters."""
    if model._engine is not None:
        model._engine.push_structure()
        model._engine.push_params()
        return model._engine.link_col_sqerr()
    eng = _fit.Engine(model, device=getattr(model, "_device", 0))
    try:
        return eng.link_col_sqerr()
    finally:
        eng.close()


def _device_batch_stats(model):
    """(finite-entry counts, squared link-space errors) per (batch, column) of every batched view."""
    if model._engine is not None:
        model._engine.push_structure()
        model._engine.push_params()
        return model._engine.batch_stats()
    eng = _fit.Engine(model, device=getattr(model, "_device", 0))
    try:
        return eng.batch_stats()
    finally:
        eng.close()


def _device_acquire(model) -> bool:
    """fit! runs on a device-resident model (gpu(model) in the reference's driver); returns whether this call made it so"""
    if model._engine is not None:
        return False
    _fit.gpu(model)
    return True


def _device_release(model, made: bool):
    if made:
        _fit.cpu(model)


DEVICE = SimpleNamespace(acquire=_device_acquire, release=_device_release, mf_fit_adapt_lr=_fit.mf_fit_adapt_lr, init_mu=_fit.init_mu, init_logsigma=_fit.init_logsigma,
                         reweight_col_losses=_fit.reweight_col_losses, theta_delta_em=_fit.theta_delta_em,
                         link_col_sqerr=_device_link_col_sqerr, batch_stats=_device_batch_stats,
                         update_A=_fsard.update_A)


def _note(history, name, **kw):
    """``history!`` (src/util.jl:607-622): a named, time-stamped entry."""
    if history is not None:
        history.append(dict(name=name, time=time.time(), **kw))


def _has_batch_layers(model) -> bool:
    return isinstance(model.matfac.col_transform.unwrapped(1), BatchScale)


# ---- parameter initialisation (src/fit.jl:105-123, 248-296) ------------------------------------------------

def init_theta(model, max_epochs=500, lr_theta=1.0, history=None, backend=DEVICE, **kwargs):
    """``init_theta!`` (src/fit.jl:105-122): everything but the batch shift frozen, then the hot loop."""
    kwargs.pop("lr", None)
    ct = model.matfac.col_transform
    freeze_layer(ct, [1, 2, 3])
    backend.mf_fit_adapt_lr(model, lr=lr_theta, update_col_layers=True, max_epochs=max_epochs, history=history, **kwargs)
    unfreeze_layer(ct, [1, 2, 3])
    _note(history, "init_theta")


def init_factors(model, lr=1.0, max_epochs=1000, init_factors_method="adagrad", history=None, backend=DEVICE, **kwargs):
    """``init_factors!`` (src/fit.jl:248-288), AdaGrad branch (the L-BFGS initialiser is outside this build)."""
    if init_factors_method == "lbfgs":
        raise NotImplementedError("init_factors_method=\"lbfgs\" (src/fit_lbfgs.jl) is outside the hot path built here")
    backend.mf_fit_adapt_lr(model, update_X=True, update_Y=True, lr=lr, min_lr=0.05, max_epochs=max_epochs,
                            history=history, **kwargs)
    _note(history, "init_factors")


def construct_minimal_regularizer(model):
    """``construct_minimal_regularizer`` (src/regularizers.jl:750-774): one L2 group per noise model, weighted by
    K, the mean sigma^2 of the group's columns and the density of its data.  ``MF.batched_column_nanvar`` lives in
    MatFac.jl (not in the tree); taken as the corrected variance over the finite entries, like the reference's own
    ``nanvar`` (src/util.jl:38-40)."""
    mf = model.matfac
    K, M = mf.X.shape
    nm = mf.noise_model
    sigma = np.exp(mf.col_transform.unwrapped(0).logsigma)
    weights = []
    for cr in nm.col_ranges:
        block = np.asarray(model.data[:, cr.start:cr.stop], dtype=f32)
        finite = np.isfinite(block)
        m_vec = finite.sum(axis=0).astype(f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            col_vars = np.nanvar(np.where(finite, block, np.nan), axis=0, ddof=1).astype(f32)
        col_vars = np.maximum(np.nan_to_num(col_vars, nan=0.0), f32(1.0 / M))
        w = f32(K) * np.mean(sigma[cr.start:cr.stop] ** 2) / (np.sum(col_vars * m_vec) / f32(M))
        weights.append(np.full(K, w, dtype=f32))
    labels = [n.dist for n in nm.noises]
    return GroupRegularizer(labels, K=K, group_idx=list(nm.col_ranges), group_weights=weights)


# ---- batch effects (src/fit.jl:376-500) ---------------------------------------------------------------------

def init_batch_effects(model, max_epochs=5000, lr_regress=0.25, lr_mu=0.1, lr_theta=1.0, batch_method="EM",
                       batch_em_rtol=1e-8, batch_em_max_iter=100, history=None, backend=DEVICE, **kwargs):
    """``init_batch_effects!`` (src/fit.jl:376-497).  A stand-in factorisation whose X is the sample-condition
    indicator regresses every column on the conditions; its batch shift is fitted by least squares, column and
    batch scales come from the residuals, and the EM / EB step shrinks both."""
    orig = model.matfac
    resident_device = None
    if model._engine is not None:                 # the stand-in has another K: it gets handles of its own
        resident_device = model._engine.device
        _fit.cpu(model)
    try:
        sur = copy.deepcopy(orig)
        cond = ids_to_ind_mat(list(model.sample_conditions)).astype(f32)           # M x K_conditions
        M, k_cond = cond.shape
        N = orig.Y.shape[1]
        sur.X = np.asfortranarray(cond.T.copy())
        sur.Y = np.zeros((k_cond, N), dtype=f32, order="F")
        # X is data here and K is the number of conditions: the original factor penalties (K-sized weight vectors) do
        # not apply to the stand-in.  The reference drops Y_reg after init_mu! (whose M-estimation ignores penalties).
        sur.X_reg = ZeroReg()
        sur.Y_reg = ZeroReg()
        model.matfac = sur

        backend.init_mu(model, lr_mu=lr_mu, max_epochs=max_epochs, history=history)
        orig.col_transform.unwrapped(2).mu[...] = sur.col_transform.unwrapped(2).mu

        # regress every column on the conditions (no penalty)
        backend.mf_fit_adapt_lr(model, max_epochs=max_epochs, lr=lr_regress, min_lr=0.05, update_X=False, update_Y=True,
                                update_col_layers=False, history=history)
        _note(history, "regress_against_sample_conditions")

        sur.col_transform_reg = SequenceReg([ZeroReg(), ZeroReg(), ZeroReg(), ZeroReg()])   # least-squares batch shift
        init_theta(model, max_epochs=max_epochs, lr_theta=lr_theta, history=history, backend=backend)
        theta = sur.col_transform.unwrapped(3).theta
        for v in theta.values:
            v[~np.isfinite(v)] = 0

        sq, cnt = backend.link_col_sqerr(model)                                      # column scales
        with np.errstate(divide="ignore", invalid="ignore"):
            col_vars = (np.asarray(sq, dtype=f32) / np.asarray(cnt, dtype=f32)).astype(f32)
        col_vars[~np.isfinite(col_vars)] = 1

        ba_cnt, ba_sq = backend.batch_stats(model)                                   # batch scales
        delta2 = []
        for s, c, cr in zip(ba_sq, ba_cnt, theta.col_ranges):
            with np.errstate(divide="ignore", invalid="ignore"):
                v = (np.asarray(s, dtype=f32) / np.asarray(c, dtype=f32)).astype(f32)
            v[~np.isfinite(v)] = 1
            delta2.append(v / col_vars[None, cr.start:cr.stop])

        theta_values = theta.values
        if batch_method in ("EM", "EB"):
            res = backend.theta_delta_em(model, delta2, col_vars, update_priors=batch_method == "EM",
                                         batch_em_max_iter=batch_em_max_iter, batch_em_rtol=batch_em_rtol)
            theta_values, delta2 = res[0], res[1]              # (theta, delta^2[, relative changes per iteration])
    finally:
        model.matfac = orig
    ct = orig.col_transform
    with np.errstate(divide="ignore"):
        ct.unwrapped(0).logsigma[...] = np.log(np.sqrt(col_vars))
        for dst, d2 in zip(ct.unwrapped(1).logdelta.values, delta2):
            dst[...] = np.log(np.sqrt(d2))
    for dst, tv in zip(ct.unwrapped(3).theta.values, theta_values):
        dst[...] = tv
    if resident_device is not None:
        _fit.gpu(model, device=resident_device)


# ---- procedures (src/fit.jl:560-905) -------------------------------------------------------------------------

def basic_fit(model, fit_batch=False, batch_method="EM", fit_mu=False, fit_logsigma=False, reweight_losses=False,
              init_factors_=False, init_factors_method="adagrad", fit_factors=False, init_ordinal=False,
              svd_rotate=False, whiten_=False, lr=1.0, max_epochs=1000, history=None, lr_regress=1.0, lr_mu=0.1,
              lr_theta=1.0, backend=DEVICE, **kwargs):
    """``basic_fit!`` (src/fit.jl:563-672).  ``init_factors_`` / ``whiten_`` are the reference's ``init_factors`` /
    ``whiten`` keywords (renamed: they would shadow the functions of the same name)."""
    if init_ordinal:
        init_ordinal_thresholds(model)
    if fit_batch:
        ct = model.matfac.col_transform
        assert isinstance(ct.unwrapped(1), BatchScale), "Model must have batch parameters whenever `fit_batch` is true"
        assert isinstance(ct.unwrapped(3), BatchShift), "Model must have batch parameters whenever `fit_batch` is true"
        init_batch_effects(model, batch_method=batch_method, max_epochs=max_epochs, history=history,
                           lr_regress=lr_regress, lr_theta=lr_theta, backend=backend)
    else:
        if fit_mu:
            backend.init_mu(model, lr_mu=lr_mu, max_epochs=500, history=history)
        if fit_logsigma:
            backend.init_logsigma(model)
    if reweight_losses:
        backend.reweight_col_losses(model)
    if init_factors_:
        init_factors(model, lr=lr, init_factors_method=init_factors_method, history=history, max_epochs=max_epochs,
                     backend=backend, **kwargs)
    if fit_factors:
        backend.mf_fit_adapt_lr(model, update_X=True, update_Y=True, lr=lr, min_lr=0.05, max_epochs=max_epochs,
                                history=history, **kwargs)
    if whiten_:
        whiten(model)
    if svd_rotate:
        rotate_by_svd(model)
    unfreeze_layer(model.matfac.col_transform, [1, 2, 3, 4])


def basic_fit_reg_weight_eb(model, lr=1.0, max_epochs=1000, history=None, svd_rotate=True, backend=DEVICE, **kwargs):
    """``basic_fit_reg_weight_eb!`` (src/fit.jl:675-727): pre-fit under minimal penalties, set the regulariser
    weights by empirical Bayes from that fit, re-fit under the full penalties."""
    mf = model.matfac
    K = mf.X.shape[0]
    freeze_reg(mf.col_transform_reg, [1, 2, 3, 4])
    orig_X_reg, orig_Y_reg = mf.X_reg, mf.Y_reg
    mf.X_reg = L2Regularizer(K, 1.0)                           # X -> 0.5 * sum(X .* X)
    mf.Y_reg = construct_minimal_regularizer(model)
    basic_fit(model, fit_mu=True, fit_logsigma=True, reweight_losses=True, fit_batch=_has_batch_layers(model),
              init_factors_=True, svd_rotate=svd_rotate, whiten_=False, lr=lr, max_epochs=max_epochs, history=history,
              backend=backend, **kwargs)
    unfreeze_reg(mf.col_transform_reg, [1, 2, 3, 4])
    mf.X_reg, mf.Y_reg = orig_X_reg, orig_Y_reg
    reweight_eb(mf.col_transform_reg, mf.col_transform)
    reweight_eb(mf.X_reg, mf.X)
    reweight_eb(mf.Y_reg, mf.Y)
    _note(history, "reweight_eb")
    basic_fit(model, reweight_losses=True, fit_factors=True, lr=lr, max_epochs=max_epochs, history=history,
              backend=backend, **kwargs)


def fit_non_ard(model, fit_reg_weight="EB", lambda_max=None, n_lambda=8, lambda_min_frac=1e-3, backend=DEVICE, **kwargs):
    """``fit_non_ard!`` (src/fit.jl:730-747)."""
    if fit_reg_weight == "EB":
        basic_fit_reg_weight_eb(model, backend=backend, **kwargs)
    else:
        basic_fit(model, fit_mu=True, fit_logsigma=True, reweight_losses=True, fit_batch=_has_batch_layers(model),
                  fit_factors=True, backend=backend, **kwargs)


def fit_ard(model, max_epochs=1000, history=None, lr=1.0, lr_regress=1.0, lr_theta=1.0, svd_rotate=True,
            batch_method="EM", backend=DEVICE, **kwargs):
    """``fit_ard!`` (src/fit.jl:753-808): pre-fit under minimal penalties, then put the ARD prior back and
    continue."""
    mf = model.matfac
    orig_X_reg, orig_ard = mf.X_reg, mf.Y_reg
    mf.X_reg = ZeroReg()                                        # X -> 0
    mf.Y_reg = construct_minimal_regularizer(model)
    basic_fit(model, fit_batch=_has_batch_layers(model), fit_mu=True, fit_logsigma=True, init_factors_=True,
              reweight_losses=True, svd_rotate=svd_rotate, whiten_=True, lr_regress=lr_regress, lr_theta=lr_theta,
              batch_method=batch_method, max_epochs=max_epochs, history=history, backend=backend, lr=lr, **kwargs)
    mf.X_reg = orig_X_reg
    reweight_eb(mf.X_reg, mf.X)
    mf.Y_reg = orig_ard
    reweight_eb(mf.Y_reg, mf.Y)
    backend.reweight_col_losses(model)
    backend.mf_fit_adapt_lr(model, update_X=True, update_Y=True, lr=lr, min_lr=0.01, max_epochs=max_epochs,
                            history=history, **kwargs)


def fit_feature_set_ard(model, lr=1.0, max_epochs=1000, fsard_max_iter=10, fsard_max_A_iter=1000, fsard_term_rtol=1e-5,
                        svd_rotate=True, history=None, backend=DEVICE, **kwargs):
    """``fit_feature_set_ard!`` (src/fit.jl:814-895): vanilla-ARD pre-fit, then alternate ``update_A!`` with re-fits
    of the factors until beta stops moving."""
    mf = model.matfac
    orig_reg = mf.Y_reg
    mf.Y_reg = ARDRegularizer(model.feature_views)
    fit_ard(model, max_epochs=max_epochs, lr=lr, history=history, svd_rotate=svd_rotate, backend=backend, **kwargs)
    mf.Y_reg = orig_reg
    beta_old = orig_reg.beta.copy()
    beta_diff = None
    for it in range(1, fsard_max_iter + 1):
        backend.update_A(orig_reg, model, max_epochs=fsard_max_A_iter, term_iter=50)
        beta_diff = float(np.sum((beta_old - orig_reg.beta) ** 2) / np.sum(orig_reg.beta * orig_reg.beta))
        if beta_diff < fsard_term_rtol or it == fsard_max_iter:
            break
        beta_old = orig_reg.beta.copy()
        backend.mf_fit_adapt_lr(model, update_X=True, update_Y=True, lr=lr, min_lr=0.01, max_epochs=max_epochs,
                                history=history)
    return beta_diff


def fit(model, lr=1.0, fit_reg_weight="EB", n_lambda=8, lambda_max=None, lambda_min_frac=1e-3, keep_history=False,
        svd_rotate=True, fit_joint=False, fsard_max_iter=10, fsard_max_A_iter=1000, fsard_term_rtol=1e-5, rel_tol=1e-5,
        abs_tol=1e-5, max_epochs=1000, backend=DEVICE, **kwargs) -> Optional[List[dict]]:
    """``fit!`` (src/fit.jl:923-1018): the master procedure.  Returns the history list when ``keep_history``."""
    hist: Optional[List[dict]] = [] if keep_history else None
    _note(hist, "start")
    made_resident = backend.acquire(model)
    try:
        mf = model.matfac
        common = dict(history=hist, rel_tol=rel_tol, abs_tol=abs_tol, svd_rotate=svd_rotate, max_epochs=max_epochs,
                      backend=backend, **kwargs)
        # `lr` reaches only the feature-set branch and the joint adjustment: fit! names it as a keyword of its own
        # and does not hand it to fit_ard! / fit_non_ard!, which therefore run at their default of 1.0 (:962-1001)
        if isinstance(mf.Y_reg, ARDRegularizer):
            fit_ard(model, **common)
        elif isinstance(mf.Y_reg, FeatureSetARDReg):
            fit_feature_set_ard(model, lr=lr, fsard_max_iter=fsard_max_iter, fsard_max_A_iter=fsard_max_A_iter,
                                fsard_term_rtol=fsard_term_rtol, **common)
        else:
            fit_non_ard(model, fit_reg_weight=fit_reg_weight, lambda_max=lambda_max, n_lambda=n_lambda,
                        lambda_min_frac=lambda_min_frac, **common)
        if fit_joint:                                               # let the fitted parameters share information
            ct = mf.col_transform
            freeze_layer(ct, [1, 2, 3])
            unfreeze_layer(ct, 4)
            backend.mf_fit_adapt_lr(model, update_X=True, update_Y=True, lr=lr, update_col_layers=True,
                                    max_epochs=max_epochs, history=hist)
            unfreeze_layer(ct, [1, 2, 3])
        whiten(model)
        backend.reweight_col_losses(model)                          # matters for `transform`ing new samples
        reorder_by_importance(model)
        _note(hist, "reorder_factors")
        _note(hist, "finish")
    finally:
        backend.release(model, made_resident)
    return hist
