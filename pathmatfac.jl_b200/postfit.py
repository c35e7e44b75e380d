"""Host-side steps the reference runs BETWEEN hot-loop calls: empirical-Bayes
re-weighting of the regularisers (``reweight_eb!``), factor re-ordering (``reorder_reg!``), and the
``whiten!`` / ``rotate_by_svd!`` / ``reorder_by_importance!`` post-processing of a fitted model
(src/fit.jl:504-555).  All of it is K x N / K x M bookkeeping on host arrays; a device-resident model
is re-synchronised through ``Engine.push_params`` / ``push_structure`` afterwards."""
from __future__ import annotations

import numpy as np

from .layers import BatchArray, BatchScale, BatchShift, ColScale, ColShift, FrozenLayer, ViewableComposition
from .regularizers import (ARDRegularizer, BatchArrayReg, ColParamReg, CompositeRegularizer, FeatureSetARDReg,
                           FrozenRegularizer, GroupRegularizer, L1Regularizer, L2Regularizer, NetworkRegularizer,
                           SelectiveL1Reg, SequenceReg, ZeroReg)
from .util import ids_to_ranges

f32 = np.float32


def _top_sv2(X) -> np.float32:
    """largest squared singular value (``svd(X).S[1]^2``)"""
    X = np.atleast_2d(np.asarray(X, dtype=f32))
    return f32(np.linalg.svd(X, compute_uv=False)[0]) ** 2


def _row_var(X):
    return np.var(np.asarray(X, dtype=f32), axis=1, ddof=1)


def reweight_eb(reg, x, mixture_p=1.0):
    """``reweight_eb!`` -- one function, dispatching on the regulariser like the reference's methods."""
    p = f32(mixture_p)
    if isinstance(reg, ZeroReg):                                   # pure functions have no weights (:941)
        return
    if isinstance(reg, L2Regularizer):                             # src/regularizers.jl:39-51
        X = x if np.ndim(x) == 2 else np.asarray(x)[None, :]
        reg.weights[...] = p / _top_sv2(X)
    elif isinstance(reg, L1Regularizer):                           # :88-96
        X = x if np.ndim(x) == 2 else np.asarray(x)[None, :]
        reg.weights[...] = p / _row_var(X)
    elif isinstance(reg, SelectiveL1Reg):                          # :149-159
        sel = reg.l1_idx * np.asarray(x, dtype=f32)
        mean_x = sel.mean(axis=1)
        var_x = (sel * sel).mean(axis=1) - mean_x * mean_x
        with np.errstate(divide="ignore", invalid="ignore"):
            w = p * np.sqrt(f32(2) / var_x)
        w[~np.isfinite(w)] = 1
        reg.weight[...] = w
    elif isinstance(reg, NetworkRegularizer):                      # :313-328
        row_precs = p / _row_var(x)
        ratio = row_precs / reg.cur_weights
        for k in range(len(reg.AA)):
            for blocks in (reg.AA, reg.AB, reg.BB):
                blocks[k].data *= ratio[k]
        reg.cur_weights[...] = row_precs
    elif isinstance(reg, GroupRegularizer):                        # :406-420
        X = np.asarray(x, dtype=f32)
        K = X.shape[0]
        reg.group_weights = [np.full(K, p / _top_sv2(X[:, r.start:r.stop]), dtype=f32) for r in reg.group_idx]
    elif isinstance(reg, ColParamReg):                             # :490-519
        if isinstance(x, ColShift):
            x = x.mu
        elif isinstance(x, ColScale):
            x = x.logsigma
        elif isinstance(x, FrozenLayer):
            raise TypeError("reweight_eb!(::ColParamReg, ::FrozenLayer) has no method in the reference")
        v = np.asarray(x, dtype=f32)
        reg.centers = [f32(v[r.start:r.stop].mean()) for r in reg.col_ranges]
        new_vars = [f32(np.var(v[r.start:r.stop], ddof=1)) for r in reg.col_ranges]
        reg.weights = [p * f32(1e-1 + 0.5) / (f32(1e-1) + f32(0.5) * nv) for nv in new_vars]
    elif isinstance(reg, ARDRegularizer):                          # :588-609
        reg.alpha = [0.001] * len(reg.alpha)
        reg.beta = [0.001] * len(reg.beta)
    elif isinstance(reg, CompositeRegularizer):                    # :634-638 (super_mixture_p)
        for r, q in zip(reg.regularizers, reg.mixture_p):
            reweight_eb(r, x, mixture_p=f32(q) * p)
    elif isinstance(reg, BatchArrayReg):                           # :818-877
        if isinstance(x, BatchShift):
            x = x.theta
        elif isinstance(x, BatchScale):
            x = x.logdelta
        assert isinstance(x, BatchArray)
        reg.centers = [v.mean(axis=1).astype(f32) for v in x.values]
        with np.errstate(divide="ignore", invalid="ignore"):
            reg.weights = [(p / np.var(v, axis=1, ddof=1)).astype(f32) for v in x.values]
        for w, cr in zip(reg.weights, x.col_ranges):
            w[~np.isfinite(w)] = f32(1) + f32(0.5) * len(cr)       # posterior mean of a Gamma(1, 1) precision
    elif isinstance(reg, SequenceReg):                             # :928-932
        assert isinstance(x, ViewableComposition)
        for r, layer in zip(reg.regs, x.layers):
            reweight_eb(r, layer, mixture_p=mixture_p)
    else:
        raise TypeError(f"reweight_eb! has no method for {type(reg).__name__}")


def reorder_reg(reg, perm):
    """``reorder_reg!`` (src/regularizers.jl:5,53,98,161,330,449,645,1002; src/featureset_ard.jl:68): permute the
    per-factor state of a regulariser; ``perm`` 0-based."""
    perm = np.asarray(perm)
    if isinstance(reg, (L2Regularizer, L1Regularizer)):
        reg.weights[...] = reg.weights[perm]
    elif isinstance(reg, SelectiveL1Reg):
        reg.l1_idx[...] = reg.l1_idx[perm, :]                       # the weights stay where they are (:161-163)
    elif isinstance(reg, NetworkRegularizer):
        for name in ("AA", "AB", "BB", "x_virtual"):
            setattr(reg, name, [getattr(reg, name)[k] for k in perm])
        reg.cur_weights[...] = reg.cur_weights[perm]
    elif isinstance(reg, GroupRegularizer):
        reg.group_weights = [w[perm] for w in reg.group_weights]
    elif isinstance(reg, CompositeRegularizer):
        for r in reg.regularizers:
            reorder_reg(r, perm)
    elif isinstance(reg, FrozenRegularizer):
        reorder_reg(reg.reg, perm)
    elif isinstance(reg, FeatureSetARDReg):
        reg.beta[...] = reg.beta[perm, :]
        for A in reg.A:
            A[...] = A[:, perm]
        for opt in reg.A_opts:
            opt.ssq_grad[...] = opt.ssq_grad[:, perm]
            opt.lam[...] = opt.lam[perm]
    # every other regulariser: the generic no-op method (:5)


def _rms(X, axis):
    return np.sqrt(np.mean(X * X, axis=axis, keepdims=True))


def _resync(model):
    if model._engine is not None:
        model._engine.push_structure()
        model._engine.push_params()


def whiten(model):
    """``whiten!`` (src/fit.jl:504-528): rms(X_k) = 1 for every factor, the magnitude moved into Y and from there,
    per view, into logsigma (the largest row rms of the view's block of Y)."""
    mf = model.matfac
    x_rms = _rms(mf.X, 1)
    mf.X /= x_rms
    mf.Y *= x_rms
    logsigma = mf.col_transform.unwrapped(0).logsigma
    for cr in ids_to_ranges(list(model.feature_views)):
        y_max = _rms(mf.Y[:, cr.start:cr.stop], 1).max()
        if y_max > 0:
            mf.Y[:, cr.start:cr.stop] /= y_max
            logsigma[cr.start:cr.stop] += np.log(y_max)
        else:
            mf.Y[:, cr.start:cr.stop] = 0
            logsigma[cr.start:cr.stop] = f32(-1e9)
    _resync(model)


def rotate_by_svd(model):
    """``rotate_by_svd!`` (src/fit.jl:531-544): Y <- S V', X' <- X' U for Y = U S V'."""
    mf = model.matfac
    U, s, Vt = np.linalg.svd(mf.Y, full_matrices=False)
    mf.Y[...] = s[:, None] * Vt
    mf.X[...] = (mf.X.T @ U).T
    _resync(model)


def reorder_by_importance(model):
    """``reorder_by_importance!`` (src/fit.jl:547-555): factors sorted by decreasing sum_j Y_kj^2 (stable), the
    regularisers' per-factor state permuted with them."""
    mf = model.matfac
    y_ssq = np.sum(mf.Y * mf.Y, axis=1)
    idx = np.argsort(-y_ssq, kind="stable")
    mf.X[...] = mf.X[idx, :]
    mf.Y[...] = mf.Y[idx, :]
    reorder_reg(mf.Y_reg, idx)
    reorder_reg(mf.X_reg, idx)
    _resync(model)
    return idx


def inv_logistic(x):
    """src/util.jl:8-10 (the reference's damped logit)."""
    x = f32(x)
    return f32(np.log(f32(0.5) + f32(0.99) * (x / (f32(1) - x) - f32(0.5))))


def rec_set_thresholds(p_vec, l_p, r_p):
    """src/fit.jl:190-219.  The reference drops ALL remaining probabilities after the first step
    (``p_vec[2:1-end]`` is an empty range), so the result has two thresholds -- exactly what its only ordinal
    model, three levels, needs."""
    if len(p_vec) < 2:
        return []
    l_p = f32(l_p) + f32(p_vec[0])
    r_p = f32(r_p) + f32(p_vec[-1])
    return [inv_logistic(l_p)] + rec_set_thresholds([], l_p, r_p) + [inv_logistic(f32(1) - r_p)]


def init_ordinal_thresholds(model):
    """``init_ordinal_thresholds!`` (src/fit.jl:222-246): interior thresholds of every OrdinalNoise range from the
    add-one-smoothed level frequencies of its columns."""
    nm = model.matfac.noise_model
    for n, cr in zip(nm.noises, nm.col_ranges):
        if n.dist != "ordinal3":                       # isa(n, MF.OrdinalNoise); the squared-hinge variant is left alone
            continue
        levels = len(n.ext_thresholds) - 1
        block = model.data[:, cr.start:cr.stop]
        p = np.array([np.sum(block == f32(k)) + 1 for k in range(1, levels + 1)], dtype=f32)
        p /= p.sum()
        n.ext_thresholds[1:-1] = rec_set_thresholds(list(p), 0.0, 0.0)
    if model._engine is not None:
        model._engine.push_structure()

