"""A stand-in backend for ``pathmatfac_b200.staging``: every device-touching call of the stage functions is served by
the NumPy oracle on the host, so the orchestration (regulariser swapping, freezing, call order, bookkeeping between
calls) can be executed end to end on a machine without a GPU.  Test infrastructure only."""
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

from oracle import pmf_oracle as O
from pathmatfac_b200.layers import BatchScale, BatchShift, FrozenLayer


def _f64(v):
    if isinstance(v, np.ndarray):
        return v.astype(np.float64) if v.dtype.kind == "f" else v.copy()
    if sp.issparse(v):
        return sp.csc_matrix(v, dtype=np.float64)
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (list, tuple)):
        return [_f64(x) for x in v]
    if hasattr(v, "__dict__") and hasattr(O, type(v).__name__):
        return conv_reg(v)
    return v


def conv_reg(reg):
    """Mirror regulariser object -> the oracle's object of the same name (float64 copies of the same fields)."""
    name = type(reg).__name__
    if name in ("ZeroReg", "FrozenRegularizer"):        # a frozen regulariser evaluates to 0 (regularizers.jl:950-959)
        return O.ZeroReg()
    if name == "CompositeRegularizer":
        return O.CompositeRegularizer([conv_reg(r) for r in reg.regularizers], list(reg.mixture_p))
    new = object.__new__(getattr(O, name))
    for k, v in reg.__dict__.items():
        setattr(new, k, _f64(v))
    return new


def _batch_array(ba):
    return O.BatchArray(ba.col_ranges, ba.col_range_ids, ba.row_batches, ba.row_batch_ids,
                        [v.astype(np.float64) for v in ba.values])


def to_oracle(model):
    mf = model.matfac
    ct = mf.col_transform
    nm = O.NoiseModel.from_distributions(list(model.feature_distributions))
    nm.weights = mf.noise_model.weights().astype(np.float64)
    nm.thresholds = [None if n.ext_thresholds is None else n.ext_thresholds.astype(np.float64) for n in mf.noise_model.noises]
    l1, l3 = ct.unwrapped(1), ct.unwrapped(3)
    om = O.OracleModel(X=np.array(mf.X, dtype=np.float64), Y=np.array(mf.Y, dtype=np.float64),
                       logsigma=ct.unwrapped(0).logsigma.astype(np.float64), mu=ct.unwrapped(2).mu.astype(np.float64),
                       logdelta=_batch_array(l1.logdelta) if isinstance(l1, BatchScale) else None,
                       theta=_batch_array(l3.theta) if isinstance(l3, BatchShift) else None, noise=nm)
    om.X_reg, om.Y_reg = conv_reg(mf.X_reg), conv_reg(mf.Y_reg)
    om.layer_regs = [conv_reg(r) for r in mf.col_transform_reg.regs]
    om.frozen = [isinstance(l, FrozenLayer) for l in ct.layers]
    return om, np.asarray(model.data, dtype=np.float64)


def from_oracle(om, model):
    mf = model.matfac
    ct = mf.col_transform
    mf.X[...] = om.X
    mf.Y[...] = om.Y
    ct.unwrapped(0).logsigma[...] = om.logsigma
    ct.unwrapped(2).mu[...] = om.mu
    if om.logdelta is not None:
        for dst, src in zip(ct.unwrapped(1).logdelta.values, om.logdelta.values):
            dst[...] = src
        for dst, src in zip(ct.unwrapped(3).theta.values, om.theta.values):
            dst[...] = src
    mf.noise_model.set_weight(om.noise.weights.astype(np.float32))
    for reg, oreg in ((mf.X_reg, om.X_reg), (mf.Y_reg, om.Y_reg)):       # warm-start state of the graph penalty
        regs = zip(reg.regularizers, oreg.regularizers) if type(reg).__name__ == "CompositeRegularizer" else [(reg, oreg)]
        for r, o in regs:
            if type(r).__name__ == "NetworkRegularizer":
                for dst, src in zip(r.x_virtual, o.x_virtual):
                    dst[...] = src


CALLS = []          # (name, detail) trace of every backend call, for the orchestration tests


def _mf_fit_adapt_lr(model, lr=1.0, min_lr=0.001, max_epochs=1000, history=None, **kw):
    CALLS.append(("mf_fit_adapt_lr", dict(lr=lr, min_lr=min_lr, max_epochs=max_epochs,
                                          flags=tuple(k for k in ("update_X", "update_Y", "update_col_layers") if kw.get(k)),
                                          X_reg=type(model.matfac.X_reg).__name__, Y_reg=type(model.matfac.Y_reg).__name__,
                                          frozen=tuple(isinstance(l, FrozenLayer) for l in model.matfac.col_transform.layers))))
    om, D = to_oracle(model)
    keep = {k: kw[k] for k in ("update_X", "update_Y", "update_col_layers", "rel_tol", "abs_tol") if k in kw}
    hs = O.mf_fit_adapt_lr(om, D, lr=lr, min_lr=min_lr, max_epochs=max_epochs, **keep)
    from_oracle(om, model)
    if history is not None:
        history.extend(hs)
    return hs


def _init_mu(model, lr_mu=0.1, max_epochs=500, history=None, **kw):
    CALLS.append(("init_mu", dict(K=model.matfac.X.shape[0])))
    om, D = to_oracle(model)
    h = O.init_mu(om, D, lr_mu=lr_mu, max_epochs=max_epochs)
    model.matfac.col_transform.unwrapped(2).mu[...] = om.mu
    if history is not None:
        history.append(h)


def _init_logsigma(model):
    CALLS.append(("init_logsigma", {}))
    om, D = to_oracle(model)
    O.init_logsigma(om, D)
    model.matfac.col_transform.unwrapped(0).logsigma[...] = om.logsigma


def _reweight_col_losses(model):
    CALLS.append(("reweight_col_losses", {}))
    om, D = to_oracle(model)
    O.reweight_col_losses(om, D)
    model.matfac.noise_model.set_weight(om.noise.weights.astype(np.float32))


def _link_col_sqerr(model):
    CALLS.append(("link_col_sqerr", {}))
    om, D = to_oracle(model)
    return O.link_col_sqerr(om, D), np.isfinite(D).sum(axis=0).astype(np.float64)


def _batch_stats(model):
    CALLS.append(("batch_stats", {}))
    om, D = to_oracle(model)
    Z = O.forward(om)
    cnt = O.ba_map(lambda d: np.isfinite(d).astype(float), om.theta, D)
    sq = O.ba_map(lambda z, d: np.where(np.isfinite(d), (z - d) ** 2, 0.0), om.theta, Z, D)
    return cnt, sq


def _theta_delta_em(model, delta2, sigma2, update_priors=True, batch_em_max_iter=100, batch_em_rtol=1e-8):
    CALLS.append(("theta_delta_em", dict(update_priors=update_priors)))
    om, D = to_oracle(model)
    th, d2, _ = O.theta_delta_em(om, [np.asarray(d, float) for d in delta2], np.asarray(sigma2, float), D,
                                 update_priors=update_priors, batch_em_max_iter=batch_em_max_iter, batch_em_rtol=batch_em_rtol)
    return [t.astype(np.float32) for t in th], [d.astype(np.float32) for d in d2]


def _update_A(reg, model, max_epochs=1000, term_iter=20, **kw):
    CALLS.append(("update_A", {}))
    oreg = conv_reg(reg)
    out = O.update_A(oreg, np.array(model.matfac.Y, dtype=np.float64), max_epochs=max_epochs, term_iter=term_iter)
    reg.beta[...] = oreg.beta
    for dst, src in zip(reg.A, oreg.A):
        dst[...] = src
    return out


BACKEND = SimpleNamespace(acquire=lambda model: False, release=lambda model, made: None,
                          mf_fit_adapt_lr=_mf_fit_adapt_lr, init_mu=_init_mu, init_logsigma=_init_logsigma,
                          reweight_col_losses=_reweight_col_losses, theta_delta_em=_theta_delta_em,
                          link_col_sqerr=_link_col_sqerr, batch_stats=_batch_stats, update_A=_update_A)
