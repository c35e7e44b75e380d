"""Prints (one JSON line per case) the device-vs-oracle differences behind the looser tolerances of the GPU suite:
the FSARD update_A trajectory after 20 and after 150 ISTA epochs, and the fitted parameters of the short fits.  The
tests assert with a margin over these numbers."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.featureset_ard import update_A
from oracle import pmf_oracle as O
from tests.helpers import make_pair, relerr


def fsard_case(max_epochs):
    N, K = 40, 6
    sets = {"methylation": [list(range(1, 6)), list(range(6, 11)), list(range(11, 16)), list(range(16, 21))],
            "mrnaseq": [list(range(21, 26)), list(range(25, 31)), list(range(31, 36)), list(range(36, 41))]}
    model, om, D = make_pair(50, {"methylation": ("normal", 20), "mrnaseq": ("normal", 20)}, K=K, seed=7, feature_sets=sets)
    rng = np.random.default_rng(3)
    beta = (0.001 + 0.2 * rng.random((K, N))).astype(np.float32)
    model.matfac.Y_reg.beta[...] = beta
    om.Y_reg.beta[...] = beta
    P.gpu(model)
    try:
        res = update_A(model.matfac.Y_reg, model, max_epochs=max_epochs, term_iter=20, atol=1e-5)
    finally:
        P.cpu(model)
    ref = O.update_A(om.Y_reg, om.Y.astype(np.float32), max_epochs=max_epochs, term_iter=20, atol=1e-5)
    out = {"case": f"update_A {max_epochs} epochs", "views": []}
    for (bl, ep), (rbl, rep), A, Ar in zip(res, ref, model.matfac.Y_reg.A, om.Y_reg.A):
        out["views"].append({"best_loss": bl, "best_loss_ref": rbl, "loss_rel": abs(bl - rbl) / max(abs(rbl), 1e-30),
                             "epochs": [int(ep), int(rep)], "A_relerr": relerr(A, Ar)})
    out["beta_relerr"] = relerr(model.matfac.Y_reg.beta, om.Y_reg.beta)
    return out


def fit_case(kernel, name, M, views, K, epochs, lr, **kw):
    model, om, D = make_pair(M, views, K=K, missing=0.25, lambda_X_l2=1.0, **kw)
    href = O.mf_fit(om, D, O.AdaGrad(lr), max_epochs=epochs, update_X=True, update_Y=True, update_col_layers=True,
                    rel_tol=0, abs_tol=0)
    h = P.mf_fit(model, lr=lr, max_epochs=epochs, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0,
                 kernel=kernel, verbosity=0)
    layers = model.matfac.col_transform.layers
    out = {"case": name, "loss_curve_max_rel": float(np.max(np.abs(np.array(h["loss"]) / np.array(href["loss"]) - 1))),
           "X": relerr(model.matfac.X, om.X), "Y": relerr(model.matfac.Y, om.Y), "mu": relerr(layers[2].mu, om.mu),
           "logsigma": relerr(layers[0].logsigma, om.logsigma)}
    if om.theta is not None:
        out["theta"] = [relerr(layers[3].theta.values[v], om.theta.values[v]) for v in range(len(om.theta.values))]
        out["logdelta"] = [relerr(layers[1].logdelta.values[v], om.logdelta.values[v]) for v in range(len(om.theta.values))]
    return out


if __name__ == "__main__":
    for ep in (20, 150):
        print(json.dumps(fsard_case(ep)), flush=True)
    small = {"mutation": ("bernoulli", 30), "methylation": ("normal", 60), "counts": ("poisson", 25)}
    print(json.dumps(fit_case(_lib.KERNEL_FFMA, "FP32 kernel, 120 x 115, K = 5, 25 epochs", 120, small, 5, 25, 0.3, seed=8,
                              batch_views=["methylation"], n_batches=4, n_conditions=3)), flush=True)
    mid = {"mutation": ("bernoulli", 600), "methylation": ("normal", 1400), "mrnaseq": ("normal", 1300), "counts": ("poisson", 800)}
    for kern, nm in ((_lib.KERNEL_TC, "tcgen05"), (_lib.KERNEL_FFMA, "FP32")):
        print(json.dumps(fit_case(kern, f"{nm} kernel, 1100 x 4100, K = 16, batch layers, 6 epochs", 1100, mid, 16, 6, 0.1, seed=73,
                                  batch_views=["methylation", "mrnaseq", "counts"], n_batches=6, n_conditions=4)), flush=True)
