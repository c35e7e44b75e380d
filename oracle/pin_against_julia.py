"""Pinning tool for the "parity unpinned" part of the oracle (TEST INFRASTRUCTURE, like everything under oracle/).

The noise-model losses, the epoch order and the termination test of ``MF.fit!`` live in MatFac.jl, which is not in the
build image; every assumption the oracle makes about them is an option.  The day a Julia runtime
with the reference is available:

  1. ``pathmatfac_b200.simulate.export_problem(model, dir)``              (this repo; writes the inputs as flat binaries)
  2. ``julia julia/run_reference_fit.jl dir K lr epochs [ctor kwargs]``     (the UNMODIFIED reference; writes ref_*.bin)
  3. ``python -m oracle.pin_against_julia dir --K K --lr lr --epochs epochs [--ctor lambda_X_l2=1.0 ...]``

Step 3 rebuilds the oracle's model from the same files, runs ``oracle.mf_fit`` under every combination of the open
options (``alternating`` D1, ``update_noise_models`` D7) and reports, per combination, the largest relative difference
between its parameters and the reference's after the same number of epochs.  Exit status 0 = some combination agrees
within ``--tol`` (default 1e-4, north_star's bar): that combination is what MatFac.jl does, and the oracle's defaults and
``pmf_fit_opts`` defaults are to be set to it.  tests/test_oracle_golden.py exercises the tool end to end with a
stand-in for step 2."""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pmf_oracle as O  # noqa: E402


def read_export(directory):
    """manifest.txt + *.bin + the three text files of export_problem (column-major arrays)."""
    arrays = {}
    for line in open(os.path.join(directory, "manifest.txt")):
        name, ty, r, c = line.split()
        dt = np.float32 if ty == "Float32" else np.int32
        arrays[name] = np.fromfile(os.path.join(directory, name + ".bin"), dtype=dt).reshape((int(r), int(c)), order="F")

    def lines(f):
        return [l for l in open(os.path.join(directory, f)).read().split("\n") if l != ""]
    return arrays, lines("feature_views.txt"), lines("feature_distributions.txt"), lines("sample_conditions.txt")


def oracle_model_from_export(directory, K, lambda_X_l2=None, lambda_X_condition=1.0, lambda_Y_l2=1.0, lambda_layer=1.0):
    """The model the reference's constructor builds from the exported inputs (src/model.jl:92-196 defaults: condition
    group penalty on X when conditions are given, per-view L2 on Y, column / batch layer penalties)."""
    arrays, views, dists, conds = read_export(directory)
    M, N = arrays["data"].shape
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    batch_views = [k.split("__", 1)[1] for k in arrays if k.startswith("batch_of_sample__")]
    logdelta = theta = None
    batch_dict = None
    if batch_views:
        batch_dict = {v: [int(b) for b in arrays[f"batch_of_sample__{v}"][:, 0]] for v in batch_views}
        unq, ranges = O.unique_in_order(views), O.ids_to_ranges(views)

        def mk(prefix):
            vds = [({b: np.zeros(len(cr)) for b in O.unique_in_order(batch_dict[v])} if v in batch_dict else {})
                   for v, cr in zip(unq, ranges)]
            ba = O.BatchArray.construct(views, batch_dict, vds)
            ba.values = [f64(arrays[f"{prefix}__{name}"]) for name in ba.col_range_ids]
            return ba
        logdelta, theta = mk("logdelta"), mk("theta")
    nm = O.NoiseModel.from_distributions(dists)
    if "col_weights" in arrays:
        nm.weights = f64(arrays["col_weights"][:, 0])
    m = O.OracleModel(X=f64(arrays["X"]), Y=f64(arrays["Y"]), logsigma=f64(arrays["logsigma"][:, 0]),
                      mu=f64(arrays["mu"][:, 0]), logdelta=logdelta, theta=theta, noise=nm)
    conditions = conds if conds else None
    m.X_reg = O.construct_X_reg(K, M, list(range(1, M + 1)), conditions, None, lambda_X_l2, lambda_X_condition, 1.0, False, False)
    m.Y_reg = O.construct_Y_reg(K, N, list(range(1, N + 1)), views, None, None, lambda_Y_l2, None, None, False, False, None,
                                np.float32(1.001), np.float32(0.8))
    regs = [O.ColParamReg(views, weight=lambda_layer), O.ZeroReg(), O.ColParamReg(views, weight=lambda_layer), O.ZeroReg()]
    if logdelta is not None:
        regs[1], regs[3] = O.BatchArrayReg(logdelta, weight=lambda_layer), O.BatchArrayReg(theta, weight=lambda_layer)
    m.layer_regs = regs
    return m, f64(arrays["data"])


def read_reference_outputs(directory, m):
    def rd(name, shape):
        return np.fromfile(os.path.join(directory, name), dtype=np.float32).reshape(shape, order="F").astype(np.float64)
    ref = {"X": rd("ref_X.bin", m.X.shape), "Y": rd("ref_Y.bin", m.Y.shape),
           "logsigma": rd("ref_logsigma.bin", m.logsigma.shape), "mu": rd("ref_mu.bin", m.mu.shape)}
    if m.theta is not None:
        for v, name in enumerate(m.theta.col_range_ids):
            ref[f"theta/{name}"] = rd(f"ref_theta__{name}.bin", m.theta.values[v].shape)
            ref[f"logdelta/{name}"] = rd(f"ref_logdelta__{name}.bin", m.logdelta.values[v].shape)
    return ref


def oracle_outputs(m):
    out = {"X": m.X, "Y": m.Y, "logsigma": m.logsigma, "mu": m.mu}
    if m.theta is not None:
        for v, name in enumerate(m.theta.col_range_ids):
            out[f"theta/{name}"] = m.theta.values[v]
            out[f"logdelta/{name}"] = m.logdelta.values[v]
    return out


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


OPTIONS = {"alternating": (False, True), "update_noise_models": (False, True)}


def pin(directory, K, lr, epochs, tol=1e-4, **ctor):
    """[(options, {parameter: relative difference}, worst)] sorted by the worst difference, and the matching options
    (None when no combination is within ``tol``)."""
    rows = []
    for combo in itertools.product(*OPTIONS.values()):
        opts = dict(zip(OPTIONS, combo))
        m, D = oracle_model_from_export(directory, K, **ctor)
        if opts["update_noise_models"] and not any(d.startswith("ordinal") for d in m.noise.dists):
            continue                                     # no trainable noise parameter: identical to the run without it
        O.mf_fit(m, D, O.AdaGrad(lr), max_epochs=epochs, rel_tol=0.0, abs_tol=0.0, update_X=True, update_Y=True,
                 update_col_layers=True, **opts)
        ref = read_reference_outputs(directory, m)
        diffs = {k: relerr(v, ref[k]) for k, v in oracle_outputs(m).items()}
        rows.append((opts, diffs, max(diffs.values())))
    rows.sort(key=lambda r: r[2])
    match = rows[0][0] if rows and rows[0][2] < tol else None
    return rows, match


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("directory")
    ap.add_argument("--K", type=int, required=True)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--ctor", nargs="*", default=[], help="constructor keywords, e.g. lambda_X_l2=1.0")
    a = ap.parse_args(argv)
    ctor = {k: float(v) for k, v in (kv.split("=") for kv in a.ctor)}
    rows, match = pin(a.directory, a.K, a.lr, a.epochs, a.tol, **ctor)
    for opts, diffs, worst in rows:
        print(json.dumps({"options": opts, "worst": worst, "differences": diffs}))
    hist = os.path.join(a.directory, "ref_history.json")
    if os.path.exists(hist):
        h = json.load(open(hist)).get("history", {})
        print(json.dumps({"reference_history_keys": sorted(h) if isinstance(h, dict) else str(type(h)),
                          "term_code": h.get("term_code") if isinstance(h, dict) else None,
                          "epochs": h.get("epochs") if isinstance(h, dict) else None}))
    print(json.dumps({"match": match, "tol": a.tol}))
    return 0 if match is not None else 1


if __name__ == "__main__":
    sys.exit(main())
