"""Shared test scaffolding: build the same synthetic problem for the CPU oracle and for the
CUDA path (through the host mirror + C ABI)."""
import numpy as np

import pathmatfac_b200 as P
from oracle import pmf_oracle as O


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_pair(M, views, K, seed, batch_views=(), n_batches=0, n_conditions=0, missing=0.0,
              lambda_X_l2=None, lambda_Y_l2=1.0, lambda_layer=1.0, feature_graphs=None,
              lambda_Y_selective_l1=None, lambda_Y_graph=None, Y_ard=False, feature_sets=None,
              feature_ids=None, sort_batches=False):
    """Returns (product model, oracle model, D float32).  ``views``: name -> (dist, n cols),
    listed in (distribution, view) sorted order."""
    om, D, meta = O.simulate_model(M, views, K, seed, batch_views=batch_views, n_batches=n_batches,
                                   n_conditions=n_conditions, missing=missing, dtype=np.float64,
                                   sort_batches=sort_batches)
    D32 = np.asfortranarray(D.astype(np.float32))
    # round parameters to float32 so both paths start from identical numbers
    for name in ("X", "Y", "logsigma", "mu"):
        setattr(om, name, getattr(om, name).astype(np.float32).astype(np.float64))
    for ba in (om.logdelta, om.theta):
        if ba is not None:
            ba.values = [v.astype(np.float32).astype(np.float64) for v in ba.values]
    N = len(meta["views"])
    conditions = meta["conditions"]
    if batch_views and conditions is None:
        conditions = [0] * M
    fids = list(range(1, N + 1)) if feature_ids is None else feature_ids
    Y_fsard = feature_sets is not None
    model = P.PathMatFacModel(D32.copy(order="F"), K=K, sample_conditions=conditions, feature_ids=fids,
                              feature_views=meta["views"], feature_distributions=meta["dists"],
                              batch_dict=meta["batch_dict"], lambda_X_l2=lambda_X_l2, lambda_Y_l2=lambda_Y_l2,
                              lambda_layer=lambda_layer, feature_graphs=feature_graphs,
                              lambda_Y_selective_l1=lambda_Y_selective_l1, lambda_Y_graph=lambda_Y_graph,
                              Y_ard=Y_ard, Y_fsard=Y_fsard, feature_sets_dict=feature_sets)
    mf = model.matfac
    mf.X[...] = om.X
    mf.Y[...] = om.Y
    mf.col_transform.layers[0].logsigma[...] = om.logsigma
    mf.col_transform.layers[2].mu[...] = om.mu
    if om.logdelta is not None:
        for v in range(len(om.logdelta.values)):
            mf.col_transform.layers[1].logdelta.values[v][...] = om.logdelta.values[v]
            mf.col_transform.layers[3].theta.values[v][...] = om.theta.values[v]
    # oracle-side regularisers through the oracle's own constructors
    om.X_reg = O.construct_X_reg(K, M, list(range(1, M + 1)), conditions, None, lambda_X_l2, 1.0, 1.0, Y_ard, Y_fsard)
    om.Y_reg = O.construct_Y_reg(K, N, fids, meta["views"], feature_sets, feature_graphs, lambda_Y_l2,
                                 lambda_Y_selective_l1, lambda_Y_graph, Y_ard, Y_fsard, None,
                                 np.float32(1.001), np.float32(0.8))
    regs = [O.ColParamReg(meta["views"], weight=lambda_layer), O.ZeroReg(),
            O.ColParamReg(meta["views"], weight=lambda_layer), O.ZeroReg()]
    if om.logdelta is not None:
        regs[1] = O.BatchArrayReg(om.logdelta, weight=lambda_layer)
        regs[3] = O.BatchArrayReg(om.theta, weight=lambda_layer)
    om.layer_regs = regs
    return model, om, D.astype(np.float32).astype(np.float64)


def random_graphs(N, K, rng, n_virtual=6, n_edges=60):
    """Per-factor edge lists over features 1..N with a few virtual nodes (signed edges), as the model constructor
    takes them (``feature_graphs``)."""
    graphs = []
    for k in range(K):
        el = []
        for _ in range(n_edges):
            a, b = rng.integers(1, N + 1, size=2)
            if a != b:
                el.append([int(a), int(b), float(rng.choice([-1.0, 1.0]))])
        for v in range(n_virtual):
            for _ in range(3):
                el.append([int(rng.integers(1, N + 1)), f"virt{k}_{v}", 1.0])
        el.append([f"virt{k}_0", f"virt{k}_1", -1.0])
        graphs.append(el)
    return graphs
