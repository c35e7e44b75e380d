"""Host-side mirror of the reference's regulariser objects (src/regularizers.jl,
src/featureset_ard.jl).  They hold weights and sparse bookkeeping and know how to
install themselves on a libpmf handle; value / pullback evaluation happens in CUDA."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import scipy.sparse as sp

from .layers import BatchArray, FrozenLayer
from .util import (compute_nongraph_nodes, csc_select, edgelist_to_spmat, featuresets_to_csc,
                   get_all_nodes, ids_to_ranges, unique, value_to_idx)


class ZeroReg:
    """``x -> 0`` (src/regularizers.jl:674,715; src/fit.jl:415,769)."""


class L2Regularizer:
    """src/regularizers.jl:11-55."""

    def __init__(self, K, w):
        self.weights = np.full(K, w, dtype=np.float32)


class L1Regularizer:
    """src/regularizers.jl:61-100 (not reachable from the model constructor)."""

    def __init__(self, K, w):
        self.weights = np.full(K, w, dtype=np.float32)


class GroupRegularizer:
    """src/regularizers.jl:345-456."""

    def __init__(self, group_labels, weight=1.0, K=1, group_idx=None, group_weights=None):
        self.group_labels = unique(group_labels)
        self.group_idx = ids_to_ranges(group_labels) if group_idx is None else list(group_idx)
        if group_weights is None:
            group_weights = [np.full(K, weight, dtype=np.float32) for _ in self.group_idx]
        self.group_weights = [np.asarray(w, dtype=np.float32) for w in group_weights]


class SelectiveL1Reg:
    """src/regularizers.jl:106-163: l1_idx[k, j] = feature j is absent from graph k."""

    def __init__(self, feature_ids, edgelists, weight=1.0):
        outside = compute_nongraph_nodes(feature_ids, edgelists)
        self.l1_idx = np.array([[f in s for f in feature_ids] for s in outside], dtype=bool)
        self.weight = np.full(len(outside), weight, dtype=np.float32)


class NetworkRegularizer:
    """src/regularizers.jl:169-240: per factor the Laplacian blocks AA (observed x observed),
    AB (observed x virtual), BB (virtual x virtual); virtual nodes sorted."""

    def __init__(self, feature_ids, edgelists, epsilon=0.1, weight=1.0):
        feature_ids = list(feature_ids)
        n = len(feature_ids)
        observed = set(feature_ids)
        self.AA, self.AB, self.BB, self.x_virtual = [], [], [], []
        for el in edgelists:
            virtual = sorted(get_all_nodes(el) - observed)
            lap = edgelist_to_spmat(el, value_to_idx(feature_ids + virtual), epsilon=epsilon) * float(weight)
            nt = lap.shape[0]
            self.AA.append(csc_select(lap, range(0, n), range(0, n)))
            self.AB.append(csc_select(lap, range(0, n), range(n, nt)))
            self.BB.append(csc_select(lap, range(n, nt), range(n, nt)))
            self.x_virtual.append(np.zeros(nt - n, dtype=np.float32))
        self.cur_weights = np.full(len(edgelists), weight, dtype=np.float32)
        self.cg_rtol = 0.0   # <= 0: Krylov.jl default sqrt(eps(Float32))
        self.cg_atol = 0.0
        self.cg_itmax = 0


class ARDRegularizer:
    """src/regularizers.jl:526-609."""

    def __init__(self, column_groups, alpha=np.float32(1.001), beta=np.float32(0.001), weight=1.0):
        self.col_ranges = ids_to_ranges(column_groups)
        self.alpha = [np.float32(alpha)] * len(self.col_ranges)
        self.beta = [np.float32(beta)] * len(self.col_ranges)
        self.weight = weight


class ISTAOptimiser:
    """src/optimizers.jl:26-44 (state only; the update rule runs on the device)."""

    def __init__(self, target, lr, l1_lambda):
        self.lr = np.float32(lr)
        self.ssq_grad = np.zeros_like(target) + np.float32(1e-8)
        self.lam = np.asarray(l1_lambda, dtype=np.float32)


class FeatureSetARDReg:
    """src/featureset_ard.jl:19-65."""

    def __init__(self, K, feature_views, S_vec, featureset_ids_vec, alpha0=1.01, v0=0.8, lr=0.05):
        N = len(feature_views)
        self.col_ranges = ids_to_ranges(feature_views)
        for cr, S, fid in zip(self.col_ranges, S_vec, featureset_ids_vec):
            assert len(cr) == S.shape[1], "columns of each feature_view must match size(S, 2)"
            assert len(fid) == S.shape[0], "featureset_ids incompatible with size(S, 1)"
        self.S = [sp.csr_matrix(S, dtype=np.float32) for S in S_vec]
        self.A = [np.zeros((S.shape[0], K), dtype=np.float32) for S in self.S]
        self.alpha0 = np.float32(alpha0)
        self.v0 = np.float32(v0)
        self.featureset_ids = [list(f) for f in featureset_ids_vec]
        self.alpha = np.full(N, self.alpha0, dtype=np.float32)
        self.beta = np.full((K, N), self.alpha0 - np.float32(1), dtype=np.float32)
        self.A_opts = [ISTAOptimiser(A, lr, np.ones(K, np.float32)) for A in self.A]


def construct_featureset_ard(K, feature_ids, feature_views, feature_sets_dict, featureset_ids=None,
                             alpha0=np.float32(1.001), v0=np.float32(0.8), lr=np.float32(0.05)):
    """src/featureset_ard.jl:111-132."""
    feature_ids = list(feature_ids)
    S_vec, names = [], []
    for cr, uv in zip(ids_to_ranges(feature_views), unique(feature_views)):
        sets = feature_sets_dict[uv]
        S_vec.append(featuresets_to_csc(feature_ids[cr.start:cr.stop], sets))
        names.append(list(range(1, len(sets) + 1)) if featureset_ids is None else featureset_ids[uv])
    return FeatureSetARDReg(K, feature_views, S_vec, names, alpha0=alpha0, v0=v0, lr=lr)


class CompositeRegularizer:
    """src/regularizers.jl:616-649."""

    def __init__(self, regularizers, mixture_p):
        self.regularizers = tuple(regularizers)
        self.mixture_p = tuple(float(p) for p in mixture_p)


def construct_X_reg(K, M, sample_ids, sample_conditions, sample_graphs, lambda_X_l2,
                    lambda_X_condition, lambda_X_graph, Y_ard, Y_geneset_ard):
    """src/regularizers.jl:655-689."""
    if Y_ard or Y_geneset_ard:
        if sample_conditions is not None:
            return GroupRegularizer(sample_conditions, weight=1.0, K=K)
        return L2Regularizer(K, 1.0)
    slots = [ZeroReg(), ZeroReg(), ZeroReg()]
    p = np.zeros(3)
    if lambda_X_l2 is not None:
        slots[0], p[0] = L2Regularizer(K, lambda_X_l2), 1
    if sample_conditions is not None:
        slots[1], p[1] = GroupRegularizer(sample_conditions, weight=lambda_X_condition, K=K), 1
    if sample_graphs is not None:
        slots[2], p[2] = NetworkRegularizer(sample_ids, sample_graphs, weight=lambda_X_graph), 1
    total = p.sum()
    # the reference divides unguarded here (NaN mixture when nothing is enabled); an all-zero
    # mixture is the only sensible reading of "no regulariser", so guard like construct_Y_reg
    return CompositeRegularizer(slots, p / (total if total > 0 else 1))


def construct_Y_reg(K, N, feature_ids, feature_views, feature_sets_dict, feature_graphs,
                    lambda_Y_l2, lambda_Y_selective_l1, lambda_Y_graph, Y_ard, Y_geneset_ard,
                    featureset_names, alpha0, v0):
    """src/regularizers.jl:696-739."""
    if Y_geneset_ard:
        return construct_featureset_ard(K, feature_ids, feature_views, feature_sets_dict,
                                        featureset_ids=featureset_names, alpha0=alpha0, v0=v0)
    if Y_ard:
        return ARDRegularizer(feature_views)
    slots = [ZeroReg(), ZeroReg(), ZeroReg()]
    p = np.zeros(3)
    if lambda_Y_l2 is not None:
        slots[0], p[0] = GroupRegularizer(feature_views, K=K, weight=lambda_Y_l2), 1
    if feature_ids is not None and feature_graphs is not None:
        if lambda_Y_selective_l1 is not None:
            slots[1], p[1] = SelectiveL1Reg(feature_ids, feature_graphs, weight=lambda_Y_selective_l1), 1
        if lambda_Y_graph is not None:
            slots[2], p[2] = NetworkRegularizer(feature_ids, feature_graphs, weight=lambda_Y_graph), 1
    total = p.sum()
    return CompositeRegularizer(slots, p / (total if total > 0 else 1))


class ColParamReg:
    """src/regularizers.jl:462-519."""

    def __init__(self, feature_views, weight=1.0, center=0.0):
        self.col_ranges = ids_to_ranges(feature_views)
        self.weights = [np.float32(weight)] * len(self.col_ranges)
        self.centers = [np.float32(center)] * len(self.col_ranges)

    def expanded(self, N):
        w = np.zeros(N, np.float32)
        c = np.zeros(N, np.float32)
        for r, wi, ci in zip(self.col_ranges, self.weights, self.centers):
            w[r.start:r.stop] = wi
            c[r.start:r.stop] = ci
        return w, c


class BatchArrayReg:
    """src/regularizers.jl:781-792."""

    def __init__(self, ba: BatchArray, center=0.0, weight=1.0):
        self.centers = [np.full(v.shape[0], center, dtype=np.float32) for v in ba.values]
        self.weights = [np.full(v.shape[0], weight, dtype=np.float32) for v in ba.values]


class FrozenRegularizer:
    """src/regularizers.jl:950-959: evaluates to 0 while wrapped."""

    def __init__(self, reg):
        self.reg = reg


class SequenceReg:
    """src/regularizers.jl:896-906."""

    def __init__(self, regs):
        self.regs = tuple(regs)


def construct_layer_reg(feature_views, batch_dict, layers, lambda_layer) -> SequenceReg:
    """src/regularizers.jl:908-926."""
    regs = [ZeroReg(), ZeroReg(), ZeroReg(), ZeroReg()]
    if feature_views is not None:
        regs[0] = ColParamReg(feature_views, weight=lambda_layer)
        regs[2] = ColParamReg(feature_views, weight=lambda_layer)
    if batch_dict is not None:
        regs[1] = BatchArrayReg(layers.layers[1].logdelta, weight=lambda_layer)
        regs[3] = BatchArrayReg(layers.layers[3].theta, weight=lambda_layer)
    return SequenceReg(regs)


def _as_list(idx):
    return [idx] if isinstance(idx, int) else list(idx)


def freeze_reg(sr: SequenceReg, idx):
    """freeze_reg! (src/regularizers.jl:975-987), 1-based slots."""
    regs = list(sr.regs)
    for i in _as_list(idx):
        if not isinstance(regs[i - 1], FrozenRegularizer):
            regs[i - 1] = FrozenRegularizer(regs[i - 1])
    sr.regs = tuple(regs)


def unfreeze_reg(sr: SequenceReg, idx):
    """unfreeze_reg! (src/regularizers.jl:989-1001)."""
    regs = list(sr.regs)
    for i in _as_list(idx):
        if isinstance(regs[i - 1], FrozenRegularizer):
            regs[i - 1] = regs[i - 1].reg
    sr.regs = tuple(regs)
