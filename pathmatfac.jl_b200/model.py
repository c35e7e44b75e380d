"""``PathMatFacModel`` -- host-side mirror of the reference's model object and
constructor (src/model.jl:6-196), plus the minimal ``MatFacModel`` / ``CompositeNoise``
surface the reference uses from MatFac.jl."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .layers import construct_model_layers
from .regularizers import construct_layer_reg, construct_X_reg, construct_Y_reg
from .util import VALID_LOSSES, ids_to_ranges, is_contiguous, unique


class Noise:
    """One noise model (NormalNoise, BernoulliNoise, ... of MatFac.jl)."""

    def __init__(self, dist: str, n_cols: int):
        self.dist = dist
        self.weight = np.ones(n_cols, dtype=np.float32)
        # OrdinalNoise.ext_thresholds = [-Inf, t1, t2, +Inf]  (src/impute.jl:15-24)
        self.ext_thresholds = (np.array([-np.inf, -1.0, 1.0, np.inf], dtype=np.float32)
                               if dist.startswith("ordinal") else None)


class CompositeNoise:
    """``col_ranges`` / ``noises`` (src/fit.jl:227, src/regularizers.jl:756-757)."""

    def __init__(self, feature_distributions):
        self.col_ranges = ids_to_ranges(feature_distributions)
        self.noises = [Noise(d, len(r)) for d, r in zip(unique(feature_distributions), self.col_ranges)]

    def set_weight(self, w):
        """MF.set_weight!(noise_model, w) (src/fit.jl:157,180)."""
        w = np.asarray(w, dtype=np.float32)
        for r, n in zip(self.col_ranges, self.noises):
            n.weight[...] = w[r.start:r.stop]

    def weights(self) -> np.ndarray:
        return np.concatenate([n.weight for n in self.noises]).astype(np.float32)


class MatFacModel:
    """MatFacModel(M, N, K, feature_distributions; col_transform, X_reg, Y_reg,
    col_transform_reg) (src/model.jl:69-72).  X is K x M, Y is K x N."""

    def __init__(self, M, N, K, feature_distributions, col_transform, X_reg, Y_reg, col_transform_reg,
                 rng=None):
        rng = np.random.default_rng(0) if rng is None else rng
        # MatFac.jl's random init is external and unreproducible; callers that need parity set
        # X and Y explicitly
        # column-major like the reference's K x M / K x N Julia arrays: the device layout ([M][K], [N][K]) is then
        # the arrays' own memory and uploads / downloads need no transposed copy
        self.X = np.asfortranarray((rng.standard_normal((K, M)) * 0.01).astype(np.float32))
        self.Y = np.asfortranarray((rng.standard_normal((K, N)) * 0.01).astype(np.float32))
        self.col_transform = col_transform
        self.X_reg = X_reg
        self.Y_reg = Y_reg
        self.col_transform_reg = col_transform_reg
        self.noise_model = CompositeNoise(feature_distributions)


class PathMatFacModel:
    """src/model.jl:6-28 + the keyword constructor :92-196.  ``D`` is reordered in place so
    that (distribution, view) blocks are contiguous; ``data_idx`` is that permutation."""

    def __init__(self, D, K: int = 10, sample_ids=None, sample_conditions=None, feature_ids=None,
                 feature_views=None, feature_distributions=None, batch_dict=None, sample_graphs=None,
                 feature_sets_dict=None, featureset_names=None, feature_graphs=None,
                 lambda_X_l2=None, lambda_X_condition=1.0, lambda_X_graph=1.0, lambda_Y_l2=1.0,
                 lambda_Y_selective_l1=None, lambda_Y_graph=None, lambda_layer=1.0, Y_ard=False,
                 Y_fsard=False, fsard_alpha0=np.float32(1.001), fsard_v0=np.float32(0.8), rng=None):
        D = np.asarray(D)
        M, N = D.shape
        # -- validation, same messages as src/model.jl:118-186 ---------------------------------
        if feature_graphs is not None:
            K = len(feature_graphs)
            if sample_graphs is not None:
                assert K == len(sample_graphs), \
                    "`sample_graphs` and `feature_graphs` must have equal length; or one of them must be nothing"
        elif sample_graphs is not None:
            K = len(sample_graphs)
        if sample_ids is not None:
            assert len(sample_ids) == len(set(sample_ids)), "`sample_ids` must be unique"
            assert len(sample_ids) == M, "`sample_ids` must be nothing or have length equal to size(D,1)"
            sample_ids = list(sample_ids)
        else:
            sample_ids = list(range(1, M + 1))
        if sample_conditions is not None:
            assert len(sample_conditions) == M, \
                "`sample_conditions` must be nothing or have length equal to size(D,1)"
            assert is_contiguous(sample_conditions), \
                "`sample_conditions` must be contiguous; I.e., samples must be grouped by condition."
            sample_conditions = list(sample_conditions)
        if feature_ids is not None:
            assert len(feature_ids) == len(set(feature_ids)), \
                "`feature_ids` must be left default, or set to a vector of unique identifiers"
            assert len(feature_ids) == N, "`feature_ids` must have length equal to dim(D,2)"
            feature_ids = list(feature_ids)
        else:
            feature_ids = list(range(1, N + 1))
        if batch_dict is not None:
            assert feature_views is not None, "`feature_views` must be provided whenever `batch_dict` is provided"
            assert sample_conditions is not None, \
                "`sample_conditions` must be provided whenever `batch_dict` is provided"
            assert set(batch_dict) <= set(feature_views), "The `batch_dict` keys must be a subset of `feature_views`"
            for v in batch_dict.values():
                assert len(v) == M, "Each value of `batch_dict` must be a vector of length size(D,1)"
        if feature_views is not None:
            assert len(feature_views) == N, "`feature_views` must be nothing or have length equal to size(D,2)"
            feature_views = list(feature_views)
        else:
            feature_views = [1] * N
        if feature_distributions is not None:
            assert len(feature_distributions) == N, \
                "`feature_distributions` must (a) be nothing or have length equal to size(D,2)"
            assert all(d in VALID_LOSSES for d in feature_distributions), \
                f"Each entry of `feature_distributions` must be one of {set(VALID_LOSSES)}"
            feature_distributions = list(feature_distributions)
        else:
            feature_distributions = ["normal"] * N
        if Y_fsard:
            assert feature_sets_dict is not None, "`feature_sets_dict` must be provided whenever `Y_fsard` is true."

        # -- assemble_model (src/model.jl:37-81) ---------------------------------------------------
        order = sorted(range(N), key=lambda j: (feature_distributions[j], feature_views[j]))   # stable sortperm
        data_idx = np.asarray(order, dtype=np.int64)
        feature_ids = [feature_ids[j] for j in order]
        feature_views = [feature_views[j] for j in order]
        feature_distributions = [feature_distributions[j] for j in order]
        if not np.array_equal(data_idx, np.arange(N)):
            if isinstance(D, np.ndarray) and D.flags.writeable:
                D[...] = D[:, data_idx]      # the reference permutes the caller's matrix in place
            else:
                D = D[:, data_idx].copy()

        col_layers = construct_model_layers(feature_views, batch_dict)
        layer_reg = construct_layer_reg(feature_views, batch_dict, col_layers, lambda_layer)
        X_reg = construct_X_reg(K, M, sample_ids, sample_conditions, sample_graphs, lambda_X_l2,
                                lambda_X_condition, lambda_X_graph, Y_ard, Y_fsard)
        Y_reg = construct_Y_reg(K, N, feature_ids, feature_views, feature_sets_dict, feature_graphs,
                                lambda_Y_l2, lambda_Y_selective_l1, lambda_Y_graph, Y_ard, Y_fsard,
                                featureset_names, fsard_alpha0, fsard_v0)
        self.matfac = MatFacModel(M, N, K, feature_distributions, col_layers, X_reg, Y_reg, layer_reg, rng=rng)
        self.data = D
        self.sample_ids = sample_ids
        self.sample_conditions = sample_conditions
        self.feature_ids = feature_ids
        self.feature_views = feature_views
        self.feature_distributions = feature_distributions
        self.data_idx = data_idx
        self._engine = None     # device residency (the reference's gpu(model)); see fit.py
