"""Host-side mirror of the reference's column / batch layers and the ``BatchArray``
container (src/batch_array.jl, src/layers.jl).  These objects only *hold* parameters
and bookkeeping; the arithmetic (forward, pullbacks) runs in the CUDA data pass.

BatchArray keeps, per batched view, the column range, the batch names in ``unique``
order, the int32 ordinal of every sample's batch (the dense form of the reference's
sparse Bool indicator ``row_batches[v]``) and the ``n_b x N_v`` value matrix."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .util import ids_to_index, ids_to_ranges, subset_ranges, unique


class BatchArray:
    """src/batch_array.jl:5-15."""

    def __init__(self, col_ranges, col_range_ids, batch_index, row_batch_ids, values):
        self.col_ranges: List[range] = list(col_ranges)
        self.col_range_ids = list(col_range_ids)
        self.batch_index: List[np.ndarray] = list(batch_index)     # int32[M] per view
        self.row_batch_ids = list(row_batch_ids)
        self.values: List[np.ndarray] = list(values)               # float32 n_b x N_v

    @classmethod
    def from_dicts(cls, feature_views, row_batch_dict, value_dicts):
        """Constructor of src/batch_array.jl:47-79: views absent from the dict are skipped."""
        views = unique(feature_views)
        ranges = ids_to_ranges(feature_views)
        keep = [k for k, v in enumerate(views) if v in row_batch_dict]
        out = cls([], [], [], [], [])
        for k in keep:
            idx, names = ids_to_index(list(row_batch_dict[views[k]]))
            vals = np.zeros((len(names), len(ranges[k])), dtype=np.float32)
            for b, name in enumerate(names):
                vals[b, :] = value_dicts[k][name]
            out.col_ranges.append(ranges[k])
            out.col_range_ids.append(views[k])
            out.batch_index.append(idx)
            out.row_batch_ids.append(names)
            out.values.append(vals)
        return out

    @property
    def row_batches(self):
        """The reference's indicator matrices (dense Bool M x n_b)."""
        mats = []
        for idx, names in zip(self.batch_index, self.row_batch_ids):
            m = np.zeros((len(idx), len(names)), dtype=bool)
            m[np.arange(len(idx)), idx] = True
            mats.append(m)
        return mats

    def zero(self):
        return BatchArray(self.col_ranges, self.col_range_ids, self.batch_index, self.row_batch_ids,
                          [np.zeros_like(v) for v in self.values])

    def view(self, rows, cols: range):
        """src/batch_array.jl:83-106 (values are numpy views, like the reference)."""
        kept, lo, hi = subset_ranges(self.col_ranges, cols)
        if lo > hi:
            return BatchArray([], [], [], [], [])
        rows = np.arange(rows.start, rows.stop) if isinstance(rows, range) else np.asarray(rows)
        vals = [v[:, k.start - r.start:k.stop - r.start]
                for k, r, v in zip(kept, self.col_ranges[lo:hi + 1], self.values[lo:hi + 1])]
        return BatchArray([range(k.start - cols.start, k.stop - cols.start) for k in kept],
                          self.col_range_ids[lo:hi + 1], [b[rows] for b in self.batch_index[lo:hi + 1]],
                          self.row_batch_ids[lo:hi + 1], vals)


class ColScale:
    """Z .* exp.(logsigma)'   (src/layers.jl:9-48)."""

    def __init__(self, N: int):
        self.logsigma = np.zeros(N, dtype=np.float32)


class ColShift:
    """Z .+ mu'   (src/layers.jl:53-90); the reference initialises mu ~ 1e-4 randn."""

    def __init__(self, N: int, rng=None):
        rng = np.random.default_rng(0) if rng is None else rng
        self.mu = (rng.standard_normal(N) * 1e-4).astype(np.float32)


def _zero_batch_array(col_batches, batch_dict):
    views = unique(col_batches)
    ranges = ids_to_ranges(col_batches)
    vds = []
    for v, cr in zip(views, ranges):
        vds.append({rb: np.zeros(len(cr), np.float32) for rb in unique(batch_dict[v])} if v in batch_dict else {})
    return BatchArray.from_dicts(col_batches, batch_dict, vds)


class BatchScale:
    """Z * exp(logdelta)   (src/layers.jl:95-152)."""

    def __init__(self, col_batches, batch_dict):
        self.logdelta = _zero_batch_array(col_batches, batch_dict)


class BatchShift:
    """Z + theta   (src/layers.jl:158-214)."""

    def __init__(self, col_batches, batch_dict):
        self.theta = _zero_batch_array(col_batches, batch_dict)


class Identity:
    """The ``x->x`` closure in slots 2 / 4 when there is no batch_dict (layers.jl:243)."""


class FrozenLayer:
    """src/layers.jl:299-334: parameters get no gradient, the pullback still flows."""

    def __init__(self, layer):
        self.layer = layer


class ViewableComposition:
    """Fixed order (ColScale, BatchScale|identity, ColShift, BatchShift|identity)
    (src/layers.jl:221-253)."""

    def __init__(self, layers):
        self.layers = tuple(layers)

    def unwrapped(self, idx):
        l = self.layers[idx]
        return l.layer if isinstance(l, FrozenLayer) else l


def construct_model_layers(feature_views, batch_dict) -> ViewableComposition:
    """src/layers.jl:240-253."""
    N = len(feature_views)
    layers = [ColScale(N), Identity(), ColShift(N), Identity()]
    if batch_dict is not None:
        layers[1] = BatchScale(feature_views, batch_dict)
        layers[3] = BatchShift(feature_views, batch_dict)
    return ViewableComposition(layers)


def _as_list(idx):
    return [idx] if isinstance(idx, int) else list(idx)


def freeze_layer(vc: ViewableComposition, idx):
    """freeze_layer! (src/layers.jl:337-349); ``idx`` is 1-based like the reference."""
    ls = list(vc.layers)
    for i in _as_list(idx):
        if not isinstance(ls[i - 1], (FrozenLayer, Identity)):
            ls[i - 1] = FrozenLayer(ls[i - 1])
    vc.layers = tuple(ls)


def unfreeze_layer(vc: ViewableComposition, idx):
    """unfreeze_layer! (src/layers.jl:351-363)."""
    ls = list(vc.layers)
    for i in _as_list(idx):
        if isinstance(ls[i - 1], FrozenLayer):
            ls[i - 1] = ls[i - 1].layer
    vc.layers = tuple(ls)
