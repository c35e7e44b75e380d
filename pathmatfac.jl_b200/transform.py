"""``transform`` (src/transform.jl:6-106): embed NEW samples with a fitted model -- clone the model
around the new data (NaN-padded to the training columns), drop the batch layers and every
regulariser, freeze the column layers and fit X only through the same boundary (``mf_fit_adapt_lr``,
i.e. the fused device pass with ``update_X`` alone)."""
from __future__ import annotations

import copy

import numpy as np

from .fit import mf_fit_adapt_lr
from .layers import Identity, freeze_layer, unfreeze_layer
from .regularizers import ZeroReg, freeze_reg


def keymatch(l_keys, r_keys):
    """src/util.jl:168-184: positions (0-based here) of the keys of ``l_keys`` found in ``r_keys``."""
    rkey_to_idx = {k: i for i, k in enumerate(r_keys)}
    l_idx, r_idx = [], []
    for i, lk in enumerate(l_keys):
        if lk in rkey_to_idx:
            l_idx.append(i)
            r_idx.append(rkey_to_idx[lk])
    return l_idx, r_idx


def set_layer(vc, idx, layer):
    """set_layer! (src/layers.jl:255-259); ``idx`` is 1-based like the reference."""
    ls = list(vc.layers)
    ls[idx - 1] = layer
    vc.layers = tuple(ls)


def transform(model, D, feature_ids=None, sample_ids=None, verbosity=1, print_prefix="", max_epochs=1000,
              lr=1.0, capacity=10 ** 8, **fit_kwargs):
    K, N = model.matfac.Y.shape
    D = np.asarray(D)
    M_new, N_new = D.shape
    # column attributes (src/transform.jl:20-34)
    old_idx, new_idx = list(range(N)), list(range(N_new))
    if feature_ids is None:
        assert N_new == N, ("Columns of D do not match columns of training data. "
                            "Provide `feature_ids` to ensure they match.")
    else:
        old_idx, new_idx = keymatch(model.feature_ids, list(feature_ids))
    if sample_ids is not None:
        assert len(sample_ids) == M_new, "`sample_ids` must have length == size(D,1)"
        sample_ids = list(sample_ids)
    else:
        sample_ids = list(range(1, M_new + 1))
    # a model around the new dataset (src/transform.jl:46-71)
    old_data, old_engine = model.data, model._engine
    model.data, model._engine = None, None
    try:
        new_model = copy.deepcopy(model)
    finally:
        model.data, model._engine = old_data, old_engine
    new_data = np.full((M_new, N), np.nan, dtype=np.float32, order="F")
    new_data[:, old_idx] = D[:, new_idx]
    new_model.data = new_data
    mf = new_model.matfac
    mf.Y_reg = ZeroReg()                                   # Y is not updated: drop its regulariser
    set_layer(mf.col_transform, 2, Identity())             # batch effects are ignored on new data
    set_layer(mf.col_transform, 4, Identity())
    mf.X = np.zeros((K, M_new), dtype=np.float32, order="F")
    mf.X_reg = ZeroReg()
    new_model.sample_ids = sample_ids
    new_model.sample_conditions = None
    # freeze the column layers and their regularisers, fit X only (src/transform.jl:82-90)
    freeze_layer(mf.col_transform, [1, 2, 3, 4])
    freeze_reg(mf.col_transform_reg, [1, 2, 3, 4])
    mf_fit_adapt_lr(new_model, update_X=True, verbosity=verbosity, print_prefix="    " + print_prefix,
                    max_epochs=max_epochs, lr=lr, capacity=capacity, **fit_kwargs)
    unfreeze_layer(mf.col_transform, [1, 2, 3, 4])
    return new_model
