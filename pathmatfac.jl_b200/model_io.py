"""``save_model`` / ``load_model`` -- the reference's whole-model serialisation (src/model_io.jl:9-19:
``BSON.@save filename model`` with ``model.data = nothing`` unless ``save_data``; ``BSON.load``) for the mirror's
model object, and ``write_model_arrays`` -- the flat parameter export of
analyses/scripts/julia/bson_to_hdf.jl:18-71 (same dataset names) into a NumPy ``.npz`` container.

Host-only code off the hot path.  In a Julia deployment the model object lives in Julia and
the reference's own ``save_model`` keeps working: the shim's ``pull_params!`` (and ``pmf_get_opt_state`` for the AdaGrad
accumulators) brings every trained array back from the device first.  A device-resident mirror model is synchronised the
same way before it is written; the device handle itself is never serialised."""
from __future__ import annotations

import pickle

import numpy as np

FORMAT = "pathmatfac_b200.model.v1"


def save_model(model, filename, save_data: bool = False):
    """src/model_io.jl:9-14.  Like the reference, ``save_data=False`` drops ``model.data`` from the saved object
    (the reference sets the field of the caller's model to nothing; here the caller's model keeps its data)."""
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.pull_params()                    # trained parameters live on the device while the model is resident
    data = model.data
    model._engine = None
    if not save_data:
        model.data = None
    try:
        with open(filename, "wb") as f:
            pickle.dump({"format": FORMAT, "model": model}, f, protocol=pickle.HIGHEST_PROTOCOL)
    finally:
        model.data = data
        model._engine = eng


def load_model(filename):
    """src/model_io.jl:16-19.  Resume = call ``fit`` / ``mf_fit_adapt_lr`` again on the loaded model."""
    with open(filename, "rb") as f:
        d = pickle.load(f)
    if not isinstance(d, dict) or d.get("format") != FORMAT:
        raise ValueError(f"{filename}: not a {FORMAT} file")
    model = d["model"]
    model._engine = None
    return model


def _layers(model):
    return model.matfac.col_transform.layers


def model_arrays(model) -> dict:
    """Flat dict of the trained arrays under the dataset names of bson_to_hdf.jl:18-71."""
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.pull_params()
    from .layers import FrozenLayer
    from .regularizers import FeatureSetARDReg

    def unwrap(layer):
        return layer.layer if isinstance(layer, FrozenLayer) else layer

    lay = [unwrap(l) for l in _layers(model)]
    out = {
        "feature_ids": np.asarray(model.feature_ids), "feature_views": np.asarray(model.feature_views),
        "sample_ids": np.asarray(model.sample_ids), "data_idx": np.asarray(model.data_idx) + 1,   # 1-based like Julia's
        "X": np.asarray(model.matfac.X), "Y": np.asarray(model.matfac.Y),
        "logsigma": np.asarray(lay[0].logsigma), "mu": np.asarray(lay[2].mu),
    }
    if model.sample_conditions is not None:
        out["sample_conditions"] = np.asarray(model.sample_conditions)
    for name, idx, attr in (("logdelta", 1, "logdelta"), ("theta", 3, "theta")):
        ba = getattr(lay[idx], attr, None)           # Identity layers (no batch_dict) have none
        if ba is not None:
            for i, v in enumerate(ba.values):
                out[f"{name}/values_{i + 1}"] = np.asarray(v)
            out[f"{name}/col_ranges"] = np.asarray([[r.start + 1, r.stop] for r in ba.col_ranges])
    regs = getattr(model.matfac.Y_reg, "regularizers", [model.matfac.Y_reg])
    for r in regs:
        if isinstance(r, FeatureSetARDReg):
            for i, (A, S) in enumerate(zip(r.A, r.S)):
                out[f"fsard/A/{i + 1}"] = np.asarray(A)
                out[f"fsard/S/{i + 1}"] = np.asarray(S.todense() if hasattr(S, "todense") else S)
    return out


def write_model_arrays(filename, model):
    np.savez(filename, **{k.replace("/", "__"): v for k, v in model_arrays(model).items()})
