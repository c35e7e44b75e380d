"""``save_model`` / ``load_model`` -- the reference's whole-model serialisation (src/model_io.jl:9-19:
``BSON.@save filename model`` with ``model.data = nothing`` unless ``save_data``; ``BSON.load``) for the mirror's
model object, and ``write_model_arrays`` -- the flat parameter export of
analyses/scripts/julia/bson_to_hdf.jl:18-71 (same dataset names) into a NumPy ``.npz`` container and, through
``h5lite`` (the HDF5 subset libhdf5 writes by default, restated in NumPy because the image has no HDF5 library), into
the very HDF5 layout of that script (``write_model_to_hdf``); ``load_omic_data`` / ``load_batches`` /
``save_omic_data`` / ``save_transformed`` read and write the study's input / output files
(analyses/scripts/julia/fit_matfac.jl:60-117, script_util.jl:147-180).

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
    """Flat dict of the trained arrays under the dataset names of bson_to_hdf.jl:18-71: ids / views / conditions /
    ``data_idx`` (1-based), ``X``, ``Y``, ``logsigma``, ``mu``; per batched view i (1-based) ``logdelta/values_i``,
    ``logdelta/col_range_i`` (``collect(cr)``: every 1-based column index of the view), ``theta/values_i``,
    ``theta/col_range_i``, ``theta/batch_ids_i``; with a feature-set ARD ``fsard/A/i`` and the dense Float32
    ``fsard/S/i``."""
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.pull_params()
    from .layers import FrozenLayer
    from .regularizers import FeatureSetARDReg

    def unwrap(layer):
        return layer.layer if isinstance(layer, FrozenLayer) else layer

    def strings(v):
        return np.array([str(x) for x in v], dtype=object)

    lay = [unwrap(l) for l in _layers(model)]
    out = {
        "feature_ids": strings(model.feature_ids), "feature_views": strings(model.feature_views),
        "sample_ids": strings(model.sample_ids), "data_idx": np.asarray(model.data_idx, dtype=np.int64) + 1,
        "X": np.asarray(model.matfac.X), "Y": np.asarray(model.matfac.Y),
        "logsigma": np.asarray(lay[0].logsigma), "mu": np.asarray(lay[2].mu),
    }
    # the reference writes safe_convert(model.sample_conditions) unconditionally (:32); a model built without
    # conditions has none to write
    if model.sample_conditions is not None:
        out["sample_conditions"] = strings(model.sample_conditions)
    for name, idx, attr in (("logdelta", 1, "logdelta"), ("theta", 3, "theta")):
        ba = getattr(lay[idx], attr, None)           # Identity layers (no batch_dict) have none
        if ba is not None:
            for i, (v, cr, ids) in enumerate(zip(ba.values, ba.col_ranges, ba.row_batch_ids)):
                out[f"{name}/values_{i + 1}"] = np.asarray(v)
                out[f"{name}/col_range_{i + 1}"] = np.arange(cr.start + 1, cr.stop + 1, dtype=np.int64)
                if name == "theta":
                    out[f"theta/batch_ids_{i + 1}"] = strings(ids)
    regs = getattr(model.matfac.Y_reg, "regularizers", [model.matfac.Y_reg])
    for r in regs:
        if isinstance(r, FeatureSetARDReg):
            for i, (A, S) in enumerate(zip(r.A, r.S)):
                out[f"fsard/A/{i + 1}"] = np.asarray(A)
                out[f"fsard/S/{i + 1}"] = np.asarray(S.todense() if hasattr(S, "todense") else S, dtype=np.float32)
    return out


def write_model_arrays(filename, model):
    """The same arrays as a NumPy ``.npz`` container ("/" in a name becomes "__")."""
    np.savez(filename, **{k.replace("/", "__"): v for k, v in model_arrays(model).items()})


# ---- the reference's HDF5 layouts (analyses/scripts/julia) ---------------------------------------------------------

BATCHED_ASSAYS = ("mrnaseq", "methylation")                     # script_util.jl:25


def write_model_to_hdf(out_hdf, model):
    """analyses/scripts/julia/bson_to_hdf.jl:18-71, dataset for dataset (``model_arrays``), as an HDF5 file.  Matrices
    are stored the way HDF5.jl stores Julia's (dimensions reversed: h5lite's ``julia=True``)."""
    from . import h5lite
    w = h5lite.Writer()
    for name, value in model_arrays(model).items():
        w.write(name, value)
    return w.save(out_hdf)


def read_model_hdf(path) -> dict:
    """Every dataset of a ``write_model_to_hdf`` / bson_to_hdf.jl file under its slash-separated name."""
    from . import h5lite
    f = h5lite.File(path)
    return {k: f.read(k) for k in f.visit()}


def barcode_to_batch(barcode: str) -> str:
    """script_util.jl:147-157: the last two dash-separated terms of a TCGA barcode ("" stays "")."""
    if barcode == "":
        return ""
    return "-".join(barcode.split("-")[-2:])


def load_omic_data(omic_hdf, omic_types):
    """fit_matfac.jl:60-82: ``omic_data/{feature_assays, feature_genes, data, instances, instance_groups}`` with the
    features filtered to ``omic_types``.  Returns (data M x N', sample_ids, sample_conditions, feature_genes,
    feature_assays)."""
    from . import h5lite
    f = h5lite.File(omic_hdf)
    assays = np.asarray(f.read("omic_data/feature_assays"), dtype=object)
    keep = np.isin(assays, list(set(omic_types)))
    genes = np.asarray(f.read("omic_data/feature_genes"), dtype=object)[keep]
    data = f.read("omic_data/data")[:, keep]
    return (data, list(f.read("omic_data/instances")), list(f.read("omic_data/instance_groups")), list(genes),
            list(assays[keep]))


def load_batches(omic_hdf, omic_types):
    """fit_matfac.jl:85-101: ``barcodes/data`` (M x n_assays strings) and ``barcodes/features`` -> the ``batch_dict`` of
    the model constructor for the batched assays, or None."""
    from . import h5lite
    f = h5lite.File(omic_hdf)
    barcodes = f.read("barcodes/data")
    cols = {a: i for i, a in enumerate(f.read("barcodes/features"))}
    out = {a: [barcode_to_batch(b) for b in barcodes[:, cols[a]]] for a in omic_types if a in BATCHED_ASSAYS}
    return out or None


def save_omic_data(output_hdf, feature_assays, feature_genes, instance_names, instance_groups, omic_matrix,
                   barcodes=None, barcode_features=None):
    """script_util.jl:165-180 (+ the ``barcodes`` group ``load_batches`` reads, when given)."""
    from . import h5lite
    omic_matrix = np.asarray(omic_matrix)
    assert omic_matrix.shape[1] == len(feature_assays)
    assert omic_matrix.shape[0] == len(instance_names)
    assert len(instance_names) == len(instance_groups)
    w = h5lite.Writer()
    w.write("omic_data/feature_assays", list(feature_assays))
    w.write("omic_data/feature_genes", list(feature_genes))
    w.write("omic_data/instances", list(instance_names))
    w.write("omic_data/instance_groups", list(instance_groups))
    w.write("omic_data/data", omic_matrix)
    if barcodes is not None:
        w.write("barcodes/data", np.asarray(barcodes, dtype=object))
        w.write("barcodes/features", list(barcode_features))
    return w.save(output_hdf)


def save_transformed(transformed_X, instances, instance_groups, target, output_hdf):
    """fit_matfac.jl:104-116."""
    from . import h5lite
    w = h5lite.Writer()
    w.write("X", np.asarray(transformed_X))
    w.write("instances", list(instances))
    w.write("instance_groups", list(instance_groups))
    w.write("target", target)
    return w.save(output_hdf)
