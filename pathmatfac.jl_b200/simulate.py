"""Seeded synthetic inputs restating ``simulate_params!`` / ``simulate_data!``
(src/simulate_params.jl:7-255) for the BASELINE configs.  NumPy RNG (Julia's
Xoshiro stream is not reproducible here).  Float32 and blocked by view so the 10k x 30k
matrix is generated in a few seconds.

Extensions the reference lacks, documented in DESIGN.md: bernoulli / poisson samplers
(``Bern(sigmoid(z))``, ``Pois(exp(z))``) and a "counts" view for the poisson assay."""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from .model import PathMatFacModel

# per-view tables, src/simulate_params.jl:124-142 (+ the "counts" extension)
VIEW_MU_MEAN = {"mrnaseq": 10.0, "methylation": 0.0, "cna": 0.0, "mutation": -1.5, "counts": 1.0}
VIEW_MU_STD = {"mrnaseq": 2.0, "methylation": 0.1, "cna": 0.1, "mutation": 0.1, "counts": 0.5}
VIEW_LOGSIGMA_MEAN = {"mrnaseq": 0.5, "methylation": 0.1, "cna": math.log(2.0), "mutation": 0.0, "counts": -1.0}
VIEW_LOGSIGMA_STD = {"mrnaseq": 0.1, "methylation": 0.1, "cna": 0.001, "mutation": 0.001, "counts": 0.1}

# (view, distribution, columns) blocks of the TCGA-scale configs, already in the constructor's
# (distribution, view) sorted order (src/model.jl:50)
C2_BLOCKS = (("mutation", "bernoulli", 5000), ("methylation", "normal", 10000),
             ("mrnaseq", "normal", 10000), ("counts", "poisson", 5000))


def scale_blocks(blocks, N):
    tot = sum(b[2] for b in blocks)
    out = [(v, d, max(1, int(round(n * N / tot)))) for v, d, n in blocks]
    out[-1] = (out[-1][0], out[-1][1], N - sum(b[2] for b in out[:-1]))
    return tuple(out)


def simulate_problem(M: int, blocks=C2_BLOCKS, K: int = 64, seed: int = 2, missing: float = 0.3,
                     batch_views: Sequence[str] = (), n_batches: int = 0, n_conditions: int = 0,
                     noise: float = 0.1, data_out: Optional[np.ndarray] = None, model_kwargs=None,
                     sort_batches: bool = False):
    """Build a PathMatFacModel on synthetic data.

    Parameters are drawn like simulate_params! (X, Y ~ N(0,1), :7-11; per-view logsigma / mu
    tables, :144-174; logsigma -= log sqrt(K), :212; batch values = per-batch centre
    N(0, 0.25) + N(0, 0.25), :96-101), data = forward(model) + per-assay sampling
    (:224-255), missingness iid Bernoulli(missing).  The model is then re-initialised at
    fresh random X, Y (a fit needs a starting point away from the generating one).
    ``data_out``: optional preallocated (M, N) float32 Fortran-ordered array (e.g. a view of
    pinned memory) that receives the data."""
    rng = np.random.default_rng(seed)
    N = sum(b[2] for b in blocks)
    views, dists = [], []
    for v, d, n in blocks:
        views += [v] * n
        dists += [d] * n
    f32 = np.float32
    Xg = rng.standard_normal((K, M), dtype=f32)
    D = data_out if data_out is not None else np.empty((M, N), dtype=f32, order="F")
    assert D.shape == (M, N) and D.dtype == f32
    logsigma = np.empty(N, f32)
    mu = np.empty(N, f32)
    batch_dict = None
    if batch_views:
        batch_dict = {v: [int(b) for b in rng.integers(0, n_batches, size=M)] for v in batch_views}
        if sort_batches:      # samples grouped by batch in every view (the same sample order works for all)
            batch_dict = {v: sorted(b) for v, b in batch_dict.items()}
    conditions = None
    if n_conditions or batch_views:
        conditions = list(np.sort(rng.integers(0, max(n_conditions, 1), size=M)))
    gen_batch = {}
    c0 = 0
    for v, d, n in blocks:
        sl = slice(c0, c0 + n)
        Yg = rng.standard_normal((K, n), dtype=f32)
        logsigma[sl] = rng.standard_normal(n, dtype=f32) * f32(VIEW_LOGSIGMA_STD[v]) + f32(VIEW_LOGSIGMA_MEAN[v])
        mu[sl] = rng.standard_normal(n, dtype=f32) * f32(VIEW_MU_STD[v]) + f32(VIEW_MU_MEAN[v])
        logsigma[sl] -= f32(math.log(math.sqrt(K)))
        Z = (Xg.T @ Yg) * np.exp(logsigma[sl])[None, :]
        if batch_dict is not None and v in batch_dict:
            names = list(dict.fromkeys(batch_dict[v]))
            lut = {b: i for i, b in enumerate(names)}
            bidx = np.fromiter((lut[b] for b in batch_dict[v]), dtype=np.int64, count=M)
            nb = len(names)
            ld = (rng.standard_normal(nb, dtype=f32) * f32(0.25))[:, None] + rng.standard_normal((nb, n), dtype=f32) * f32(0.25)
            th = (rng.standard_normal(nb, dtype=f32) * f32(0.25))[:, None] + rng.standard_normal((nb, n), dtype=f32) * f32(0.25)
            gen_batch[v] = (ld, th)
            Z *= np.exp(ld)[bidx, :]
            Z += mu[sl][None, :]
            Z += th[bidx, :]
        else:
            Z += mu[sl][None, :]
        if d == "normal":
            Z += rng.standard_normal(Z.shape, dtype=f32) * f32(noise)
        elif d == "bernoulli":
            Z = (rng.random(Z.shape, dtype=f32) < 1.0 / (1.0 + np.exp(-Z))).astype(f32)
        elif d == "poisson":
            Z = rng.poisson(np.exp(np.minimum(Z, 10.0))).astype(f32)
        elif d == "bernoulli_sq_hinge":
            Z = (Z > 0).astype(f32)
        elif d in ("ordinal3", "ordinal_sq_hinge3"):
            Z = (1.0 + (Z > -1.0) + (Z > 1.0)).astype(f32)
        if missing > 0:
            Z[rng.random(Z.shape, dtype=f32) < missing] = np.nan
        D[:, sl] = Z
        c0 += n
    kw = dict(model_kwargs or {})
    model = PathMatFacModel(D, K=K, sample_conditions=conditions, feature_views=views,
                            feature_distributions=dists, batch_dict=batch_dict, rng=rng, **kw)
    assert list(model.data_idx) == list(range(N)), "blocks must be listed in (distribution, view) order"
    mf = model.matfac
    mf.X[...] = rng.standard_normal((K, M), dtype=f32)
    mf.Y[...] = rng.standard_normal((K, N), dtype=f32)
    mf.col_transform.layers[0].logsigma[...] = logsigma
    mf.col_transform.layers[2].mu[...] = mu
    if batch_dict is not None:
        for i, v in enumerate(mf.col_transform.layers[1].logdelta.col_range_ids):
            mf.col_transform.layers[1].logdelta.values[i][...] = gen_batch[v][0] * f32(0.5)
            mf.col_transform.layers[3].theta.values[i][...] = gen_batch[v][1] * f32(0.5)
    return model


# ---- factor simulators of the ARD priors (src/simulate_params.jl:14-79) -------------------------------------------

def simulate_factor_ard(K: int, N: int, rng, ard_alpha=1.01, ard_beta=0.01) -> np.ndarray:
    """simulate_params!(X, ::ARDRegularizer) (src/simulate_params.jl:14-21): tau ~ Gamma(shape alpha, scale 1/beta) per
    entry, X = randn / sqrt(tau)."""
    tau = rng.gamma(shape=ard_alpha, scale=1.0 / ard_beta, size=(K, N))
    return (rng.standard_normal((K, N)) / np.sqrt(tau)).astype(np.float32)


def corrupt_S(S, S_add_corruption: float, S_rem_corruption: float, rng) -> np.ndarray:
    """corrupt_S (src/simulate_params.jl:24-43): per feature set remove round(rem * |set|) members, add
    round(add * |set|) non-members, renormalise the row to 1 / sqrt(|corrupted set|).  Dense L x N float32 result."""
    S_new = np.asarray(S.todense() if hasattr(S, "todense") else S, dtype=np.float32).copy()
    L, N = S_new.shape
    for l in range(L):
        cur = np.flatnonzero(S_new[l] != 0)
        comp = np.flatnonzero(S_new[l] == 0)
        rm_n = int(round(S_rem_corruption * len(cur)))
        add_n = int(round(len(cur) * S_add_corruption))
        to_remove = rng.choice(cur, size=rm_n, replace=False) if rm_n else np.zeros(0, np.int64)
        to_add = rng.choice(comp, size=min(add_n, len(comp)), replace=False) if add_n else np.zeros(0, np.int64)
        corrupted = np.union1d(np.setdiff1d(cur, to_remove), to_add)
        S_new[l, :] = 0
        if len(corrupted):
            S_new[l, corrupted] = 1.0 / np.sqrt(len(corrupted))
    return S_new


def simulate_factor_fsard(Y: np.ndarray, reg, rng, ard_alpha=1.01, beta0=0.001, S_add_corruption=0.1,
                          S_rem_corruption=0.1) -> None:
    """simulate_params!(Y, ::FeatureSetARDReg) (src/simulate_params.jl:45-79), in place on Y (K x N) and on the
    regulariser's A / S: per view one random feature set per factor (A[l, k] = |5 randn|), the sets corrupted, Y small
    (randn / sqrt(tau), tau ~ Gamma(alpha, 1 / beta0)) outside the assigned sets and +-1 inside them."""
    import scipy.sparse as sp
    for v, (cr, A, S) in enumerate(zip(reg.col_ranges, reg.A, reg.S)):
        Yv = Y[:, cr.start:cr.stop]
        K, N = Yv.shape
        L = S.shape[0]
        A[...] = 0
        for k in range(K):
            A[int(rng.integers(0, L)), k] = abs(rng.standard_normal() * 5.0)
        S_new = corrupt_S(S, S_add_corruption, S_rem_corruption, rng)
        reg.S[v] = sp.csr_matrix(S_new, dtype=np.float32)
        tau = rng.gamma(shape=ard_alpha, scale=1.0 / beta0, size=(K, N))
        beta = A.T @ S_new
        Yv[...] = (rng.standard_normal((K, N)) / np.sqrt(tau)) * (beta == 0)
        Yv += (beta > 0) * rng.choice([-1.0, 1.0], size=(K, N))


def add_missingness(data: np.ndarray, view_cols: Dict[str, slice], batch_of_sample: Dict[str, Sequence], rng,
                    missingness=0.1) -> None:
    """add_missingness! of the study's simulator (analyses/scripts/julia/simulate_matfac.jl:106-130): per batched view
    whole rows go missing, batch by batch in random batch order, until round(M * missingness) rows are gone (the last
    batch touched loses a random subset of its rows).  ``data`` is modified in place (NaN = missing)."""
    M = data.shape[0]
    for view, cols in view_cols.items():
        if view not in batch_of_sample:
            continue
        b = np.asarray(batch_of_sample[view])
        to_remove = int(round(M * missingness))
        unq = list(dict.fromkeys(b.tolist()))
        for ub in [unq[i] for i in rng.permutation(len(unq))]:
            rows = np.flatnonzero(b == ub)
            n_rem = min(len(rows), to_remove)
            if n_rem < len(rows):
                rows = rng.choice(rows, size=n_rem, replace=False)
            data[rows, cols] = np.nan
            to_remove -= n_rem


# ---- flat binary export: the same bytes for a Julia run of the reference --------------------------------

def export_problem(model, directory: str) -> str:
    """Write the inputs of a fit as raw little-endian arrays plus a text manifest (one line per array: name, element
    type, rows, cols; Julia column-major order, so ``read!(io, Matrix{T}(undef, rows, cols))`` restores each one).
    ``julia/load_exported_problem.jl`` rebuilds the PathMatFacModel from it.  Returns the manifest path."""
    import os
    os.makedirs(directory, exist_ok=True)
    mf = model.matfac
    ct = mf.col_transform
    arrays = {"data": np.asarray(model.data, np.float32), "X": np.asarray(mf.X, np.float32), "Y": np.asarray(mf.Y, np.float32),
              "logsigma": np.asarray(ct.unwrapped(0).logsigma, np.float32)[:, None],
              "mu": np.asarray(ct.unwrapped(2).mu, np.float32)[:, None],
              "col_weights": mf.noise_model.weights()[:, None]}
    from .layers import BatchShift
    l4 = ct.unwrapped(3)
    if isinstance(l4, BatchShift):
        ld, th = ct.unwrapped(1).logdelta, l4.theta
        for v, name in enumerate(th.col_range_ids):
            arrays[f"batch_of_sample__{name}"] = (np.asarray(th.batch_index[v], np.int32) + 1)[:, None]    # 1-based for Julia
            arrays[f"theta__{name}"] = np.asarray(th.values[v], np.float32)
            arrays[f"logdelta__{name}"] = np.asarray(ld.values[v], np.float32)
    lines = []
    for name, a in arrays.items():
        a = np.asfortranarray(a)
        with open(os.path.join(directory, name + ".bin"), "wb") as f:
            f.write(a.tobytes(order="F"))
        lines.append(f"{name} {'Float32' if a.dtype == np.float32 else 'Int32'} {a.shape[0]} {a.shape[1]}")
    with open(os.path.join(directory, "feature_views.txt"), "w") as f:
        f.write("\n".join(str(v) for v in model.feature_views) + "\n")
    with open(os.path.join(directory, "feature_distributions.txt"), "w") as f:
        f.write("\n".join(str(d) for d in model.feature_distributions) + "\n")
    with open(os.path.join(directory, "sample_conditions.txt"), "w") as f:
        f.write("\n".join(str(c) for c in (model.sample_conditions if model.sample_conditions is not None else [])) + "\n")
    manifest = os.path.join(directory, "manifest.txt")
    with open(manifest, "w") as f:
        f.write("\n".join(lines) + "\n")
    return manifest


def import_problem_arrays(directory: str) -> Dict[str, np.ndarray]:
    """Read back what export_problem wrote (round-trip check of the format)."""
    import os
    out = {}
    for line in open(os.path.join(directory, "manifest.txt")):
        name, ty, r, c = line.split()
        dt = np.float32 if ty == "Float32" else np.int32
        out[name] = np.fromfile(os.path.join(directory, name + ".bin"), dtype=dt).reshape((int(r), int(c)), order="F")
    return out


def export_problem_hdf(model, path: str, barcodes=None, barcode_features=None) -> str:
    """The same inputs in the study's own HDF5 layout (``omic_data/{data, feature_assays, feature_genes, instances,
    instance_groups}`` [+ ``barcodes/*``], analyses/scripts/julia/script_util.jl:165-180), so that the reference's
    driver script fit_matfac.jl -- not only a custom loader -- can read a simulated problem (``load_omic_data``,
    fit_matfac.jl:60-82): views become the assays, feature ids the genes, sample conditions the instance groups."""
    from .model_io import save_omic_data
    M = np.asarray(model.data).shape[0]
    groups = model.sample_conditions if model.sample_conditions is not None else ["all"] * M
    return save_omic_data(path, [str(v) for v in model.feature_views], [str(f) for f in model.feature_ids],
                          [str(x) for x in model.sample_ids], [str(g) for g in groups], np.asarray(model.data),
                          barcodes=barcodes, barcode_features=barcode_features)
