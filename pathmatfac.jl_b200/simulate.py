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
