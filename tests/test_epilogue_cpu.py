"""The per-entry math of the SHIPPED kernels on CPU: csrc/pmf_epilogue.cuh (noise-model loss, dloss/dz, the ordinal
threshold derivatives, the NaN = missing bit test -- shared by the FP32 kernel and both tensor-core kernels) is compiled
with g++ against a host stand-in for the CUDA headers (tests/cuda_stub/epilogue/; the fast-math intrinsics map to libm)
and compared with the oracle's restatement entry by entry: every distribution, z from -60 to 60 (saturation of the
sigmoid / softplus / exp forms), every category, infinite outer thresholds, missing entries.  The GPU suite compares the
kernels themselves with the same oracle; this pins the source's formulas without a device."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pmf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIST = {"normal": 0, "bernoulli": 1, "poisson": 2, "ordinal3": 3, "bernoulli_sq_hinge": 4, "ordinal_sq_hinge3": 5}   # pmf_dist
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def epi(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("epi") / "libepi.so")
    stub = os.path.join(ROOT, "tests", "cuda_stub", "epilogue")
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-x", "c++", "-I", stub, "-I",
           os.path.join(ROOT, "pathmatfac.jl_b200", "csrc"), os.path.join(stub, "epilogue_harness.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(fp)


def _data(dist, n, rng):
    if dist == "normal":
        return rng.standard_normal(n) * 3
    if dist in ("bernoulli", "bernoulli_sq_hinge"):
        return (rng.random(n) < 0.5).astype(float)
    if dist == "poisson":
        return rng.poisson(3.0, n).astype(float)
    return rng.integers(1, 4, n).astype(float)


@pytest.mark.parametrize("dist", list(DIST))
def test_noise_models_of_the_shipped_epilogue_match_the_oracle(epi, dist):
    rng = np.random.default_rng(DIST[dist])
    z = np.concatenate([np.linspace(-60, 60, 481), rng.standard_normal(2000) * 4, [0.0, -0.0, 1e-6, -1e-6]])
    if dist == "poisson":
        z = np.clip(z, -60, 20)                                   # exp(z) stays finite in float32 products with the datum
    n = len(z)
    a = _data(dist, n, rng)
    a[::17] = np.nan                                              # missing entries
    a[5] = np.inf                                                 # non-finite data never contribute either
    thr = np.array([-np.inf, -0.7, 0.9, np.inf])
    z32, a32, th32 = z.astype(np.float32), a.astype(np.float32), thr.astype(np.float32)
    l, g = np.empty(n, np.float32), np.empty(n, np.float32)
    obs = np.empty(n, np.int32)
    epi.epi_noise_eval(DIST[dist], n, _p(z32), _p(a32), _p(th32), C.c_float(O.ORDINAL_EPS), C.c_float(O.SQ_HINGE_MARGIN),
                       _p(l), _p(g), obs.ctypes.data_as(C.POINTER(C.c_int32)))
    assert np.array_equal(obs.astype(bool), np.isfinite(a32))     # the mask IS the NaN pattern, bit-exact
    lo, go = O.noise_loss_grad(dist, z32.astype(np.float64), a32.astype(np.float64), thr)
    assert np.all(l[~np.isfinite(a32)] == 0) and np.all(g[~np.isfinite(a32)] == 0)
    assert np.all(np.isfinite(l)) and np.all(np.isfinite(g))
    # float32 evaluation against the float64 restatement: relative to the size of the terms that are subtracted
    a0 = np.where(np.isfinite(a32), a32, 0.0).astype(np.float64)
    scale_l = np.maximum(np.abs(lo), 1.0) + np.abs(a0 * z32)
    scale_g = np.maximum(np.abs(go), 1.0)
    # ordinal3: l = -log(s(hi - z) - s(lo - z) + eps) in float32 cancels once a sigmoid is within a few ulp of 1 (an entry
    # whose category lies far from z): inherent to the formula in FP32 (the gradient 1 - s(hi - z) - s(lo - z) does not
    # cancel), so the loss and the threshold derivatives (which divide by the same difference) are compared where every
    # |threshold - z| <= 8, and over the whole range the loss must stay finite and below -log(eps)
    near = np.abs(z32) <= 7.0 if dist == "ordinal3" else np.ones(n, bool)
    tol_l = 1e-4 if dist == "ordinal3" else 2e-6
    assert np.max((np.abs(l - lo) / scale_l)[near]) < tol_l, dist
    assert np.max(np.abs(g - go) / scale_g) < 2e-6, dist
    if dist == "ordinal3":
        assert l.min() > -1e-6 and l.max() <= -np.log(O.ORDINAL_EPS) + 1e-3
    if dist.startswith("ordinal"):
        g1, g2 = np.empty(n, np.float32), np.empty(n, np.float32)
        epi.epi_threshold_grads(DIST[dist], n, _p(z32), _p(a32), _p(th32), C.c_float(O.ORDINAL_EPS),
                                C.c_float(O.SQ_HINGE_MARGIN), _p(g1), _p(g2))
        o1, o2 = O.noise_threshold_grads(dist, z32.astype(np.float64), np.where(np.isfinite(a32), a32, np.nan).astype(np.float64), thr)
        for got, ref in ((g1, o1), (g2, o2)):
            assert np.all(np.isfinite(got))
            assert np.max((np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))[near]) < 1e-4, dist
        # a category touches only its own two thresholds (1: t1; 2: t1 and t2; 3: t2)
        cat = np.where(np.isfinite(a32), a32, 0).astype(int)
        assert not g2[cat == 1].any() and not g1[cat == 3].any()
        assert epi.epi_is_ordinal(DIST[dist]) == 1
    else:
        assert epi.epi_is_ordinal(DIST[dist]) == 0
