"""Worker of tests/test_abi_on_fake_runtime.py::test_baseline_configs_at_full_size: the host side of the REAL libpmf (built
with `-cudart shared`) on the host-only CUDA runtime stand-in, at the FULL dimensions of BASELINE.json's configs -- the
sizes the GPU suite only touches through scripts/config_times.py and bench.py.  No kernel runs; what is recorded per
config is everything the host decides at that size: which kernels `PMF_KERNEL_AUTO` picks, launch geometry, dynamic shared
memory, every TMA descriptor (extents, boxes, strides as the driver would check them), the launch counter, and the device
memory the handle holds.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fake = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)
os.environ["PMF_LIB"] = sys.argv[2]

import numpy as np  # noqa: E402

import pathmatfac_b200 as P  # noqa: E402
from pathmatfac_b200 import _lib  # noqa: E402
from pathmatfac_b200.simulate import C2_BLOCKS, simulate_problem  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
import fake_runtime_util as U  # noqa: E402

U.fake = fake
counters, launches, maps, short = U.counters, U.launches, U.maps, U.short

assert "torch" not in sys.modules
lib = _lib.load()
rng = np.random.default_rng(0)


def graphs(N, K, n_edges, n_virtual):
    """Per-factor edge lists of the size of C4 (BASELINE configs[3]: ~2 M edges over K = 256 factors, virtual nodes)."""
    out = []
    for k in range(K):
        a, b = rng.integers(1, N + 1, size=n_edges), rng.integers(1, N + 1, size=n_edges)
        el = [[int(x), int(y), 1.0] for x, y in zip(a, b) if x != y]
        el += [[int(rng.integers(1, N + 1)), f"virt{k}_{v}", 1.0] for v in range(n_virtual)]
        out.append(el)
    return out


def build(name):
    if name == "C2":
        return simulate_problem(10000, blocks=C2_BLOCKS, K=64, seed=2, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
    if name == "C3":
        return simulate_problem(10000, blocks=C2_BLOCKS, K=64, seed=3, missing=0.3, batch_views=[b[0] for b in C2_BLOCKS],
                                n_batches=40, n_conditions=20)
    if name == "C4a":
        g = graphs(30000, 256, 7800, 780)
        return simulate_problem(10000, blocks=(("mrnaseq", "normal", 30000),), K=256, seed=4, missing=0.3,
                                model_kwargs=dict(feature_graphs=g, lambda_Y_graph=1.0, lambda_Y_selective_l1=0.5))
    if name == "C5":
        return simulate_problem(10000, blocks=(("mutation", "bernoulli", 20000), ("mrnaseq", "normal", 30000)), K=128, seed=5,
                                missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
    raise SystemExit(name)


OUT = {}
for name in sys.argv[3:]:
    t0 = time.time()
    model = build(name)
    M, N = model.data.shape
    rec = {"M": M, "N": N, "K": int(model.matfac.X.shape[0]), "error": None}
    try:
        eng = P.Engine(model)
        eng.reset_opt_state(1e-8)
        launches()
        rec["live_bytes"] = counters()["live_bytes"]
        h = eng.fit(eng.make_opts(epoch=1, max_epochs=2, kernel=_lib.KERNEL_AUTO, lr=0.05, update_X=1, update_Y=1,
                                  update_col_layers=1, no_terminate=1, check_every=1 << 20, rel_tol=0.0, abs_tol=0.0))
        mp = maps()                      # before the launch log is cleared (that clears the descriptors too)
        ls = launches()
        rec["live_bytes_after_fit"] = counters()["live_bytes"]
        eng.close()
        rec["names"] = short([x["name"] for x in ls])
        rec["launches"] = [dict(x, name=n) for x, n in zip(ls, rec["names"])]
        rec["maps"] = mp
        rec["reported"], rec["term"] = h["kernel_launches"], h["term_code"]
    except _lib.PmfError as e:
        rec["error"] = str(e)
    del model
    lib.pmf_release_cached_memory()
    rec["counters_after_close"] = counters()
    rec["seconds"] = round(time.time() - t0, 1)
    OUT[name] = rec
print(json.dumps(OUT))
