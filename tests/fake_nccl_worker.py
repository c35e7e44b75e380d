"""Worker of tests/test_abi_on_fake_runtime.py for the exchange step: three ranks as THREADS of one interpreter, each with
its own handle on a row shard, the host-only CUDA runtime and NCCL stand-ins (tests/cuda_stub/fake_cudart, fake_nccl)
underneath, the REAL libpmf in between.  ``pmf_fit`` runs its sample-sharded epoch loop with ncclAllReduce inside; the
"gradients" a data pass would have produced are planted by the runtime stand-in when the data-pass kernel is launched
(rank r: r + 1 everywhere in the shared gradient buffer, 10 (r + 1) in the two rank-local loss scalars)."""
import ctypes as C
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fake = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)
nccl = C.CDLL(sys.argv[2], mode=C.RTLD_GLOBAL)
os.environ["PMF_LIB"] = sys.argv[3]

import numpy as np  # noqa: E402

import pathmatfac_b200 as P  # noqa: E402
from pathmatfac_b200 import _lib  # noqa: E402
from pathmatfac_b200.dist import shard_rows  # noqa: E402

assert "torch" not in sys.modules
lib = _lib.load()
fake.fake_fill_on_launch.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_double, C.c_int]
rng = np.random.default_rng(0)
v6 = (C.c_longlong * 6)()


def scenario(R, EPOCHS, model, M, kernel, plant_on, check_every):
    """R ranks as threads, each with its own handle on a row shard of `model`; returns what the exchange did."""
    ident = (C.c_uint8 * 128)()
    assert lib.pmf_comm_unique_id(ident) == 0
    log0 = nccl.fake_nccl_log_count()
    engines, bufs = [], []
    for r in range(R):
        eng = P.Engine(model, rows=shard_rows(M, r, R))
        eng._ck(lib.pmf_comm_init_rank(eng.h, R, r, ident))
        p, n = C.c_void_p(), C.c_int64()
        eng._ck(lib.pmf_shared_grad_buffer(eng.h, C.byref(p), C.byref(n)))
        ps, ns = C.c_void_p(), C.c_int64()
        eng._ck(lib.pmf_shared_scalar_buffer(eng.h, C.byref(ps), C.byref(ns)))
        fake.fake_fill_on_launch(plant_on, p.value, n.value, float(r + 1), 0)
        fake.fake_fill_on_launch(plant_on, ps.value, ns.value, 10.0 * (r + 1), 1)
        engines.append(eng)
        bufs.append((p.value, n.value, ps.value, ns.value))
    hist, errors = [None] * R, []

    def run(r):
        try:
            eng = engines[r]
            hist[r] = eng.fit(eng.make_opts(epoch=1, max_epochs=EPOCHS, lr=0.1, update_X=1, update_Y=1, update_col_layers=1,
                                            kernel=kernel, rel_tol=0.0, abs_tol=0.0, check_every=check_every))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    alive = [t.is_alive() for t in threads]
    out = {"alive": alive, "errors": errors, "ranks": R, "epochs": EPOCHS}
    if any(alive) or errors:
        return out
    sums, scal = [], []
    for p, n, ps, ns in bufs:
        sums.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy())
        scal.append(np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_double)), shape=(ns,)).copy())
    out["grad_buffer_len"] = [int(b[1]) for b in bufs]
    out["grad_buffers_all_equal_sum"] = bool(all(np.all(s == float(sum(range(1, R + 1)))) for s in sums))
    out["scalars"] = [s.tolist() for s in scal]
    log = []
    for i in range(log0, nccl.fake_nccl_log_count()):
        nccl.fake_nccl_log(i, v6)
        log.append(dict(zip(("rank", "count", "dtype", "in_place", "grouped", "nranks"), list(v6))))
    out["nccl_log"] = log
    out["kernel_launches"] = [h["kernel_launches"] for h in hist]
    out["term"] = [h["term_code"] for h in hist]
    for eng in engines:
        eng._ck(lib.pmf_comm_destroy(eng.h))
        eng.close()
    st = (C.c_int * 2)()
    nccl.fake_nccl_state(st)
    out["nccl_mismatches"], out["nccl_live_comms"] = st[0], st[1]
    lib.pmf_release_cached_memory()
    c = (C.c_long * 10)()
    fake.fake_counters(c)
    out["live_blocks"], out["bad_frees"], out["oob_copies"] = c[2], c[8], c[9]
    return out


M, N, K = 31, 24, 4
D = rng.standard_normal((M, N)).astype(np.float32)
model = P.PathMatFacModel(D, K=K, feature_views=["a"] * 10 + ["b"] * 14, lambda_X_l2=1.0)
out = scenario(3, 4, model, M, _lib.KERNEL_FFMA, b"data_pass", 2)

# BASELINE configs[4] (80 000 x 50 000, K = 128 over 8 GPUs) at its real feature count and latent dimension: two ranks with
# 1 000 samples each -- the exchanged payload (dY + column-parameter gradients) does not depend on the sample count -- on the
# tcgen05 kernels for K > 64 (the "gradients" are planted when the link kernel is launched)
M5, N5, K5 = 2000, 50000, 128
D5 = np.zeros((M5, N5), np.float32)
D5[::7, ::5] = np.nan
model5 = P.PathMatFacModel(D5, K=K5, feature_views=["mutation"] * 20000 + ["mrnaseq"] * 30000,
                           feature_distributions=["bernoulli"] * 20000 + ["normal"] * 30000, lambda_X_l2=1.0)
fake.fake_clear_launches()
out["c5"] = scenario(2, 4, model5, M5, _lib.KERNEL_TC, b"zlink", 1)
out["c5"]["N"], out["c5"]["K"] = N5, K5
print(json.dumps(out))
os._exit(0)
