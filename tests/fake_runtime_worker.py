"""Worker of tests/test_abi_on_fake_runtime.py: a fresh interpreter (no torch, hence no real libcudart in the process)
loads the host-only libcudart stand-in (tests/cuda_stub/fake_cudart), then the REAL libpmf built with `-cudart shared`
(argv), and drives it through the Python mirror exactly as on a GPU box.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fake = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)
os.environ["PMF_LIB"] = sys.argv[2]

import numpy as np  # noqa: E402

import pathmatfac_b200 as P  # noqa: E402
from pathmatfac_b200 import _lib  # noqa: E402

assert "torch" not in sys.modules
lib = _lib.load()
OUT = {}


sys.path.insert(0, os.path.join(ROOT, "tests"))
import fake_runtime_util as U  # noqa: E402

U.fake = fake
launches, maps, counters, short = U.launches, U.maps, U.counters, U.short


def model_(M, N, K, batch_views=0, ordinal=False, seed=0, **kw):
    rng = np.random.default_rng(seed)
    nv = max(batch_views, 2)
    cuts = np.linspace(0, N, nv + 1).astype(int)
    views = [f"v{i}" for i in range(nv) for _ in range(cuts[i + 1] - cuts[i])]
    dists = ["normal"] * N
    if ordinal:
        dists = ["normal"] * (N - 6) + ["ordinal3"] * 6
        views = views[:N - 6] + ["zord"] * 6
    D = rng.standard_normal((M, N)).astype(np.float32)
    if ordinal:
        D[:, N - 6:] = rng.integers(1, 4, (M, 6))
    D[rng.random((M, N)) < 0.1] = np.nan
    batch = {f"v{i}": [f"b{int(b)}" for b in rng.integers(0, 3 + i, M)] for i in range(batch_views)} or None
    cond = ["c"] * M if batch else None
    m = P.PathMatFacModel(D, K=K, feature_views=views, feature_distributions=dists, sample_conditions=cond, batch_dict=batch, **kw)
    mf = m.matfac
    mf.X[...] = rng.standard_normal(mf.X.shape)
    mf.Y[...] = rng.standard_normal(mf.Y.shape)
    ct = mf.col_transform
    ct.layers[0].logsigma[...] = rng.standard_normal(N)
    ct.layers[2].mu[...] = rng.standard_normal(N)
    if batch:
        for v in ct.layers[1].logdelta.values + ct.layers[3].theta.values:
            v[...] = rng.standard_normal(v.shape)
    return m


def snapshot(m):
    mf, ct = m.matfac, m.matfac.col_transform
    out = {"X": mf.X.copy(), "Y": mf.Y.copy(), "logsigma": ct.unwrapped(0).logsigma.copy(), "mu": ct.unwrapped(2).mu.copy()}
    l1 = ct.unwrapped(1)
    if hasattr(l1, "logdelta"):
        for i, (a, b) in enumerate(zip(l1.logdelta.values, ct.unwrapped(3).theta.values)):
            out[f"logdelta{i}"], out[f"theta{i}"] = a.copy(), b.copy()
    return out


def scramble(m):
    for v in snapshot(m).values():
        pass
    mf, ct = m.matfac, m.matfac.col_transform
    mf.X[...] = -7
    mf.Y[...] = -7
    ct.unwrapped(0).logsigma[...] = -7
    ct.unwrapped(2).mu[...] = -7
    l1 = ct.unwrapped(1)
    if hasattr(l1, "logdelta"):
        for a, b in zip(l1.logdelta.values, ct.unwrapped(3).theta.values):
            a[...] = -7
            b[...] = -7


def same(a, b):
    return sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)


fit_kw = dict(lr=0.1, update_X=1, update_Y=1, update_col_layers=1, rel_tol=0.0, abs_tol=0.0)

# ---- S1: every parameter goes to the "device" and comes back unchanged (K = 5 -> Kp = 8, ragged M / N) ---------------------
m = model_(53, 41, 5, batch_views=2, ordinal=True, lambda_X_l2=1.0)
want = snapshot(m)
eng = P.Engine(m)
launches()
scramble(m)
eng.pull_params()
OUT["s1_round_trip"] = same(snapshot(m), want)
# thresholds: interior values survive, the outer ones stay infinite
nm = m.matfac.noise_model
OUT["s1_thresholds"] = [[float(x) for x in n.ext_thresholds] for n in nm.noises if n.ext_thresholds is not None]

# ---- S2: re-sending an unchanged layout keeps batch parameters and optimiser state; a new layout re-allocates them ------------
acc = np.full((m.matfac.Y.shape[1], 5), 3.25, np.float32)
eng._ck(lib.pmf_set_opt_state(eng.h, 1, 0, _lib.fptr(acc)))
eng.push_structure()                       # what reweight_col_losses / a second mf_fit on a resident model do
scramble(m)
eng.pull_params()
back = np.zeros_like(acc)
eng._ck(lib.pmf_get_opt_state(eng.h, 1, 0, _lib.fptr(back)))
OUT["s2_same_layout_keeps_values"] = same(snapshot(m), want)
OUT["s2_same_layout_keeps_opt_state"] = bool(np.array_equal(back, acc))
eng.reset_opt_state(1e-8)
eng._ck(lib.pmf_get_opt_state(eng.h, 1, 0, _lib.fptr(back)))
accx = np.zeros((53, 5), np.float32)
eng._ck(lib.pmf_get_opt_state(eng.h, 0, 0, _lib.fptr(accx)))
OUT["s2_reset_gives_epsilon"] = bool(np.all(back == np.float32(1e-8)) and np.all(accx == np.float32(1e-8)))
ld = m.matfac.col_transform.unwrapped(1).logdelta
ld.batch_index[0] = (ld.batch_index[0] + 1) % ld.values[0].shape[0]            # another assignment of samples to batches
eng.push_structure()
scramble(m)
eng.pull_params()
s = snapshot(m)
OUT["s2_new_layout_keeps_column_params"] = bool(np.array_equal(s["logsigma"], want["logsigma"]) and np.array_equal(s["mu"], want["mu"])
                                                and np.array_equal(s["X"], want["X"]))
OUT["s2_new_layout_zeroes_batch_params"] = bool(not s["theta0"].any() and not s["logdelta1"].any())
eng.close()
lib.pmf_release_cached_memory()
OUT["s2_counters_after_close"] = counters()

# ---- S3: which kernels an epoch launches, per configuration -------------------------------------------------------------------------


def epoch_launches(m, kernel, epochs=3, **extra):
    eng = P.Engine(m)
    launches()
    h = eng.fit(eng.make_opts(epoch=1, max_epochs=epochs, kernel=kernel, **{**fit_kw, **extra}))
    mp = maps()
    ls = launches()
    eng.close()
    names = short([x["name"] for x in ls])
    return {"names": names, "launches": ls, "maps": mp, "reported": h["kernel_launches"], "term": h["term_code"]}


OUT["s3_ffma"] = epoch_launches(model_(60, 45, 6, lambda_X_l2=1.0), _lib.KERNEL_FFMA)
OUT["s3_tc"] = epoch_launches(model_(300, 400, 16, lambda_X_l2=1.0), _lib.KERNEL_TC)
OUT["s3_tc_batch"] = epoch_launches(model_(300, 400, 16, batch_views=2, lambda_X_l2=1.0), _lib.KERNEL_TC)
OUT["s3_tc_wide"] = epoch_launches(model_(300, 400, 128, lambda_X_l2=1.0), _lib.KERNEL_TC)
OUT["s3_ffma_alternating"] = epoch_launches(model_(60, 45, 6, lambda_X_l2=1.0), _lib.KERNEL_FFMA, alternating=1)
OUT["s3_auto_small"] = epoch_launches(model_(1999, 3100, 64), _lib.KERNEL_AUTO, epochs=1)["names"]
OUT["s3_auto_large"] = epoch_launches(model_(2000, 3000, 64), _lib.KERNEL_AUTO, epochs=1)["names"]
try:
    epoch_launches(model_(80, 90, 300), _lib.KERNEL_TC, epochs=1)
    OUT["s3_tc_refuses_k300"] = False
except _lib.PmfError as e:
    OUT["s3_tc_refuses_k300"] = str(e)
lib.pmf_release_cached_memory()
OUT["s3_counters_after_close"] = counters()

# ---- S3b: graph regulariser, statistics passes, the end-to-end call -------------------------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests"))
rng = np.random.default_rng(9)
Ng, Kg = 30, 3
fids = list(range(1, Ng + 1))
graphs = []
for k in range(Kg):
    el = [[int(a), int(b), 1.0] for a, b in rng.integers(1, Ng + 1, (20, 2)) if a != b]
    el += [[int(rng.integers(1, Ng + 1)), f"virt{k}", 1.0] for _ in range(3)]
    graphs.append(el)
mg = P.PathMatFacModel(rng.standard_normal((25, Ng)).astype(np.float32), feature_ids=fids, feature_graphs=graphs, lambda_Y_graph=1.0)
OUT["s3_network"] = epoch_launches(mg, _lib.KERNEL_FFMA)

m = model_(50, 40, 4, batch_views=2, lambda_X_l2=1.0)
eng = P.Engine(m)
launches()
eng.column_stats()
a = short([x["name"] for x in launches()])
eng.link_col_sqerr()
b = short([x["name"] for x in launches()])
cnt, sq = eng.batch_stats()
c = short([x["name"] for x in launches()])
OUT["s3_stats"] = {"column_stats": a, "link_col_sqerr": b, "batch_stats": c, "batch_shapes": [list(x.shape) for x in cnt]}
eng.close()

# mf_fit on a host-resident model = create handle, upload, fit, read back, destroy; from the second call on the allocator
# cache serves every block
m = model_(64, 48, 8, lambda_X_l2=1.0)
lib.pmf_release_cached_memory()
per_call = []
for _ in range(3):
    c0 = counters()
    h = P.mf_fit(m, lr=0.1, max_epochs=3, update_X=True, update_Y=True, update_col_layers=True, verbosity=0, kernel=_lib.KERNEL_FFMA)
    c1 = counters()
    per_call.append({"mallocs": c1["mallocs"] - c0["mallocs"], "frees": c1["frees"] - c0["frees"],
                     "host_allocs": c1["host_allocs"] - c0["host_allocs"], "h2d": h["h2d_bytes"], "d2h": h["d2h_bytes"]})
OUT["s3_mf_fit_calls"] = per_call
lib.pmf_release_cached_memory()
OUT["s3b_counters"] = counters()

# ---- S4: the host-driven sharded step (dist.ShardedFit's sequence) --------------------------------------------------------------------
m = model_(40, 30, 4, lambda_X_l2=1.0)
eng = P.Engine(m, rows=range(10, 30))
o = eng.make_opts(epoch=1, max_epochs=2, kernel=_lib.KERNEL_FFMA, **fit_kw)
launches()
eng._ck(lib.pmf_fit_start(eng.h, C.byref(o)))
seq = {"start": short([x["name"] for x in launches()])}
eng._ck(lib.pmf_epoch_begin(eng.h, C.byref(o)))
seq["begin"] = short([x["name"] for x in launches()])
p, n = C.c_void_p(), C.c_int64()
eng._ck(lib.pmf_shared_grad_buffer(eng.h, C.byref(p), C.byref(n)))
seq["grad_floats"] = n.value
eng._ck(lib.pmf_shared_scalar_buffer(eng.h, C.byref(p), C.byref(n)))
seq["scalar_doubles"] = n.value
eng._ck(lib.pmf_epoch_end(eng.h, C.byref(o)))
seq["end"] = short([x["name"] for x in launches()])
OUT["s4_sharded"] = seq
eng.close()

# ---- S5: error paths -------------------------------------------------------------------------------------------------------------------------


def create_error():
    try:
        P.Engine(model_(20, 10, 3)).close()
        return None
    except _lib.PmfError as e:
        return str(e)


lib.pmf_release_cached_memory()
before = counters()
fake.fake_set(0, 10, -1)
OUT["s5_no_device"] = create_error()
fake.fake_set(1, 9, -1)
OUT["s5_wrong_architecture"] = create_error()
leaks = []
for k in (0, 3, 9, 20):                              # every cudaMalloc after the k-th fails (the cache is empty: nothing is served from it)
    lib.pmf_release_cached_memory()
    fake.fake_set(1, 10, k)
    err = create_error()
    fake.fake_set(1, 10, -1)
    lib.pmf_release_cached_memory()
    c = counters()
    leaks.append({"fail_after": k, "error": err, "live_blocks": c["live_blocks"] - before["live_blocks"], "bad_frees": c["bad_frees"]})
OUT["s5_alloc_failures"] = leaks
eng = P.Engine(model_(20, 10, 3))
rc = lib.pmf_set_batch_values(eng.h, 3, None, None)
OUT["s5_bad_view"] = [rc, lib.pmf_last_error(eng.h).decode()]
bad = np.array([0, 5], np.int32)
rc = lib.pmf_set_noise(eng.h, 2, _lib.iptr(bad), _lib.iptr(np.array([5, 9], np.int32)), _lib.iptr(np.array([0, 0], np.int32)), None, None)
OUT["s5_noise_ranges_must_cover"] = [rc, lib.pmf_last_error(eng.h).decode()]
eng.close()

# ---- S7: launch geometry and tensor maps over a sweep of ragged / degenerate shapes, both kernel families ------------------------
sweep = []
rs = np.random.default_rng(77)
shapes = [(1, 1, 1), (1, 300, 8), (50, 1, 3), (2, 2, 64), (129, 257, 64), (3000, 3, 2), (3, 3000, 5), (127, 129, 72), (64, 128, 128),
          (65, 130, 200), (33, 600, 256)]
shapes += [(int(rs.integers(1, 700)), int(rs.integers(2, 900)), int(rs.choice([1, 3, 8, 10, 25, 64, 65, 100, 128, 256]))) for _ in range(14)]
for (Ms, Ns, Ks) in shapes:
    for bv in (0, 2):
        if bv and (Ns < 4 or Ms < 4):
            continue
        for kern in (_lib.KERNEL_FFMA, _lib.KERNEL_TC):
            rec = {"shape": [Ms, Ns, Ks], "batch_views": bv, "kernel": kern, "error": None}
            try:
                r = epoch_launches(model_(Ms, Ns, Ks, batch_views=bv, seed=Ms + Ns), kern, epochs=1)
                rec["bad_geometry"] = [x for x in r["launches"] if min(x["grid"]) < 1 or x["block"][0] * x["block"][1] * x["block"][2] > 1024
                                       or x["smem"] > 232448]
                rec["bad_maps"] = [mm for mm in r["maps"] if mm["rc"] != 0]
                rec["n_launches"], rec["reported"] = len(r["launches"]), r["reported"]
                rec["tc"] = any("_tc_" in n or "zlink" in n for n in r["names"])
            except _lib.PmfError as e:
                rec["error"] = str(e)
            sweep.append(rec)
OUT["s7_sweep"] = sweep
lib.pmf_release_cached_memory()
OUT["s7_counters"] = counters()

# ---- S8: the ABI in the order julia/PathMatFacB200.jl calls it (create_handle: create, set_data, set_batch_layout; every
# mf_fit!: factors, column parameters, batch values, regularisers, layer penalties, frozen masks, noise LAST, reset of the
# optimiser state on first use, fit, read-back) -- the shim cannot be run here, its call order can ------------------------------
from pathmatfac_b200._lib import pmf_dims  # noqa: E402

m = model_(45, 36, 5, batch_views=2, ordinal=True, lambda_X_l2=1.0)
want = snapshot(m)
eng = P.Engine.__new__(P.Engine)
eng.lib, eng.model, eng.rows, eng.device = lib, m, range(0, 45), 0
eng.M, eng.N, eng.K, eng.h, eng.n_views, eng.h2d_bytes, eng.d2h_bytes = 45, 36, 5, C.c_void_p(), 0, 0, 0
steps = []


def step(name, rc):
    steps.append([name, int(rc), lib.pmf_last_error(eng.h).decode() if rc else ""])


step("pmf_create", lib.pmf_create(C.byref(pmf_dims(45, 36, 5, 0)), C.byref(eng.h)))
A = np.ascontiguousarray(np.asarray(m.data, dtype=np.float32).T)
step("pmf_set_data", lib.pmf_set_data(eng.h, _lib.fptr(A)))
ld = m.matfac.col_transform.unwrapped(1).logdelta
bcs = np.array([r.start for r in ld.col_ranges], np.int32)
bce = np.array([r.stop for r in ld.col_ranges], np.int32)
nbv = np.array([v.shape[0] for v in ld.values], np.int32)
bos = np.ascontiguousarray(np.stack(ld.batch_index).astype(np.int32))
step("pmf_set_batch_layout", lib.pmf_set_batch_layout(eng.h, 2, _lib.iptr(bcs), _lib.iptr(bce), _lib.iptr(nbv), _lib.iptr(bos)))
eng.n_views = 2
shim_fits = []
for call in range(2):
    try:
        eng.push_params()                                   # pmf_set_factors, pmf_set_col_params, pmf_set_batch_values
        eng.push_regs()                                     # pmf_clear_reg / pmf_set_reg_*, layer penalties, pmf_set_frozen, pmf_set_noise
        if call == 0:
            eng.reset_opt_state(1e-8)
        launches()
        h = eng.fit(eng.make_opts(epoch=1, max_epochs=2, **fit_kw))
        names = short([x["name"] for x in launches()])
        scramble(m)
        eng.pull_params()
        shim_fits.append({"ok": True, "round_trip": same(snapshot(m), want), "names": names, "term": h["term_code"]})
    except _lib.PmfError as e:
        shim_fits.append({"ok": False, "error": str(e)})
OUT["s8_shim_order"] = {"steps": steps, "fits": shim_fits}
eng.close()

# ---- S9: how often the host waits for the device inside pmf_fit: the epoch loop only enqueues (termination is evaluated on the
# device); the host synchronises once per `check_every` epochs to read the stop flag, and once at the end ----------------
fake.fake_syncs.restype = C.c_long
m = model_(300, 400, 16, lambda_X_l2=1.0)
eng = P.Engine(m)
eng.reset_opt_state(1e-8)
eng.fit(eng.make_opts(epoch=1, max_epochs=3, kernel=_lib.KERNEL_TC, **fit_kw))
sync_rule = []
launches()
for kw in (dict(no_terminate=1, check_every=1 << 20), dict(check_every=8), dict(check_every=1)):
    for E in (40, 400):
        s0, c0 = fake.fake_syncs(), counters()
        eng.fit(eng.make_opts(epoch=1, max_epochs=E, kernel=_lib.KERNEL_TC, **{**fit_kw, **kw}))
        c1 = counters()
        sync_rule.append({"check_every": kw["check_every"], "epochs": E, "syncs": fake.fake_syncs() - s0, "launches": len(launches()),
                          "mallocs": c1["mallocs"] - c0["mallocs"], "host_allocs": c1["host_allocs"] - c0["host_allocs"]})
OUT["s9_sync_rule"] = sync_rule
eng.close()

# ---- S6: guard zones (PMF_GUARD=1 in the environment; without it there is nothing to check) ----------------------------------------
m = model_(70, 60, 12, batch_views=2, ordinal=True, lambda_X_l2=1.0)
eng = P.Engine(m)
eng.fit(eng.make_opts(epoch=1, max_epochs=2, kernel=_lib.KERNEL_FFMA, **fit_kw))
eng.fit(eng.make_opts(epoch=1, max_epochs=2, kernel=_lib.KERNEL_TC, **fit_kw))
eng.column_stats(); eng.batch_stats(); eng.loss_grad(); eng.pull_params(); eng.push_structure(); eng.push_params()
nb, bad = C.c_int64(0), C.c_int64(0)
rc = lib.pmf_check_guards(C.byref(nb), C.byref(bad))
OUT["s6_guards"] = {"rc": rc, "buffers": nb.value, "bad": bad.value, "enabled": os.environ.get("PMF_GUARD") == "1"}
eng.close()
lib.pmf_release_cached_memory()
OUT["final_counters"] = counters()
print(json.dumps(OUT))
