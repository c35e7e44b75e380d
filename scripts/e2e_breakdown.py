"""Where the time of the end-to-end call goes (bench.py's e2e leg: mf_fit on a host-resident C2 model, 20 epochs):
wall-clock per phase of Engine / pmf_* calls, and the pinned H2D copy rate of the box as the floor of the data upload."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import numpy as np
import torch

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem

M, N, K = 10000, 30000, 64
pinned = torch.empty((N, M), dtype=torch.float32, pin_memory=True)
D = pinned.numpy().T
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=K, seed=2, missing=0.3, data_out=D, model_kwargs=dict(lambda_X_l2=1.0))

# PCIe floor: pinned -> device copy of the same 1.2 GB through torch
dev = torch.empty((N, M), dtype=torch.float32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(pinned, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pinned H2D 1.2 GB: {dt * 1e3:.2f} ms = {pinned.numel() * 4 / dt / 1e9:.1f} GB/s", flush=True)
del dev
torch.cuda.empty_cache()


def timed(label, fn, acc):
    t0 = time.perf_counter()
    r = fn()
    acc.append((label, (time.perf_counter() - t0) * 1e3))
    return r


lib = _lib.load()
for rep in range(4):
    acc = []
    t_all = time.perf_counter()
    eng = timed("Engine(): pmf_create", lambda: P.Engine.__new__(P.Engine), acc)
    # Engine.__init__ by hand, phase by phase
    eng.lib = lib; eng.model = model
    eng.rows = range(0, M); eng.M, eng.N, eng.K = M, N, K
    eng.h = C.c_void_p(); eng.device = 0
    dims = _lib.pmf_dims(M, N, K, 0)
    timed("pmf_create", lambda: lib.pmf_create(C.byref(dims), C.byref(eng.h)), acc)
    eng.n_views = 0; eng.h2d_bytes = 0; eng.d2h_bytes = 0
    timed("push_data", lambda: eng.push_data(model.data), acc)
    timed("push_structure", eng.push_structure, acc)
    timed("push_params", eng.push_params, acc)
    timed("reset_opt_state", lambda: eng.reset_opt_state(1e-8), acc)
    o = eng.make_opts(max_epochs=20, epoch=1, lr=0.05, update_X=1, update_Y=1, update_col_layers=1, rel_tol=-1.0, abs_tol=-1.0,
                      check_every=1 << 20)
    h = timed("fit (20 epochs)", lambda: eng.fit(o), acc)
    timed("pull_params", eng.pull_params, acc)
    timed("close", eng.close, acc)
    total = (time.perf_counter() - t_all) * 1e3
    print(f"rep {rep}: total {total:.2f} ms | " + " | ".join(f"{k} {v:.2f}" for k, v in acc[1:]) + f" | device_ms {h.get('device_ms')}", flush=True)

# the call bench.py times, for comparison
X0, Y0 = model.matfac.X.copy(), model.matfac.Y.copy()
for rep in range(3):
    model.matfac.X[...] = X0; model.matfac.Y[...] = Y0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    P.mf_fit(model, lr=0.05, max_epochs=20, update_X=True, update_Y=True, update_col_layers=True, rel_tol=-1.0, abs_tol=-1.0,
             verbosity=0, check_every=1 << 20)
    torch.cuda.synchronize()
    print(f"mf_fit call {rep}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
