import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem
M, N = 10000, 30000
pinned = torch.empty((N, M), dtype=torch.float32, pin_memory=True)
D = pinned.numpy().T
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=2, missing=0.3, data_out=D, model_kwargs=dict(lambda_X_l2=1.0))
for rep in range(3):
    t0 = time.perf_counter()
    eng = P.Engine(model, device=0, upload_data=False)
    t1 = time.perf_counter()
    eng.push_data(model.data)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    o = eng.make_opts(max_epochs=20, epoch=1, lr=0.05, update_X=1, update_Y=1, update_col_layers=1, rel_tol=-1.0, abs_tol=-1.0, check_every=1 << 20)
    eng.reset_opt_state(1e-8)
    h = eng.fit(o)
    t3 = time.perf_counter()
    eng.pull_params()
    t4 = time.perf_counter()
    eng.close()
    t5 = time.perf_counter()
    print(f"rep {rep}: create+structure+params {t1-t0:.3f}s  push_data {t2-t1:.3f}s  fit(20) {t3-t2:.3f}s  pull {t4-t3:.3f}s  close {t5-t4:.3f}s  epochs {h['epochs']}", flush=True)
