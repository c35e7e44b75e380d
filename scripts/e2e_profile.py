"""cProfile of the end-to-end call bench.py times (mf_fit on a host-resident C2 model, 200 epochs)."""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, "/root/repo")
import numpy as np
import torch

import pathmatfac_b200 as P
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem

M, N = 10000, 30000
pinned = torch.empty((N, M), dtype=torch.float32, pin_memory=True)
D = pinned.numpy().T
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=2, missing=0.3, data_out=D,
                         model_kwargs=dict(lambda_X_l2=1.0))
X0, Y0 = model.matfac.X.copy(), model.matfac.Y.copy()
kw = dict(lr=0.05, max_epochs=200, update_X=True, update_Y=True, update_col_layers=True, rel_tol=-1.0, abs_tol=-1.0,
          verbosity=0, check_every=1 << 20)
for rep in range(3):
    model.matfac.X[...] = X0
    model.matfac.Y[...] = Y0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rep == 2:
        pr = cProfile.Profile()
        pr.enable()
    h = P.mf_fit(model, **kw)
    if rep == 2:
        pr.disable()
    torch.cuda.synchronize()
    print(f"rep {rep}: {time.perf_counter() - t0:.4f} s, epochs {h['epochs']}, device part {h.get('device_ms')}", flush=True)
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
