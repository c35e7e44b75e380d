import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem
M, N = 10000, 30000
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
eng = P.Engine(model)
eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
ref = eng.loss_grad(include_reg=False)
for prec in (0, 2, 0, 2, 2, 0):
    eng.set_loss_grad_kernel(_lib.KERNEL_TC, prec)
    got = eng.loss_grad(include_reg=False)
    dX = got["dX"]
    bad = ~np.isfinite(dX)
    rows = np.unique(np.nonzero(bad)[1])
    err = np.abs(dX - ref["dX"]).max(axis=0)
    worst = np.argsort(-np.nan_to_num(err, nan=1e30))[:5]
    print("prec", prec, "nonfinite", bad.sum(), "samples with NaN:", rows[:10], len(rows),
          "worst cols", worst, err[worst], "dY ok", np.isfinite(got["dY"]).all(), flush=True)
eng.close()
