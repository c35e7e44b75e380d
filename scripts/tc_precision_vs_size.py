"""Relative error of the tensor-core gradients (single-pass TF32 contractions) against the FP32 kernel on the same
handle, as a function of problem size and K: the data behind the PMF_KERNEL_AUTO thresholds (pmf_abi.cu)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


cases = [(1024, 4096, 64), (2000, 3000, 64), (4000, 6000, 64), (6000, 12000, 64), (10000, 30000, 64),
         (1024, 4096, 128), (2000, 3000, 128), (4000, 6000, 128), (10000, 30000, 128), (4000, 6000, 256), (10000, 30000, 256)]
for M, N, K in cases:
    model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=K, seed=3, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
    eng = P.Engine(model)
    out = {"M": M, "N": N, "K": K}
    for phase in ("start", "fitted"):
        if phase == "fitted":      # gradients near a fitted model are residual-dominated: the harder case
            eng.reset_opt_state(1e-8)
            eng.fit(eng.make_opts(epoch=1, max_epochs=30, lr=0.1, update_X=1, update_Y=1, update_col_layers=1, kernel=_lib.KERNEL_TC,
                                  rel_tol=0.0, abs_tol=0.0))
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        ref = eng.loss_grad(include_reg=False)
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        got = eng.loss_grad(include_reg=False)
        out[phase] = {k: rel(got[k], ref[k]) for k in ("dX", "dY", "dmu", "dlogsigma")}
        out[phase]["loss"] = abs(got["loss"] - ref["loss"]) / abs(ref["loss"])
    eng.close()
    print(json.dumps(out), flush=True)
