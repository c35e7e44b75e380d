"""Print a few entries of the tcgen05 gradients next to the FFMA reference (descriptor debugging)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem
M, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 128)
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
eng = P.Engine(model)
eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
ref = eng.loss_grad(include_reg=False)
eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
got = eng.loss_grad(include_reg=False)
eng.close()
np.set_printoptions(precision=4, linewidth=200, suppress=True)
for k in ("dY", "dX"):
    g, r = got[k], ref[k]
    print(k, "shape", g.shape, "norm got", np.linalg.norm(g), "ref", np.linalg.norm(r), "nonfinite", (~np.isfinite(g)).sum())
    print(" got[:8,:6]\n", g[:8, :6], "\n ref[:8,:6]\n", r[:8, :6])
    # best matching permutation hints: correlation of got rows with ref rows
    if np.isfinite(g).all() and np.linalg.norm(g) > 0:
        gn = g / (np.linalg.norm(g, axis=1, keepdims=True) + 1e-30)
        rn = r / (np.linalg.norm(r, axis=1, keepdims=True) + 1e-30)
        C = gn @ rn.T
        print(" row k of got best matches ref row:", np.argmax(np.abs(C), axis=1)[:16], np.max(np.abs(C), axis=1)[:8])
        gn = g / (np.linalg.norm(g, axis=0, keepdims=True) + 1e-30)
        rn = r / (np.linalg.norm(r, axis=0, keepdims=True) + 1e-30)
        C = gn.T @ rn
        print(" col of got best matches ref col:", np.argmax(np.abs(C), axis=1)[:24], np.max(np.abs(C), axis=1)[:8])
err = np.linalg.norm(got["dX"] - ref["dX"], axis=0) / (np.linalg.norm(ref["dX"], axis=0) + 1e-30)
print("per-sample relerr by 16-sample block:", [f"{err[i:i+16].max():.1e}" for i in range(0, M, 16)])
