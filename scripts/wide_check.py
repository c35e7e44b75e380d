"""K > 64 tensor-core path (wide_tc.cu) against the FP32 kernel on the same handle: prints the relative differences
instead of asserting (first look at a new kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import simulate_problem


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


cases = [(300, 385, 72), (523, 385, 128), (300, 385, 256), (2000, 3000, 128), (2000, 3000, 256)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for M, N, K in cases:
    nb, nn = N // 4, N // 2
    blocks = (("mutation", "bernoulli", nb), ("methylation", "normal", nn), ("counts", "poisson", N - nb - nn))
    model = simulate_problem(M, blocks=blocks, K=K, seed=7, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
    eng = P.Engine(model)
    try:
        eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
        ref = eng.loss_grad(include_reg=False)
        eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
        got = eng.loss_grad(include_reg=False)
        print(f"M={M} N={N} K={K}: loss {abs(got['loss'] - ref['loss']) / abs(ref['loss']):.2e} "
              + " ".join(f"{k} {rel(got[k], ref[k]):.2e}" for k in ("dmu", "dlogsigma", "dY", "dX")), flush=True)
        for k in ("dY", "dX"):
            g, r = got[k], ref[k]
            if rel(g, r) > 1e-3:
                # where is it wrong?  per-row (factor) and per-column-block error
                e = np.abs(g - r)
                print(f"   {k}: worst factor rows {np.argsort(-e.sum(1))[:6]}, column blocks of 128 with error:",
                      [int(b) for b in np.nonzero(np.add.reduceat(e.sum(0), np.arange(0, e.shape[1], 128)) > 1e-3 * np.abs(r).sum() / max(1, e.shape[1] // 128))[0][:12]])
    finally:
        eng.close()
