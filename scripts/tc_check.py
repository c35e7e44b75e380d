"""Compare the tcgen05 data pass with the FP32 FFMA data pass (and the oracle at small sizes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run(M, N, precision, seed=5, missing=0.3):
    model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=seed, missing=missing,
                             model_kwargs=dict(lambda_X_l2=1.0))
    eng = P.Engine(model)
    eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
    ref = eng.loss_grad(include_reg=False)
    eng.set_loss_grad_kernel(_lib.KERNEL_TC, precision)
    t0 = time.time()
    got = eng.loss_grad(include_reg=False)
    dt = time.time() - t0
    eng.close()
    print(f"M={M} N={N} prec={precision}: loss ref {ref['loss']:.6e} tc {got['loss']:.6e} rel {abs(got['loss']-ref['loss'])/abs(ref['loss']):.2e} | "
          f"dmu {relerr(got['dmu'], ref['dmu']):.2e} dls {relerr(got['dlogsigma'], ref['dlogsigma']):.2e} "
          f"dY {relerr(got['dY'], ref['dY']):.2e} dX {relerr(got['dX'], ref['dX']):.2e}  ({dt*1e3:.1f} ms incl. copies)", flush=True)
    return got, ref


if __name__ == "__main__":
    sizes = [(128, 128), (300, 260), (1000, 2000)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        sizes = [(10000, 30000)]
    for M, N in sizes:
        for prec in (0, 2):
            run(M, N, prec)
