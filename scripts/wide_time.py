"""Time the K > 64 data pass (wide_tc.cu) under PMF_WIDE_FLAGS experiment settings in one process."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import simulate_problem
M, N, K = (int(x) for x in sys.argv[1].split("x"))
settings = [int(a) for a in sys.argv[2:]] or [0]
blocks = (("mutation", "bernoulli", 2 * N // 5), ("mrnaseq", "normal", N - 2 * N // 5))
model = simulate_problem(M, blocks=blocks, K=K, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
eng = P.Engine(model)
eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
for fl in settings:
    os.environ["PMF_WIDE_FLAGS"] = str(fl)
    for _ in range(3):
        eng.loss_grad(include_reg=False)
    eng.set_profiling(True)
    for _ in range(8):
        eng.loss_grad(include_reg=False)
    n, mean_ms, min_ms = eng.get_profile()
    eng.set_profiling(False)
    print(f"{M}x{N} K={K} flags={fl}: n={n} data pass mean {mean_ms:.4f} ms min {min_ms:.4f} ms", flush=True)
eng.close()
