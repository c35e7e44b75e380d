"""Time the tcgen05 data pass at the C2 shape under several PMF_TC_ABLATE settings in one process (experiments;
results are wrong when a bit is set, only the time means something)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem
M, N = 10000, 30000
settings = [int(a) for a in sys.argv[1:]] or [0]
model = simulate_problem(M, blocks=scale_blocks(C2_BLOCKS, N), K=64, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
eng = P.Engine(model)
eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
for ab in settings:
    os.environ["PMF_TC_ABLATE"] = str(ab)
    for _ in range(3):
        eng.loss_grad(include_reg=False)
    eng.set_profiling(True)
    for _ in range(10):
        eng.loss_grad(include_reg=False)
    n, mean_ms, min_ms = eng.get_profile()
    eng.set_profiling(False)
    print(f"ablate={ab}: n={n} mean {mean_ms:.4f} ms min {min_ms:.4f} ms", flush=True)
eng.close()
