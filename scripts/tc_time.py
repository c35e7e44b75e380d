"""Time the tcgen05 data pass alone at the C2 shape (CUDA events inside libpmf); PMF_TC_ABLATE for experiments."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem
M, N = 10000, 30000
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
blocks = scale_blocks(C2_BLOCKS, N)
if os.environ.get("PMF_BLOCKS"):      # e.g. PMF_BLOCKS=normal : one all-<dist> block (per-noise-model timing)
    blocks = (({"normal": "methylation", "bernoulli": "mutation", "poisson": "counts"}[os.environ["PMF_BLOCKS"]], os.environ["PMF_BLOCKS"], N),)
model = simulate_problem(M, blocks=blocks, K=64, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0))
eng = P.Engine(model)
eng.set_loss_grad_kernel(_lib.KERNEL_TC, prec)
for _ in range(3):
    eng.loss_grad(include_reg=False)
eng.set_profiling(True)
for _ in range(10):
    eng.loss_grad(include_reg=False)
n, mean_ms, min_ms = eng.get_profile()
print(f"flags={os.environ.get('PMF_TC_FLAGS','0')} blocks={os.environ.get('PMF_BLOCKS','C2')} ablate={os.environ.get('PMF_TC_ABLATE','0')} prec={prec}: n={n} mean {mean_ms:.4f} ms min {min_ms:.4f} ms", flush=True)
eng.close()
