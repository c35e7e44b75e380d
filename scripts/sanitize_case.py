"""Small invocations of every kernel family at ragged shapes -- the fused tensor-core data pass without and with batch
layers, the K > 64 kernels, the FP32 kernel, the statistics pass, the graph regulariser, three fit epochs through the
fused epoch pass -- written for compute-sanitizer (closed on this pool) and run with PMF_GUARD=1 instead: every device
buffer of the library then carries guard zones, verified after each case (pmf_check_guards)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import simulate_problem

cases = [("fused <0,0> 200x700 K=64", 200, 700, 64, None), ("fused <0,0> 1203x953 K=40", 1203, 953, 40, None),
         ("fused batch <0,1> 1203x953 K=32", 1203, 953, 32, 5), ("wide 300x385 K=72", 300, 385, 72, None),
         ("wide 523x385 K=256", 523, 385, 256, None)]
only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
for n, (name, M, N, K, nb) in enumerate(cases):
    if only and n not in only:
        continue
    nbn, nn = N // 4, N // 2
    blocks = (("mutation", "bernoulli", nbn), ("methylation", "normal", nn), ("counts", "poisson", N - nbn - nn))
    kw = dict(batch_views=["methylation", "counts"], n_batches=nb) if nb else {}
    mk = dict(lambda_X_l2=1.0)
    if K <= 40 and not nb:      # graph regulariser on Y for one of the cases
        rng = np.random.default_rng(3)
        mk.update(feature_graphs=[[[int(a), int(b), 1.0] for a, b in rng.integers(1, N + 1, size=(60, 2)) if a != b] +
                                  [[int(rng.integers(1, N + 1)), f"v{k}_{v}", 1.0] for v in range(7)] for k in range(K)],
                  lambda_Y_graph=1.0)
    model = simulate_problem(M, blocks=blocks, K=K, seed=11, missing=0.3, model_kwargs=mk, **kw)
    eng = P.Engine(model)
    eng.set_loss_grad_kernel(_lib.KERNEL_TC, 0)
    g = eng.loss_grad(include_reg=True)
    eng.reset_opt_state(1e-8)
    h = eng.fit(eng.make_opts(epoch=1, max_epochs=3, lr=0.1, update_X=1, update_Y=1, update_col_layers=1, kernel=_lib.KERNEL_TC,
                              rel_tol=0.0, abs_tol=0.0))
    eng.set_loss_grad_kernel(_lib.KERNEL_FFMA, 0)
    g2 = eng.loss_grad(include_reg=True)
    eng.column_stats()
    if nb:
        eng.batch_stats()
    nbuf, bad = C.c_int64(0), C.c_int64(0)
    rc = eng.lib.pmf_check_guards(C.byref(nbuf), C.byref(bad))
    eng.close()
    print(f"{name}: loss {g['loss']:.6e} (FP32 kernel {g2['loss']:.6e}), |dY| {np.linalg.norm(g['dY']):.4e}, fit losses "
          f"{[round(x, 2) for x in h['loss']]}; guards {'off' if rc != 0 else f'{nbuf.value} buffers, corrupt bytes {bad.value}'}", flush=True)
