"""Host time of the library per fit epoch, without a GPU: the real libpmf built with `-cudart shared` on the host-only CUDA runtime
stand-in of the CPU suite (tests/cuda_stub/fake_cudart: launches are logged, not submitted), 2 000 epochs per configuration
(DESIGN.md 6.2).  Build the two libraries the way tests/test_abi_on_fake_runtime.py does, then:

    python scripts/host_cost_per_epoch.py /tmp/fakert/libcudart.so.12 /tmp/fakert/libpmf_sharedrt.so
"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
fake = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL); os.environ["PMF_LIB"] = sys.argv[2]
import numpy as np
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
import fake_runtime_util as U
U.fake = fake
lib = _lib.load()
rng = np.random.default_rng(0)
for (M,N,K,bv) in ((2000,3000,64,0),(2000,3000,64,2),(2000,3000,128,0)):
    D = rng.standard_normal((M,N)).astype(np.float32)
    views = ["a"]*(N//2)+["b"]*(N-N//2)
    batch = {v:[f"b{int(b)}" for b in rng.integers(0,4,M)] for v in ("a","b")} if bv else None
    m = P.PathMatFacModel(D, K=K, feature_views=views, batch_dict=batch, sample_conditions=["c"]*M if bv else None, lambda_X_l2=1.0)
    eng = P.Engine(m); eng.reset_opt_state(1e-8)
    kw = dict(kernel=_lib.KERNEL_TC, lr=0.05, update_X=1, update_Y=1, update_col_layers=1, no_terminate=1, check_every=1<<20, rel_tol=0.0, abs_tol=0.0)
    eng.fit(eng.make_opts(epoch=1, max_epochs=5, **kw)); U.launches()
    E=2000
    t0=time.perf_counter(); h=eng.fit(eng.make_opts(epoch=6, max_epochs=5+E, **kw)); dt=time.perf_counter()-t0
    nm=len(U.maps()); n=len(U.launches())
    print(f"M={M} N={N} K={K} batch_views={bv}: {dt/E*1e6:.1f} us of host time per epoch ({n/E:.1f} launches, tensor maps encoded {nm})")
    eng.close()
