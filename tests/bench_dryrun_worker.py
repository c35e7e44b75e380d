"""Runs bench.py's own main() on CPU ranks (gloo) with stand-ins for the CUDA pieces: torch.cuda calls are no-ops,
events read the host clock, the engine returns canned histories.  What is exercised is bench.py's CONTROL FLOW at
N > 1 -- rendezvous, broadcast of the replicated parameters, barriers, the max-over-ranks reduction, the end-to-end leg,
the JSON line -- so that an edit which breaks the multi-GPU launch fails a CPU test (tests/test_bench_contract.py) and
not the round (round 1: a non-contiguous broadcast made every N > 1 run exit 1)."""
import os
import runpy
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import pathmatfac_b200 as P
from pathmatfac_b200 import dist as pmf_dist

# ---- CUDA stand-ins -------------------------------------------------------------------------------------
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = _Event
torch.Tensor.cuda = lambda self, *a, **k: self
_empty, _tensor, _init = torch.empty, torch.tensor, dist.init_process_group


def _empty_cpu(*a, pin_memory=False, **k):
    return _empty(*a, **k)


def _tensor_cpu(*a, device=None, **k):
    return _tensor(*a, **k)


torch.empty, torch.tensor = _empty_cpu, _tensor_cpu
dist.init_process_group = lambda backend=None, device_id=None, **k: _init("gloo", **k)


class _Engine:
    """Canned engine: every epoch takes 1 ms of host time and lowers the loss by 1 %."""
    n_created = 0

    def __init__(self, model, device=0, rows=None, upload_data=True):
        self.model, self.device = model, device
        self.h2d_bytes = self.d2h_bytes = 0
        self.loss = 1000.0
        self.profiling, self.n_prof = False, 0
        _Engine.n_created += 1

    def make_opts(self, **kw):
        return types.SimpleNamespace(**kw)

    def fit(self, o):
        n = o.max_epochs - o.epoch + 1
        out = []
        for _ in range(n):
            time.sleep(0.001)
            self.loss *= 0.99
            out.append(self.loss)
        if self.profiling:
            self.n_prof += n
        # 2 kernels per epoch + the first penalty pass, and at N > 1 the two all-reduces per epoch the library also counts
        return {"term_code": "max_epochs", "epochs": o.max_epochs, "loss": out,
                "kernel_launches": 2 * n + 1 + (2 * n if dist.is_initialized() and dist.get_world_size() > 1 else 0)}

    def reset_opt_state(self, eps=1e-8):
        self.loss = 1000.0

    def set_profiling(self, on=True):
        self.profiling, self.n_prof = bool(on), 0

    def get_profile(self):
        return self.n_prof, 0.9, 0.8

    def torch_stream(self):
        return None

    def push_data(self, D):
        self.h2d_bytes += np.asarray(D).size * 4

    def push_params(self):
        self.h2d_bytes += self.model.matfac.X.size * 4 + self.model.matfac.Y.size * 4

    def pull_params(self):
        self.d2h_bytes += self.model.matfac.X.size * 4 + self.model.matfac.Y.size * 4

    def close(self):
        pass


class _NcclFit:
    def __init__(self, engine, group=None):
        self.eng = engine
        # what NcclFit does at start-up: one broadcast of the 128-byte id
        t = torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(t, 0)

    def fit(self, o):
        return self.eng.fit(o)


P.Engine = _Engine
pmf_dist.NcclFit = _NcclFit

if __name__ == "__main__":
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
