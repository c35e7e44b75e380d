"""Two ranks on two GPUs (skipped with fewer devices): the sample-sharded fit -- NcclFit (ncclAllReduce inside pmf_fit)
and ShardedFit (host-driven torch.distributed all-reduce) -- against the single-handle fit of the concatenated problem.
Also runs bench.py's own N = 2 set-up, so a change that breaks the multi-GPU launch fails a test, not the round."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(args, n=2, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("mode,kernel,M,N,K", [("nccl", "ffma", 301, 240, 6), ("host", "ffma", 301, 240, 6),
                                                ("nccl", "tc", 2300, 1200, 32)])
def test_two_rank_fit_matches_single_handle(tmp_path, mode, kernel, M, N, K):
    out = tmp_path / "res.json"
    r = _torchrun([os.path.join(ROOT, "tests", "nccl_worker.py"), str(out), mode, kernel, str(M), str(N), str(K)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    full = res["full"]
    assert res["world"] == 2 and res["Y_equal_across_ranks"]
    assert res["term"] == full["term"] and len(res["loss"]) == len(full["loss"]) == 8
    # the sharded sums differ from the single-handle ones only in the order of the float32 atomics
    tol = 1e-6 if kernel == "ffma" else 2e-5
    assert np.max(np.abs(np.array(res["loss"]) / np.array(full["loss"]) - 1)) < tol
    ptol = 2e-4 if kernel == "ffma" else 2e-3
    assert relerr(res["Y"], full["Y"]) < ptol and relerr(res["X"], full["X"]) < ptol
    assert relerr(res["theta"], full["theta"]) < ptol


@needs2
def test_bench_two_gpus_runs():
    """bench.py --gpus 2 exactly as the driver launches it (small shape, few steps)."""
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3", "--M", "2048", "--N", "3072",
                   "--K", "64", "--no-cpu"], timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"] is not None and line["e2e"]["value"] > 0
