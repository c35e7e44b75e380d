"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with
the keys the driver reads, and the repo arm refuses to run without a CUDA device (there is no CPU path to time)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--M", "300", "--N", "400", "--K", "8"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "iter/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    # ms_per_step is the arm's own (bounded) step, so that steps x ms_per_step is its real timed region; value is in the
    # metric's unit (scaled to the workload's full sample count)
    assert abs(d["value"] * d["ms_per_full_step_extrapolated"] / 1e3 - 1.0) < 1e-9
    assert d["ms_per_step"] <= d["ms_per_full_step_extrapolated"] * (1 + 1e-12)
    assert d["steps"] * d["ms_per_step"] / 1e3 < 60                        # what the driver checks against its own clock
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "samples" in cb["sample"]
    assert d["config"]["workload"].startswith("C2") and "model" not in d["config"]


def test_reference_arm_runs_on_rank_0_only():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--M", "200", "--N", "300", "--K", "8"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_repo_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--M", "256", "--N", "256", "--K", "8", "--no-cpu", "--no-e2e"])
    assert r.returncode != 0 and "CUDA device" in (r.stderr + r.stdout)


@pytest.mark.parametrize("n", [2, 4])
def test_repo_arm_control_flow_at_two_ranks_on_cpu(tmp_path, n):
    """bench.py --gpus N launched exactly as the driver launches it (torch.distributed.run, 127.0.0.1), with the CUDA
    pieces replaced by stand-ins (tests/bench_dryrun_worker.py): rendezvous, parameter broadcast, barriers, max over
    ranks, the end-to-end leg and the ONE JSON line of rank 0."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "bench_dryrun_worker.py"), "--gpus", str(n), "--steps", "4",
           "--warmup", "3", "--M", "96", "--N", "128", "--K", "8"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["n_gpus"] == n and d["steps"] == 4 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["value"] == pytest.approx(n / (d["ms_per_step"] / 1e3)) and d["gpu_launches"] == 9 and d["nccl_collectives"] == 8
    assert d["config"]["parallelism"] == f"sample-sharded x{n}" and d["config"]["per_rank_samples"] == 96
    e = d["e2e"]
    assert e["epochs_run"] == 4 and len(e["seconds_each_call"]) == 3 and e["seconds"] == sorted(e["seconds_each_call"])[1]
    assert e["value"] == pytest.approx(n * 4 / e["seconds"]) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["roofline"]["launches_timed"] == 4 and d["cpu_baseline"] is None      # CPU baseline at N = 1 only
    # the reference arm, launched with the same arguments, prints the SAME `config` (the driver compares them)
    ref = _run(["--impl", "reference", "--gpus", str(n), "--steps", "4", "--warmup", "3", "--M", "96", "--N", "128", "--K", "8"],
               env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": str(n)})
    assert ref.returncode == 0, ref.stderr[-2000:]
    rd = json.loads(ref.stdout.strip().splitlines()[-1])
    assert rd["config"] == d["config"] and rd["metric"] == d["metric"] and rd["unit"] == d["unit"]
    assert rd["steps"] == d["steps"] and rd["warmup"] == d["warmup"] and rd["higher_is_better"] == d["higher_is_better"]
    assert rd["n_gpus"] == d["n_gpus"] and "kernel" in d["arm"] and "note" in rd["arm"]
    assert d["loss_first_last"][1] < d["loss_first_last"][0]
