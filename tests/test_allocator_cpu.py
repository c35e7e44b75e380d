"""The library's device allocator (csrc/guard.cu: size-keyed cache of freed blocks, guard zones under PMF_GUARD=1)
compiled with g++ against a host-only stand-in for the CUDA runtime (tests/cuda_stub/) and driven by a small
harness: reuse by exact size and device, zero fill on reuse, the 64 MB / 1 GB caps, release, the retry after an
allocation failure, and the guard-zone check."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "cuda_stub")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("alloc") / "alloc_harness")
    cmd = ["g++", "-std=c++17", "-O1", "-x", "c++", "-I", STUB, os.path.join(ROOT, "pathmatfac.jl_b200", "csrc", "guard.cu"),
           os.path.join(STUB, "allocator_harness.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def _run(binary, args=(), **env):
    e = {k: v for k, v in os.environ.items() if k not in ("PMF_GUARD", "PMF_ALLOC_CACHE")}
    e.update(env)
    r = subprocess.run([binary, *args], capture_output=True, text=True, timeout=120, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return {k: int(v) for k, v in (line.split() for line in r.stdout.splitlines() if line.strip())}


def test_cache_reuses_blocks_by_size_and_device(harness):
    o = _run(harness)
    assert o["mallocs_after_three"] == 3
    assert o["frees_after_cached_free"] == 0 and o["device_syncs_after_cached_free"] == 1   # parked after a device sync
    assert o["reused_same_block"] == 1 and o["reused_block_is_zero"] == 1 and o["mallocs_after_reuse"] == 3
    assert o["other_size_is_fresh"] == 1 and o["other_device_is_fresh"] == 1
    assert o["big_block_freed_at_once"] == 1                       # > 64 MB: pmf_abi.cu's own two-slot pool
    assert o["frees_beyond_the_cap"] == 3                          # 20 x 60 MB against the 1 GB cap
    assert o["retry_after_oom_ok"] == 1 and o["parked_blocks_released_on_oom"] == 1
    assert o["live_bytes_after_release"] == 0 and o["mallocs_equal_frees"] == 1 and o["free_null_ok"] == 1


def test_cache_can_be_switched_off(harness):
    o = _run(harness, PMF_ALLOC_CACHE="0")
    assert o["frees_after_cached_free"] == 1 and o["device_syncs_after_cached_free"] == 0 and o["mallocs_after_reuse"] == 4
    assert o["live_bytes_after_release"] == 0 and o["mallocs_equal_frees"] == 1


def test_guard_zones_catch_overruns_on_either_side(harness):
    o = _run(harness, args=("g",), PMF_GUARD="1")
    assert o["guard_rc"] == 0 and o["guard_buffers"] == 3 and o["guard_bad"] == 0
    assert o["guard_bad_after_overrun"] == 2 and o["frees_in_guard_mode"] == 3
