"""The HOST side of the real libpmf on a machine without a GPU.

libpmf is built a second time with `-cudart shared` (PTX only, so the device compiler does not run) and loaded in a fresh
interpreter next to tests/cuda_stub/fake_cudart: a host-only libcudart.so.12 that implements the 33 runtime entry points
the library uses on host memory, logs every kernel launch by name instead of executing it (only `fill_kernel` is carried
out), and records / sanity-checks every TMA descriptor handed to cuTensorMapEncodeTiled.  tests/fake_runtime_worker.py
then drives the library through the Python mirror exactly as on a GPU box.  What this pins on CPU:

  * parameters survive the trip host -> "device" -> host for ragged shapes and padded K (the marshalling of mirror and
    library agree for real, not by reading the header);
  * handle state: an unchanged batch layout keeps batch parameters and AdaGrad accumulators, a changed one
    re-allocates them but keeps the column parameters; `pmf_reset_opt_state` gives epsilon;
  * the kernels an epoch launches per configuration -- 2 per epoch on the FP32 and the tcgen05 path, 4 with batch layers, 6
    for K > 64, the two-pass `alternating` order, the begin / end halves of the host-driven sharded step -- and that
    `kernel_launches` (bench.py's `gpu_launches`) equals the launches actually made;
  * the exchange step inside `pmf_fit`: three ranks as threads over a host-only NCCL stand-in (tests/cuda_stub/fake_nccl) that
    really sums -- one group of two in-place all-reduces per epoch and rank (the whole shared gradient buffer as float32, the
    two rank-local loss scalars as float64), every rank left with the sum, no mismatched call, communicators destroyed;
  * the epoch loop only enqueues: 2 E + 1 launches, E / check_every + 1 synchronisations, no allocation per fit;
  * the `PMF_KERNEL_AUTO` size rule and the refusal of unsupported shapes;
  * launch geometry within the hardware limits, every tensor map within the driver's documented constraints and, whole, inside
    the device block it points into (TMA clamps to the descriptor's extents, not to the allocation);
  * error paths (no device, a device that is not sm_100, allocation failures at several depths, bad arguments) and, after
    every scenario, zero live device blocks, balanced streams / events, no foreign frees, no copy outside an allocation;
  * BASELINE.json's configs at their FULL dimensions (C2, C3, C4 with the 2 M-edge graph regulariser, the C5 shard): the kernels
    `PMF_KERNEL_AUTO` picks there, persistent grids of one CTA per SM, shared memory within 227 KB, every TMA descriptor valid
    at 30 000 / 50 000 features, the launch counter, and the device memory a handle holds."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pathmatfac.jl_b200", "csrc")
STUB = os.path.join(ROOT, "tests", "cuda_stub", "fake_cudart")
SOURCES = ["pmf_abi.cu", "fused_ffma.cu", "reg_update.cu", "fsard.cu", "fused_tc.cu", "wide_tc.cu", "guard.cu"]


@pytest.fixture(scope="module")
def out(tmp_path_factory):
    if shutil.which("nvcc") is None or shutil.which("g++") is None:
        pytest.skip("nvcc / g++ not available")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(shutil.which("nvcc"))), "include")
    d = tmp_path_factory.mktemp("fakert")
    fake, lib = str(d / "libcudart.so.12"), str(d / "libpmf_sharedrt.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", cuda_inc, os.path.join(STUB, "fake_cudart.cpp"),
                        "-Wl,-soname,libcudart.so.12", "-Wl,--version-script=" + os.path.join(STUB, "version.map"), "-o", fake],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    # host code is what is under test: PTX only (no ptxas), -O0
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=compute_100a", "-O0", "-std=c++17", "-cudart", "shared",
                        "-Xcompiler", "-fPIC", "-shared", "-o", lib] + SOURCES + ["-ldl"], cwd=CSRC, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    nccl = str(d / "libnccl.so.2")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread",
                        os.path.join(ROOT, "tests", "cuda_stub", "fake_nccl", "fake_nccl.cpp"), "-Wl,-soname,libnccl.so.2", "-o", nccl],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    env = {k: v for k, v in os.environ.items() if k not in ("PMF_GUARD", "PMF_ALLOC_CACHE", "PMF_LIB", "LD_PRELOAD")}
    # five independent worker processes, run side by side: the scenarios (plain and with PMF_GUARD=1), the exchange step, and
    # BASELINE.json's configs at their full dimensions (two processes of two configs each, ~45 s of host work per process)
    W = os.path.join(ROOT, "tests")
    jobs = {
        "main": ([os.path.join(W, "fake_runtime_worker.py"), fake, lib], env),
        "guarded": ([os.path.join(W, "fake_runtime_worker.py"), fake, lib], dict(env, PMF_GUARD="1")),
        "nccl": ([os.path.join(W, "fake_nccl_worker.py"), fake, nccl, lib], env),
        "full_a": ([os.path.join(W, "fake_runtime_fullsize_worker.py"), fake, lib, "C2", "C3"], env),
        "full_b": ([os.path.join(W, "fake_runtime_fullsize_worker.py"), fake, lib, "C4a", "C5"], env),
    }
    procs = {k: subprocess.Popen([sys.executable] + a, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, cwd=ROOT)
             for k, (a, e) in jobs.items()}
    got = {}
    try:
        for k, pr in procs.items():
            so, se = pr.communicate(timeout=900)
            assert pr.returncode == 0, k + ": " + so[-2000:] + se[-4000:]
            got[k] = json.loads(so.strip().splitlines()[-1])
    finally:
        for pr in procs.values():
            if pr.poll() is None:
                pr.kill()
    res = got["main"]
    res["nccl"], res["guarded"] = got["nccl"], got["guarded"]
    res["fullsize"] = dict(got["full_a"], **got["full_b"])
    return res


def _clean(c):
    return (c["live_blocks"] == 0 and c["live_bytes"] == 0 and c["mallocs"] == c["frees"] and c["host_allocs"] == c["host_frees"]
            and c["streams"] == 0 and c["events"] == 0 and c["bad_frees"] == 0 and c["oob_copies"] == 0)


def test_parameters_round_trip_and_handle_state(out):
    assert out["s1_round_trip"] is True
    assert out["s1_thresholds"] == [[float("-inf"), -1.0, 1.0, float("inf")]]
    assert out["s2_same_layout_keeps_values"] and out["s2_same_layout_keeps_opt_state"]
    assert out["s2_reset_gives_epsilon"]
    assert out["s2_new_layout_keeps_column_params"] and out["s2_new_layout_zeroes_batch_params"]
    assert _clean(out["s2_counters_after_close"]) and out["s2_counters_after_close"]["mallocs"] > 10


def _per_epoch(names, first):
    """The launches after the set-up prefix `first`, cut into epochs of equal content."""
    assert names[:len(first)] == first, names[:6]
    rest = names[len(first):]
    assert len(rest) % 3 == 0
    n = len(rest) // 3
    assert rest[:n] == rest[n:2 * n] == rest[2 * n:], rest
    return rest[:n]


def test_kernels_launched_per_epoch(out):
    s = out["s3_ffma"]
    assert _per_epoch(s["names"], ["multi_pass_kernel"]) == ["data_pass_ffma_kernel<1,0>", "fused_epoch_kernel"]
    s = out["s3_tc"]                                                     # two launches per epoch
    assert _per_epoch(s["names"], ["multi_pass_kernel", "prep_operands_kernel"]) == ["data_pass_tc_kernel<0,0,0>", "fused_epoch_kernel"]
    s = out["s3_tc_batch"]                                               # 4.2: + operand gather and dX combine
    assert _per_epoch(s["names"], ["multi_pass_kernel", "build_a_tc_kernel"]) == [
        "prep_operands_kernel", "data_pass_tc_kernel<0,1,0>", "combine_dx_kernel", "fused_epoch_kernel"]
    s = out["s3_tc_wide"]                                                # 4.3: Z + link, two gradient GEMMs over G'
    assert _per_epoch(s["names"], ["multi_pass_kernel"]) == [
        "prep_wide_kernel", "prep_wide_kernel", "zlink_kernel", "grad_gemm_kernel<0>", "grad_gemm_kernel<1>", "fused_epoch_kernel"]
    s = out["s3_ffma_alternating"]                                       # D1: column-side step, second pass, row-side step
    assert _per_epoch(s["names"], []) == ["data_pass_ffma_kernel<1,0>", "multi_pass_kernel", "control_kernel", "multi_pass_kernel",
                                          "data_pass_ffma_kernel<1,0>", "multi_pass_kernel", "multi_pass_kernel"]
    for k in ("s3_ffma", "s3_tc", "s3_tc_batch", "s3_tc_wide", "s3_ffma_alternating"):
        assert out[k]["reported"] == len(out[k]["names"]), k             # pmf_history.kernel_launches = bench.py's gpu_launches
        assert out[k]["term"] == "max_epochs"
    assert _clean(out["s3_counters_after_close"])


def test_auto_size_rule_and_refusals(out):
    """PMF_KERNEL_AUTO: the tensor-core kernels from M N >= 6e6 (K/64)^2 and min(M, N) >= 2000."""
    assert "data_pass_ffma_kernel<1,0>" in out["s3_auto_small"] and not any("_tc_" in n for n in out["s3_auto_small"])     # 1999 x 3100
    assert "data_pass_tc_kernel<0,0,0>" in out["s3_auto_large"] and not any("ffma" in n for n in out["s3_auto_large"])     # 2000 x 3000
    assert "K > 256" in out["s3_tc_refuses_k300"]


def test_launch_geometry_and_tensor_maps(out):
    for k in ("s3_ffma", "s3_tc", "s3_tc_batch", "s3_tc_wide", "s3_ffma_alternating"):
        for x in out[k]["launches"]:
            threads = x["block"][0] * x["block"][1] * x["block"][2]
            assert 1 <= threads <= 1024 and x["smem"] <= 232448 and min(x["grid"]) >= 1, (k, x)
            if "data_pass_tc_kernel" in x["name"]:
                assert threads == 768 and x["grid"][0] <= 148 and x["smem"] > 200 * 1024       # persistent: at most one CTA per SM
    assert not out["s3_ffma"]["maps"]
    for k in ("s3_tc", "s3_tc_batch", "s3_tc_wide"):
        maps = out[k]["maps"]
        assert maps and all(m["rc"] == 0 and m["rank"] == 2 for m in maps), k            # the driver's documented constraints
        assert all(m["box0"] * (4 if m["dtype"] == 7 else 2) <= 128 for m in maps)        # inner box within the 128-byte swizzle span
        assert {m["dtype"] for m in maps} == {7, 9}                                       # FP32 data / gradients, BF16 operand splits


def test_host_driven_sharded_step(out):
    s = out["s4_sharded"]
    assert s["start"] == [] and s["begin"] == ["data_pass_ffma_kernel<1,0>", "multi_pass_kernel"]      # data pass + rank-local X penalties
    assert s["end"] == ["multi_pass_kernel", "control_kernel", "multi_pass_kernel"]                     # replicated penalties, test, update
    Np, Kp = 128, 8
    assert s["grad_floats"] >= Np * Kp + 2 * Np and s["scalar_doubles"] == 2                            # [dY | dlogsigma | dmu | ...]


def test_error_paths_and_allocation_balance(out):
    assert "no CUDA device" in out["s5_no_device"] and "no CPU path" in out["s5_no_device"]
    assert "compute capability 9" in out["s5_wrong_architecture"] and "sm_100a" in out["s5_wrong_architecture"]
    for f in out["s5_alloc_failures"]:
        assert f["error"] and ("allocation failed" in f["error"] or "out of memory" in f["error"]), f
        assert f["live_blocks"] == 0 and f["bad_frees"] == 0, f                          # nothing leaks when set-up fails half way
    assert out["s5_bad_view"][0] == -1 and "does not exist" in out["s5_bad_view"][1]
    assert out["s5_noise_ranges_must_cover"][0] == -1 and "no noise model" in out["s5_noise_ranges_must_cover"][1]
    assert _clean(out["final_counters"])


def test_exchange_step_inside_pmf_fit_three_ranks(out):
    """SURVEY 8e on CPU: the sample-sharded epoch loop of the real library, ncclAllReduce inside pmf_fit (dist.NcclFit's
    path; GPU: tests/test_gpu_multi.py).  The planted "gradients" (rank r: r + 1) come back as 1 + 2 + 3 on every rank."""
    n = out["nccl"]
    assert n["alive"] == [False, False, False] and n["errors"] == []
    assert n["grad_buffers_all_equal_sum"] and len(set(n["grad_buffer_len"])) == 1
    assert n["scalars"] == [[60.0, 60.0]] * 3                                  # 10 + 20 + 30 in both rank-local scalars
    R, E, L = n["ranks"], n["epochs"], n["grad_buffer_len"][0]
    log = n["nccl_log"]
    assert len(log) == R * E * 2 and all(x["in_place"] == 1 and x["grouped"] == 1 and x["nranks"] == R for x in log)
    for r in range(R):
        mine = [(x["count"], x["dtype"]) for x in log if x["rank"] == r]
        assert mine == [(L, 7), (2, 8)] * E                                      # float32 gradients, then float64 scalars, per epoch
    assert n["nccl_mismatches"] == 0 and n["nccl_live_comms"] == 0
    assert n["term"] == ["max_epochs"] * 3 and len(set(n["kernel_launches"])) == 1
    assert n["live_blocks"] == 0 and n["bad_frees"] == 0 and n["oob_copies"] == 0


def test_exchange_step_at_the_payload_of_the_eight_gpu_config(out):
    """BASELINE configs[4] (80 000 x 50 000, K = 128): the exchange inside pmf_fit on the tcgen05 kernels for K > 64 at the real
    feature count and latent dimension (two ranks; the payload does not depend on the samples per rank).  One group of two
    in-place all-reduces per epoch after the gradient contractions: dY (50 000 x 128) + column-parameter gradients as ONE
    float32 buffer of 26.0 MB, and the two rank-local loss scalars as float64."""
    n = out["nccl"]["c5"]
    assert n["alive"] == [False, False] and n["errors"] == []
    R, E, L = n["ranks"], n["epochs"], n["grad_buffer_len"][0]
    assert n["grad_buffer_len"] == [L] * R and n["N"] * n["K"] + 2 * n["N"] <= L <= 1.02 * (n["N"] * n["K"] + 2 * n["N"])
    assert n["grad_buffers_all_equal_sum"] and n["scalars"] == [[30.0, 30.0]] * R
    log = n["nccl_log"]
    assert len(log) == R * E * 2 and all(x["in_place"] == 1 and x["grouped"] == 1 and x["nranks"] == R for x in log)
    for r in range(R):
        assert [(x["count"], x["dtype"]) for x in log if x["rank"] == r] == [(L, 7), (2, 8)] * E
    # per rank: the first penalty pass, then per epoch 2 operand splits + link + 2 gradient GEMMs + the fused epoch pass, and
    # the two collectives the library's counter also holds (bench.py reports them as nccl_collectives, not gpu_launches)
    assert n["kernel_launches"] == [1 + E * 6 + E * 2] * R and n["term"] == ["max_epochs"] * R
    assert n["nccl_mismatches"] == 0 and n["nccl_live_comms"] == 0
    assert n["live_blocks"] == 0 and n["bad_frees"] == 0 and n["oob_copies"] == 0


def test_graph_regulariser_statistics_passes_and_allocator_cache(out):
    s = out["s3_network"]                                               # NetworkRegularizer on a transposed copy of Y
    assert _per_epoch(s["names"], ["multi_pass_kernel"]) == ["data_pass_ffma_kernel<1,0>", "transpose_in_kernel", "network_virtual_kernel",
                                                             "network_rows_kernel", "transpose_add_kernel", "fused_epoch_kernel"]
    assert s["reported"] == len(s["names"])
    st = out["s3_stats"]                                                # one streaming pass behind each staging statistic (8f rank 1)
    assert st["column_stats"] == st["link_col_sqerr"] == st["batch_stats"] == ["data_pass_ffma_kernel<1,1>"]
    assert st["batch_shapes"] == [[3, 20], [4, 20]]                     # n_b x N_v per batched view, like the reference's BatchArray
    first, second, third = out["s3_mf_fit_calls"]                       # mf_fit on a host-resident model, three times
    assert first["mallocs"] > 10 and first["host_allocs"] == 1
    for later in (second, third):                                       # no cudaMalloc / cudaFree / pinned allocation at all
        assert later["mallocs"] == 0 and later["frees"] == 0 and later["host_allocs"] == 0
        assert later["h2d"] == first["h2d"] > 0 and later["d2h"] == first["d2h"] > 0
    assert _clean(out["s3b_counters"])


def test_guard_zones_survive_every_host_side_copy(out):
    """PMF_GUARD=1 (1 KiB guard zones around every device buffer, csrc/guard.cu): after uploads, downloads, layout changes,
    fits on both kernel families and the statistics passes no guard byte has changed -- no host-side copy of the library
    runs past the size it asked for.  (The kernels' side of the same check runs on the GPU: test_guard_zones_stay_intact.)"""
    g = out["guarded"]["s6_guards"]
    assert g["enabled"] and g["rc"] == 0 and g["buffers"] > 30 and g["bad"] == 0
    assert out["guarded"]["s1_round_trip"] and _clean(out["guarded"]["final_counters"])
    assert out["s6_guards"]["rc"] == -3 and out["s6_guards"]["buffers"] == 0          # not enabled: PMF_ERR_STATE, as documented


def test_shape_sweep_geometry_and_refusals(out):
    """88 fits over ragged / degenerate shapes (1 x 1 to 3000 x 3, K from 1 to 256), with and without batch layers, on both
    kernel families: every launch inside the hardware limits, every tensor map inside the driver's constraints, the launch
    counter equal to the launches made -- and the only refusals are the documented one (tensor-core kernels requested for
    K > 64 WITH batch layers; PMF_KERNEL_AUTO runs those on the FP32 kernel)."""
    sw = out["s7_sweep"]
    assert len(sw) >= 80
    for r in sw:
        if r["error"]:
            assert r["kernel"] == 2 and r["batch_views"] == 2 and r["shape"][2] > 64 and "without batch layers" in r["error"], r
            continue
        assert r["bad_geometry"] == [] and r["bad_maps"] == [] and r["n_launches"] == r["reported"], r
        assert r["tc"] == (r["kernel"] == 2), r                     # an explicit kernel request is honoured, never silently replaced
    assert _clean(out["s7_counters"])


def test_abi_in_the_julia_shims_call_order(out):
    """julia/PathMatFacB200.jl cannot be run here (no Julia), but the ORDER in which it calls the ABI can: create_handle =
    pmf_create, pmf_set_data, pmf_set_batch_layout (before any noise model is known); every mf_fit! = values, regularisers,
    frozen masks, pmf_set_noise last, pmf_reset_opt_state on an optimiser's first use, pmf_fit, read-back.  The library's
    state machine accepts it, twice in a row, and the parameters come back unchanged by the (not executed) kernels."""
    s = out["s8_shim_order"]
    assert [x[:2] for x in s["steps"]] == [["pmf_create", 0], ["pmf_set_data", 0], ["pmf_set_batch_layout", 0]]
    assert len(s["fits"]) == 2
    for f in s["fits"]:
        assert f["ok"] and f["round_trip"] and f["term"] == "max_epochs"
        assert f["names"] == ["multi_pass_kernel", "data_pass_ffma_kernel<1,0>", "fused_epoch_kernel", "data_pass_ffma_kernel<1,0>",
                              "fused_epoch_kernel"]


def test_baseline_configs_at_full_size(out):
    """What the host decides at the sizes BASELINE.json quotes (the GPU suite reaches them only through bench.py and
    scripts/config_times.py): kernel selection, geometry, TMA descriptors, launch counter, device memory per handle."""
    full = out["fullsize"]
    per_epoch = {
        "C2": (["multi_pass_kernel", "prep_operands_kernel"], ["data_pass_tc_kernel<0,0,0>", "fused_epoch_kernel"]),
        "C3": (["multi_pass_kernel", "build_a_tc_kernel"],
               ["prep_operands_kernel", "data_pass_tc_kernel<0,1,0>", "combine_dx_kernel", "fused_epoch_kernel"]),
        "C4a": (["multi_pass_kernel"],
                ["prep_wide_kernel", "prep_wide_kernel", "zlink_kernel", "grad_gemm_kernel<0>", "grad_gemm_kernel<1>",
                 "transpose_in_kernel", "network_virtual_kernel", "network_rows_kernel", "transpose_add_kernel", "fused_epoch_kernel"]),
        "C5": (["multi_pass_kernel"],
               ["prep_wide_kernel", "prep_wide_kernel", "zlink_kernel", "grad_gemm_kernel<0>", "grad_gemm_kernel<1>", "fused_epoch_kernel"]),
    }
    # device bytes per handle after a fit, as a multiple of the data matrix 4 M N: A alone on the fused path, A and
    # its per-view-ordered copy with batch layers, A and G' for K > 64 (+ operand splits, the transposed Y and CG scratch at C4)
    budget = {"C2": 1.05, "C3": 2.2, "C4a": 2.45, "C5": 2.1}
    dims = {"C2": (10000, 30000, 64), "C3": (10000, 30000, 64), "C4a": (10000, 30000, 256), "C5": (10000, 50000, 128)}
    elem = {7: 4, 9: 2}                                               # CU_TENSOR_MAP_DATA_TYPE_FLOAT32 / _BFLOAT16
    for name, (first, epoch) in per_epoch.items():
        r = full[name]
        assert r["error"] is None, (name, r["error"])
        assert (r["M"], r["N"], r["K"]) == dims[name]
        assert r["names"] == first + epoch + epoch, (name, r["names"])
        assert r["reported"] == len(r["launches"]) and r["term"] == "max_epochs"
        for x in r["launches"]:
            threads = x["block"][0] * x["block"][1] * x["block"][2]
            assert min(x["grid"]) >= 1 and x["grid"][1] <= 65535 and threads <= 1024 and x["smem"] <= 232448, (name, x)
            if x["name"].startswith(("data_pass_tc_kernel", "zlink_kernel", "grad_gemm_kernel")):
                assert x["grid"] == [148, 1, 1] and x["smem"] > 190000, (name, x)      # persistent: one CTA per SM
        assert r["maps"] and all(m["rc"] == 0 and m["rank"] == 2 for m in r["maps"]), (name, r["maps"])
        for m in r["maps"]:
            assert m["stride0"] == m["dim0"] * elem[m["dtype"]] and m["box0"] * elem[m["dtype"]] in (64, 128), (name, m)
        M, N, _ = dims[name]
        assert 4 * M * N <= r["live_bytes"] and r["live_bytes_after_fit"] <= budget[name] * 4 * M * N, \
            (name, r["live_bytes_after_fit"] / (4.0 * M * N))
        assert _clean(r["counters_after_close"]), (name, r["counters_after_close"])


def test_epoch_loop_only_enqueues(out):
    """pmf_fit's epoch loop does not wait for the device: termination is evaluated by the fused epoch pass on the device and every
    later kernel checks the stop flag; the host synchronises once per `check_every` epochs to read that flag, and
    once at the end for the history -- so an un-polled fit of 400 epochs waits as often as one of 40 (bench.py's timed region),
    and no fit allocates device or pinned memory from the second call on."""
    rule = {(r["check_every"], r["epochs"]): r for r in out["s9_sync_rule"]}
    for (ce, E), r in rule.items():
        assert r["launches"] == 2 * E + 1, r                            # the first penalty pass, then two launches per epoch
        assert r["host_allocs"] == 0, r
        if ce > 400:
            assert r["syncs"] <= 3, r
        else:
            assert r["syncs"] == E // ce + 1, r
    assert sum(r["mallocs"] for r in out["s9_sync_rule"][2:]) == 0
