"""What the workers of tests/test_abi_on_fake_runtime.py read back from the host-only CUDA runtime stand-in
(tests/cuda_stub/fake_cudart): the launch log, the recorded TMA descriptors, the allocation counters.  `fake` is the loaded
stand-in (set by the worker before the first call)."""
import ctypes as C
import re
import subprocess

fake = None


def launches(clear=True):
    name, dims, sm = C.create_string_buffer(512), (C.c_uint * 6)(), C.c_size_t()
    out = []
    for i in range(fake.fake_launch_count()):
        fake.fake_launch(i, name, 512, dims, C.byref(sm))
        out.append({"name": name.value.decode(), "grid": list(dims)[:3], "block": list(dims)[3:], "smem": sm.value})
    if clear:
        fake.fake_clear_launches()
    return out


def maps():
    v = (C.c_longlong * 10)()
    out = []
    for i in range(fake.fake_map_count()):
        fake.fake_map(i, v)
        out.append(dict(zip(("dtype", "rank", "swizzle", "oob", "rc", "dim0", "dim1", "box0", "box1", "stride0"), list(v))))
    return out


def counters():
    v = (C.c_long * 10)()
    fake.fake_counters(v)
    return dict(zip(("mallocs", "frees", "live_blocks", "live_bytes", "host_allocs", "host_frees", "streams", "events",
                     "bad_frees", "oob_copies"), list(v)))


def short(names):
    """Kernel names without namespaces / hashes / parameter lists: 'data_pass_tc_kernel<0,1,0>'."""
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = []
    for d in dem:
        d = re.sub(r"^void ", "", d)
        d = d.replace("(anonymous namespace)::", "").replace("(bool)", "")
        d = re.sub(r"\(.*$", "", d)
        d = d.replace("pmf::", "").replace("false", "0").replace("true", "1").replace(" ", "")
        out.append(d)
    return out
