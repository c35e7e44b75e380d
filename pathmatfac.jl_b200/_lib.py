"""ctypes binding of libpmf.so (include/pmf.h).  There is no CPU fallback: if the
shared library is missing or no CUDA device is present, every compute entry point
raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMF_LIB") or os.path.join(_HERE, "libpmf.so")   # PMF_LIB: A/B runs of experiment builds

c_int32_p = C.POINTER(C.c_int32)
c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_uint8_p = C.POINTER(C.c_uint8)


class PmfError(RuntimeError):
    pass


class pmf_dims(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("device", C.c_int32)]


class pmf_losses(C.Structure):
    _fields_ = [("data", C.c_double), ("x_reg", C.c_double), ("y_reg", C.c_double),
                ("layer_reg", C.c_double), ("total", C.c_double)]


class pmf_fit_opts(C.Structure):
    _fields_ = [("max_epochs", C.c_int32), ("epoch", C.c_int32), ("lr", C.c_float),
                ("adagrad_eps", C.c_float), ("rel_tol", C.c_double), ("abs_tol", C.c_double),
                ("update_X", C.c_int32), ("update_Y", C.c_int32), ("update_col_layers", C.c_int32),
                ("kernel", C.c_int32), ("precision", C.c_int32), ("check_every", C.c_int32),
                ("no_terminate", C.c_int32), ("update_noise_models", C.c_int32), ("alternating", C.c_int32)]


class pmf_history(C.Structure):
    _fields_ = [("term_code", C.c_int32), ("epochs", C.c_int32), ("n_recorded", C.c_int32),
                ("capacity", C.c_int32), ("loss_total", c_double_p), ("loss_data", c_double_p),
                ("loss_x_reg", c_double_p), ("loss_y_reg", c_double_p), ("loss_layer_reg", c_double_p),
                ("device_ms", C.c_float), ("kernel_launches", C.c_int64)]


TERM_CODES = ["max_epochs", "abs_tol", "rel_tol", "loss_increase", "nonfinite"]
KERNEL_AUTO, KERNEL_FFMA, KERNEL_TC = 0, 1, 2

# name -> (restype, argtypes); every symbol declared in include/pmf.h
H = C.c_void_p
SIGNATURES = {
    "pmf_create": (C.c_int, [C.POINTER(pmf_dims), C.POINTER(H)]),
    "pmf_destroy": (C.c_int, [H]),
    "pmf_release_cached_memory": (C.c_int, []),
    "pmf_plan_batch_orders": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                        c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, C.c_int64, c_int32_p,
                                        c_int32_p, C.c_int32, C.POINTER(C.c_uint16), C.c_int64]),
    "pmf_last_error": (C.c_char_p, [H]),
    "pmf_version": (C.c_char_p, []),
    "pmf_set_stream": (C.c_int, [H, C.c_void_p]),
    "pmf_set_data": (C.c_int, [H, c_float_p]),
    "pmf_set_factors": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_get_factors": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_set_noise": (C.c_int, [H, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_float_p, c_float_p]),
    "pmf_set_col_params": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_get_col_params": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_set_batch_layout": (C.c_int, [H, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "pmf_set_batch_values": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_get_batch_values": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_set_frozen": (C.c_int, [H, C.c_uint32, C.c_uint32]),
    "pmf_clear_reg": (C.c_int, [H, C.c_int32]),
    "pmf_set_reg_l2": (C.c_int, [H, C.c_int32, c_float_p, C.c_float]),
    "pmf_set_reg_group": (C.c_int, [H, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_float_p, C.c_float]),
    "pmf_set_reg_sel_l1": (C.c_int, [H, C.c_int32, c_uint8_p, c_float_p, C.c_float]),
    "pmf_set_reg_ard": (C.c_int, [H, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_float_p, c_float_p]),
    "pmf_set_reg_fsard": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_set_reg_network": (C.c_int, [H, C.c_int32, c_int32_p,
                                      c_int32_p, c_int32_p, c_float_p,
                                      c_int32_p, c_int32_p, c_float_p,
                                      c_int32_p, c_int32_p, c_float_p,
                                      c_float_p, C.c_float, C.c_float, C.c_float, C.c_int32]),
    "pmf_get_network_virtual": (C.c_int, [H, C.c_int32, c_float_p]),
    "pmf_set_layer_reg_col": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_set_layer_reg_batch": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_reset_opt_state": (C.c_int, [H, C.c_float]),
    "pmf_get_opt_state": (C.c_int, [H, C.c_int32, C.c_int32, c_float_p]),
    "pmf_set_opt_state": (C.c_int, [H, C.c_int32, C.c_int32, c_float_p]),
    "pmf_loss_grad": (C.c_int, [H, C.c_int32, C.POINTER(pmf_losses), c_float_p, c_float_p, c_float_p, c_float_p]),
    "pmf_get_batch_grads": (C.c_int, [H, C.c_int32, c_float_p, c_float_p]),
    "pmf_get_threshold_grads": (C.c_int, [H, C.c_int32, c_float_p]),
    "pmf_get_thresholds": (C.c_int, [H, C.c_int32, c_float_p]),
    "pmf_check_guards": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pmf_default_fit_opts": (None, [C.POINTER(pmf_fit_opts)]),
    "pmf_fit": (C.c_int, [H, C.POINTER(pmf_fit_opts), C.POINTER(pmf_history)]),
    "pmf_epoch_begin": (C.c_int, [H, C.POINTER(pmf_fit_opts)]),
    "pmf_epoch_end": (C.c_int, [H, C.POINTER(pmf_fit_opts)]),
    "pmf_fit_start": (C.c_int, [H, C.POINTER(pmf_fit_opts)]),
    "pmf_fit_poll": (C.c_int, [H, C.POINTER(pmf_history), c_int32_p]),
    "pmf_shared_grad_buffer": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pmf_shared_scalar_buffer": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pmf_fsard_update_A": (C.c_int, [H, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_float_p,
                                     c_float_p, c_float_p, c_float_p, C.c_float, C.c_float, C.c_float,
                                     C.c_int32, C.c_int32, C.c_float, c_double_p, c_int32_p]),
    "pmf_get_fsard_beta": (C.c_int, [H, c_float_p]),
    "pmf_column_stats": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_link_col_sqerr": (C.c_int, [H, c_float_p, c_float_p]),
    "pmf_batch_stats": (C.c_int, [H, C.c_int32, C.POINTER(c_float_p), C.POINTER(c_float_p)]),
    "pmf_set_loss_grad_kernel": (C.c_int, [H, C.c_int32, C.c_int32]),
    "pmf_set_profiling": (C.c_int, [H, C.c_int32]),
    "pmf_get_profile": (C.c_int, [H, c_int32_p, c_float_p, c_float_p]),
    "pmf_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "pmf_comm_init_rank": (C.c_int, [H, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "pmf_comm_destroy": (C.c_int, [H]),
}

_lib = None


def load():
    """Load libpmf.so (built in-tree by ``__graft_entry__.build()``) and bind the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a source checkout without the built library: compile it in-tree (nvcc, sm_100a); there is no
        # other execution path, so a failed build is a hard error
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("pmf_build", os.path.join(os.path.dirname(LIB_PATH), "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build(force=False)
        except Exception as exc:   # noqa: BLE001
            raise PmfError(f"{LIB_PATH} is missing and could not be built ({exc}): run "
                           "`python -c 'import __graft_entry__ as g; g.build()'` (libpmf has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "float32 C-contiguous buffer expected"
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_int32_p)


def check(lib, h, rc):
    if rc != 0:
        msg = lib.pmf_last_error(h)
        raise PmfError(f"libpmf error {rc}: {msg.decode() if msg else '?'}")
