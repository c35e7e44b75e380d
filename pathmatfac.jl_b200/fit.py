"""The drop-in boundary: ``mf_fit`` / ``mf_fit_adapt_lr`` with the reference's keyword
surface (src/fit.jl:9-75), executing on a libpmf handle (``Engine``).  ``gpu(model)`` /
``cpu(model)`` mirror the reference's device placement (fit_matfac.jl:325-340)."""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import (KERNEL_AUTO, TERM_CODES, c_double_p, c_float_p, check, fptr, iptr, pmf_dims, pmf_fit_opts,
                   pmf_history, pmf_losses)
from .layers import BatchScale, BatchShift, FrozenLayer, Identity
from .regularizers import (ARDRegularizer, BatchArrayReg, ColParamReg, CompositeRegularizer,
                           FeatureSetARDReg, FrozenRegularizer, GroupRegularizer, L2Regularizer,
                           NetworkRegularizer, SelectiveL1Reg, ZeroReg)
from .util import DIST_CODE


_REG_TYPES = (ARDRegularizer, BatchArrayReg, ColParamReg, CompositeRegularizer, FeatureSetARDReg, FrozenRegularizer,
              GroupRegularizer, L2Regularizer, NetworkRegularizer, SelectiveL1Reg, ZeroReg)


def _f32(a, order="C"):
    return np.ascontiguousarray(a, dtype=np.float32) if order == "C" else np.asfortranarray(a, dtype=np.float32)


def _jl(a):
    """numpy (rows, cols) array -> float32 buffer in the reference's column-major layout."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).T)


def _jl_out(a):
    """Writable [cols][rows] float32 view of a column-major (rows, cols) array -- a download can land in the
    array's own memory -- or None when the array is not laid out that way."""
    t = a.T
    return t if (a.dtype == np.float32 and t.flags["C_CONTIGUOUS"] and t.flags["WRITEABLE"]) else None


def recognise_closure(r, K):
    """The stage functions of the reference install anonymous functions as regularisers: ``X -> 0.5f0*sum(X.*X)``
    (src/fit.jl:686), ``0.05 * sum(x.^2)``-style quadratic closures (:266-267) and the zero closures ``X -> 0f0`` /
    ``y -> 0`` / ``x -> 0.0`` (:415, :769, src/transform.jl:61,70).  A closure has no structure to marshal, so it is
    identified by evaluating it on probes: identically zero -> no penalty; 0.5 w sum(x^2) with one scalar w -> an
    L2Regularizer with uniform weight w.  Anything else is an ERROR -- a penalty must never be dropped silently.
    Returns None (zero) or the uniform L2 weight."""
    rng = np.random.default_rng(12345)
    probes = [np.ones((K, 3), np.float32), rng.standard_normal((K, 5)).astype(np.float32),
              (3.0 * rng.standard_normal((K, 2))).astype(np.float32)]
    try:
        vals = [float(r(p)) for p in probes]
    except Exception as e:  # noqa: BLE001
        raise _lib.PmfError(f"regulariser closure could not be evaluated on a K x n probe: {e!r}")
    if all(v == 0.0 for v in vals):
        return None
    ssq = [float((p.astype(np.float64) ** 2).sum()) for p in probes]
    w = vals[0] / (0.5 * ssq[0])
    if w > 0 and all(abs(v - 0.5 * w * q) <= 1e-4 * abs(0.5 * w * q) for v, q in zip(vals, ssq)):
        return w
    raise _lib.PmfError("unsupported regulariser closure: only `x -> 0` and `x -> 0.5*w*sum(x.*x)` (the closures "
                        "src/fit.jl installs) can be moved to the device; use a regulariser object instead")


class AdaGrad:
    """Flux.Optimise.AdaGrad(lr) as built by construct_optimizer (src/fit.jl:41-43).  The
    accumulators live on the device (libpmf handle); a new AdaGrad object means fresh state,
    the same object across LR-halving restarts keeps it (src/fit.jl:55-64)."""

    def __init__(self, eta=1.0, epsilon=1e-8):
        self.eta = float(eta)
        self.epsilon = float(epsilon)
        self._owner = None


class Engine:
    """A PathMatFacModel resident on one GPU (one libpmf handle).  ``rows`` restricts the
    handle to a contiguous block of samples (sample-sharded multi-GPU, see dist.py)."""

    def __init__(self, model, device: int = 0, rows: Optional[range] = None, upload_data=True):
        self.lib = _lib.load()
        self.model = model
        mf = model.matfac
        K, M_all = mf.X.shape
        N = mf.Y.shape[1]
        self.rows = range(0, M_all) if rows is None else rows
        self.M, self.N, self.K = len(self.rows), N, K
        self.h = C.c_void_p()
        self.device = device
        dims = pmf_dims(self.M, N, K, device)
        rc = self.lib.pmf_create(C.byref(dims), C.byref(self.h))
        if rc != 0:
            msg = self.lib.pmf_last_error(None)
            raise _lib.PmfError(f"pmf_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.n_views = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        if upload_data:
            self.push_data(model.data)
        self.push_structure()
        self.push_params()

    # -- helpers ---------------------------------------------------------------------------
    def _ck(self, rc):
        check(self.lib, self.h, rc)

    def close(self):
        if self.h:
            self.lib.pmf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.pmf_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # -- host-driven exchange step (dist.ShardedFit) -----------------------------------------------
    def shared_buffers(self):
        """(gradients, scalars): torch tensors that alias the handle's shared gradient buffer
        [dY | dlogsigma | dmu | dlogdelta | dtheta] (float32) and its two rank-local loss scalars (float64)."""
        import torch

        class _DeviceBuffer:   # raw device pointer -> torch through __cuda_array_interface__
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = torch.device("cuda", self.device)
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.pmf_shared_grad_buffer(self.h, C.byref(p), C.byref(n)))
        grads = torch.as_tensor(_DeviceBuffer(p.value, n.value, "<f4"), device=dev)
        self._ck(self.lib.pmf_shared_scalar_buffer(self.h, C.byref(p), C.byref(n)))
        scalars = torch.as_tensor(_DeviceBuffer(p.value, n.value, "<f8"), device=dev)
        return grads, scalars

    def torch_stream(self):
        """A torch stream of the handle's device that the library runs on from now on, to be made current around
        every collective (``with torch.cuda.stream(s)``): torch.distributed orders a collective against the CURRENT
        stream.  The legacy default stream (handle 0) cannot be used -- ``pmf_set_stream(h, NULL)`` selects the handle's
        own non-blocking stream, which torch knows nothing about -- so a dedicated stream is created."""
        import torch
        s = torch.cuda.Stream(device=self.device)
        self.set_stream(s.cuda_stream)
        self._torch_stream = s          # keep it alive as long as the handle uses it
        return s

    # -- uploads ---------------------------------------------------------------------------
    def push_data(self, D):
        """model.data (M x N).  A Fortran-ordered float32 array is passed without a copy."""
        D = np.asarray(D)[self.rows.start:self.rows.stop, :]
        buf = _jl(D)        # [N][M]; no copy when D is float32 and column-major
        self.h2d_bytes += buf.nbytes
        self._ck(self.lib.pmf_set_data(self.h, fptr(buf)))

    def push_data_buffer(self, buf_NM: np.ndarray):
        """Already marshalled [N][M] float32 buffer (e.g. pinned host memory)."""
        assert buf_NM.shape == (self.N, self.M) and buf_NM.dtype == np.float32
        self.h2d_bytes += buf_NM.nbytes
        self._ck(self.lib.pmf_set_data(self.h, fptr(buf_NM)))

    def _batch_arrays(self):
        ct = self.model.matfac.col_transform
        l2, l4 = ct.unwrapped(1), ct.unwrapped(3)
        return (l2.logdelta if isinstance(l2, BatchScale) else None,
                l4.theta if isinstance(l4, BatchShift) else None)

    def push_structure(self):
        """Everything that is not a trainable value: noise ranges/weights, batch layout,
        regularisers, frozen masks."""
        mf = self.model.matfac
        nm = mf.noise_model
        nr = len(nm.col_ranges)
        cs = np.array([r.start for r in nm.col_ranges], np.int32)
        ce = np.array([r.stop for r in nm.col_ranges], np.int32)
        dc = np.array([DIST_CODE[n.dist] for n in nm.noises], np.int32)
        th = np.zeros((nr, 4), np.float32)
        for i, n in enumerate(nm.noises):
            if n.ext_thresholds is not None:
                th[i, :] = n.ext_thresholds
        w = nm.weights()
        self._ck(self.lib.pmf_set_noise(self.h, nr, iptr(cs), iptr(ce), iptr(dc), fptr(th), fptr(w)))
        # batch layout (logdelta and theta share it)
        ld, thb = self._batch_arrays()
        ba = ld if ld is not None else thb
        if ba is not None and len(ba.col_ranges) > 0:
            nv = len(ba.col_ranges)
            bcs = np.array([r.start for r in ba.col_ranges], np.int32)
            bce = np.array([r.stop for r in ba.col_ranges], np.int32)
            nb = np.array([v.shape[0] for v in ba.values], np.int32)
            bos = np.ascontiguousarray(np.stack([b[self.rows.start:self.rows.stop] for b in ba.batch_index]).astype(np.int32))
            self._ck(self.lib.pmf_set_batch_layout(self.h, nv, iptr(bcs), iptr(bce), iptr(nb), iptr(bos)))
            self.n_views = nv
        else:
            self._ck(self.lib.pmf_set_batch_layout(self.h, 0, None, None, None, None))
            self.n_views = 0
        self.push_regs()

    def push_regs(self):
        mf = self.model.matfac
        self._install_factor_reg(0, mf.X_reg)
        self._install_factor_reg(1, mf.Y_reg)
        # layer regularisers + frozen masks
        ct = mf.col_transform
        frozen_layers = 0
        for s, l in enumerate(ct.layers):
            if isinstance(l, FrozenLayer):
                frozen_layers |= 1 << s
        frozen_regs = 0
        regs = mf.col_transform_reg.regs if mf.col_transform_reg is not None else [ZeroReg()] * 4
        for s, r in enumerate(regs):
            slot = s + 1
            if isinstance(r, FrozenRegularizer):
                frozen_regs |= 1 << s
                r = r.reg
            if isinstance(r, ColParamReg):
                wv, cv = r.expanded(self.N)
                self._ck(self.lib.pmf_set_layer_reg_col(self.h, slot, fptr(wv), fptr(cv)))
            elif isinstance(r, BatchArrayReg) and self.n_views > 0:
                wv = _f32(np.concatenate(r.weights))
                cv = _f32(np.concatenate(r.centers))
                self._ck(self.lib.pmf_set_layer_reg_batch(self.h, slot, fptr(wv), fptr(cv)))
            elif slot in (1, 3):
                self._ck(self.lib.pmf_set_layer_reg_col(self.h, slot, None, None))
            elif self.n_views > 0:
                self._ck(self.lib.pmf_set_layer_reg_batch(self.h, slot, None, None))
        self._ck(self.lib.pmf_set_frozen(self.h, frozen_layers, frozen_regs))
        w = mf.noise_model.weights()
        # weights can change between calls (reweight_col_losses!, src/fit.jl:151-187)
        nm = mf.noise_model
        nr = len(nm.col_ranges)
        cs = np.array([r.start for r in nm.col_ranges], np.int32)
        ce = np.array([r.stop for r in nm.col_ranges], np.int32)
        dc = np.array([DIST_CODE[n.dist] for n in nm.noises], np.int32)
        th = np.zeros((nr, 4), np.float32)
        for i, n in enumerate(nm.noises):
            if n.ext_thresholds is not None:
                th[i, :] = n.ext_thresholds
        self._ck(self.lib.pmf_set_noise(self.h, nr, iptr(cs), iptr(ce), iptr(dc), fptr(th), fptr(w)))

    def _install_factor_reg(self, which, reg):
        lib, h = self.lib, self.h
        self._ck(lib.pmf_clear_reg(h, which))
        n = self.M if which == 0 else self.N
        lo = self.rows.start if which == 0 else 0

        def install(r, p):
            if r is None or isinstance(r, ZeroReg) or p == 0.0:
                return
            if callable(r) and not isinstance(r, _REG_TYPES):
                w = recognise_closure(r, self.K)       # raises on anything but the reference's own closures
                if w is not None:
                    self._ck(lib.pmf_set_reg_l2(h, which, fptr(np.full(self.K, w, np.float32)), p))
                return
            if isinstance(r, L2Regularizer):
                self._ck(lib.pmf_set_reg_l2(h, which, fptr(_f32(r.weights)), p))
            elif isinstance(r, GroupRegularizer):
                st, en, ws = [], [], []
                for rng, w in zip(r.group_idx, r.group_weights):
                    a, b = max(rng.start - lo, 0), min(rng.stop - lo, n)
                    if a < b:
                        st.append(a), en.append(b), ws.append(w)
                if st:
                    self._ck(lib.pmf_set_reg_group(h, which, len(st), iptr(np.array(st, np.int32)),
                                                   iptr(np.array(en, np.int32)), fptr(_f32(np.stack(ws))), p))
            elif isinstance(r, SelectiveL1Reg):
                idx = np.ascontiguousarray(r.l1_idx.T.astype(np.uint8))
                self._ck(lib.pmf_set_reg_sel_l1(h, which, idx.ctypes.data_as(_lib.c_uint8_p), fptr(_f32(r.weight)), p))
            elif isinstance(r, ARDRegularizer):
                st = np.array([x.start for x in r.col_ranges], np.int32)
                en = np.array([x.stop for x in r.col_ranges], np.int32)
                self._ck(lib.pmf_set_reg_ard(h, which, len(st), iptr(st), iptr(en),
                                             fptr(_f32(np.array(r.alpha))), fptr(_f32(np.array(r.beta)))))
            elif isinstance(r, FeatureSetARDReg):
                self._ck(lib.pmf_set_reg_fsard(h, which, fptr(_f32(r.alpha)), fptr(_jl(r.beta))))
            elif isinstance(r, NetworkRegularizer):
                if which == 0 and (self.rows.start != 0 or self.M != self.model.matfac.X.shape[1]):
                    raise _lib.PmfError("a network regulariser on X couples samples: not shardable (replicas only)")
                self._install_network(which, r, p)
            else:
                raise _lib.PmfError(f"unsupported regulariser {type(r).__name__}")

        if isinstance(reg, CompositeRegularizer):
            for r, p in zip(reg.regularizers, reg.mixture_p):
                install(r, float(p))
        else:
            install(reg, 1.0)

    def _install_network(self, which, r: NetworkRegularizer, p):
        def cat(mats):
            rp = np.concatenate([sp.csr_matrix(m).indptr.astype(np.int32) for m in mats])
            ci = np.concatenate([sp.csr_matrix(m).indices.astype(np.int32) for m in mats] + [np.zeros(0, np.int32)])
            va = np.concatenate([sp.csr_matrix(m).data.astype(np.float32) for m in mats] + [np.zeros(0, np.float32)])
            return np.ascontiguousarray(rp), np.ascontiguousarray(ci), np.ascontiguousarray(va)
        for m in list(r.AA) + list(r.AB) + list(r.BB):
            m.sort_indices()
        aa, ab, bb = cat(r.AA), cat(r.AB), cat(r.BB)
        nv = np.array([b.shape[0] for b in r.BB], np.int32)
        xv = _f32(np.concatenate(list(r.x_virtual) + [np.zeros(0, np.float32)]))
        dummy_i, dummy_f = np.zeros(1, np.int32), np.zeros(1, np.float32)
        P = lambda a, d: a if a.size else d
        self._ck(self.lib.pmf_set_reg_network(
            self.h, which, iptr(nv), iptr(aa[0]), iptr(P(aa[1], dummy_i)), fptr(P(aa[2], dummy_f)),
            iptr(ab[0]), iptr(P(ab[1], dummy_i)), fptr(P(ab[2], dummy_f)),
            iptr(bb[0]), iptr(P(bb[1], dummy_i)), fptr(P(bb[2], dummy_f)),
            fptr(P(xv, dummy_f)), p, r.cg_rtol, r.cg_atol, r.cg_itmax))

    def push_params(self):
        mf = self.model.matfac
        X = _jl(mf.X[:, self.rows.start:self.rows.stop])
        Y = _jl(mf.Y)
        self._ck(self.lib.pmf_set_factors(self.h, fptr(X), fptr(Y)))
        ct = mf.col_transform
        self._ck(self.lib.pmf_set_col_params(self.h, fptr(_f32(ct.unwrapped(0).logsigma)), fptr(_f32(ct.unwrapped(2).mu))))
        ld, th = self._batch_arrays()
        for v in range(self.n_views):
            a = _jl(ld.values[v]) if ld is not None else None
            b = _jl(th.values[v]) if th is not None else None
            self._ck(self.lib.pmf_set_batch_values(self.h, v, fptr(a), fptr(b)))
        self.h2d_bytes += X.nbytes + Y.nbytes + 8 * self.N

    # -- downloads --------------------------------------------------------------------------
    def pull_params(self):
        """Write the fitted parameters back into the host model (in place)."""
        mf = self.model.matfac
        x_own, y_own = _jl_out(mf.X[:, self.rows.start:self.rows.stop]), _jl_out(mf.Y)
        X = x_own if x_own is not None else np.empty((self.M, self.K), np.float32)
        Y = y_own if y_own is not None else np.empty((self.N, self.K), np.float32)
        self._ck(self.lib.pmf_get_factors(self.h, fptr(X), fptr(Y)))
        if x_own is None:
            mf.X[:, self.rows.start:self.rows.stop] = X.T
        if y_own is None:
            mf.Y[...] = Y.T
        ls = np.empty(self.N, np.float32)
        mu = np.empty(self.N, np.float32)
        self._ck(self.lib.pmf_get_col_params(self.h, fptr(ls), fptr(mu)))
        ct = mf.col_transform
        ct.unwrapped(0).logsigma[...] = ls
        ct.unwrapped(2).mu[...] = mu
        ld, th = self._batch_arrays()
        for v in range(self.n_views):
            shape = (ld if ld is not None else th).values[v].shape
            a = np.empty(shape[::-1], np.float32)
            b = np.empty(shape[::-1], np.float32)
            self._ck(self.lib.pmf_get_batch_values(self.h, v, fptr(a), fptr(b)))
            if ld is not None:
                ld.values[v][...] = a.T
            if th is not None:
                th.values[v][...] = b.T
        self.d2h_bytes += X.nbytes + Y.nbytes + 8 * self.N
        # ordinal thresholds (trained when update_noise_models; src/fit.jl:14)
        nm = mf.noise_model
        if any(n.ext_thresholds is not None for n in nm.noises):
            th4 = np.empty((len(nm.noises), 4), np.float32)
            self._ck(self.lib.pmf_get_thresholds(self.h, len(nm.noises), fptr(th4)))
            for n, row in zip(nm.noises, th4):
                if n.ext_thresholds is not None:
                    n.ext_thresholds[1:3] = row[1:3]
        for reg in (mf.X_reg, mf.Y_reg):
            nets = [reg] if isinstance(reg, NetworkRegularizer) else (
                [r for r in reg.regularizers if isinstance(r, NetworkRegularizer)]
                if isinstance(reg, CompositeRegularizer) else [])
            for net in nets:
                tot = sum(len(x) for x in net.x_virtual)
                if tot:
                    buf = np.empty(tot, np.float32)
                    which = 0 if reg is mf.X_reg else 1
                    self._ck(self.lib.pmf_get_network_virtual(self.h, which, fptr(buf)))
                    o = 0
                    for x in net.x_virtual:
                        x[...] = buf[o:o + len(x)]
                        o += len(x)

    # -- compute ------------------------------------------------------------------------------
    def set_loss_grad_kernel(self, kernel=KERNEL_AUTO, precision=0):
        self._ck(self.lib.pmf_set_loss_grad_kernel(self.h, kernel, precision))

    def loss_grad(self, include_reg=True) -> Dict:
        """Parity hook: loss components and every gradient at the current parameters."""
        out = pmf_losses()
        dX = np.empty((self.M, self.K), np.float32)
        dY = np.empty((self.N, self.K), np.float32)
        dls = np.empty(self.N, np.float32)
        dmu = np.empty(self.N, np.float32)
        self._ck(self.lib.pmf_loss_grad(self.h, int(include_reg), C.byref(out), fptr(dX), fptr(dY), fptr(dls), fptr(dmu)))
        res = {"loss": out.total, "components": {"data": out.data, "X_reg": out.x_reg, "Y_reg": out.y_reg,
                                                 "layer_reg": out.layer_reg},
               "dX": dX.T.copy(), "dY": dY.T.copy(), "dlogsigma": dls, "dmu": dmu,
               "dlogdelta": [], "dtheta": []}
        ld, th = self._batch_arrays()
        for v in range(self.n_views):
            shape = (ld if ld is not None else th).values[v].shape
            a = np.empty(shape[::-1], np.float32)
            b = np.empty(shape[::-1], np.float32)
            self._ck(self.lib.pmf_get_batch_grads(self.h, v, fptr(a), fptr(b)))
            res["dlogdelta"].append(a.T.copy())
            res["dtheta"].append(b.T.copy())
        nr = len(self.model.matfac.noise_model.col_ranges)
        dthr = np.zeros((nr, 2), np.float32)
        self._ck(self.lib.pmf_get_threshold_grads(self.h, nr, fptr(dthr)))
        res["dthresholds"] = dthr
        return res

    def reset_opt_state(self, epsilon=1e-8):
        self._ck(self.lib.pmf_reset_opt_state(self.h, epsilon))

    def make_opts(self, **kw) -> pmf_fit_opts:
        o = pmf_fit_opts()
        self.lib.pmf_default_fit_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def fit(self, opts: pmf_fit_opts) -> Dict:
        cap = max(1, opts.max_epochs - opts.epoch + 1)
        arrs = [np.zeros(cap, np.float64) for _ in range(5)]
        hist = pmf_history()
        hist.capacity = cap
        (hist.loss_total, hist.loss_data, hist.loss_x_reg, hist.loss_y_reg, hist.loss_layer_reg) = [
            a.ctypes.data_as(c_double_p) for a in arrs]
        self._ck(self.lib.pmf_fit(self.h, C.byref(opts), C.byref(hist)))
        n = hist.n_recorded
        return {"term_code": TERM_CODES[hist.term_code], "epochs": int(hist.epochs),
                "loss": arrs[0][:n].tolist(), "data_loss": arrs[1][:n].tolist(), "X_reg": arrs[2][:n].tolist(),
                "Y_reg": arrs[3][:n].tolist(), "layer_reg": arrs[4][:n].tolist(),
                "device_ms": float(hist.device_ms), "kernel_launches": int(hist.kernel_launches)}

    def set_profiling(self, enable=True):
        self._ck(self.lib.pmf_set_profiling(self.h, int(enable)))

    def get_profile(self):
        """(n bracketed data-pass launches, mean ms, min ms) since set_profiling(True)."""
        n, mean, mn = C.c_int32(0), C.c_float(0), C.c_float(0)
        self._ck(self.lib.pmf_get_profile(self.h, C.byref(n), C.byref(mean), C.byref(mn)))
        return int(n.value), float(mean.value), float(mn.value)

    def column_stats(self):
        """(sum_i (dl/dz)^2, count of finite entries) per column -- MF.batched_column_ssq_grads
        / MF.column_nonnan (src/fit.jl:166, :140)."""
        ssq = np.empty(self.N, np.float32)
        cnt = np.empty(self.N, np.float32)
        self._ck(self.lib.pmf_column_stats(self.h, fptr(ssq), fptr(cnt)))
        return ssq, cnt


    def link_col_sqerr(self):
        """(sum_i (D_ij - forward_ij)^2 over finite entries, count of finite entries) per column --
        MF.link_col_sqerr / MF.column_nonnan (src/fit.jl:138-140, :444-447), identity link."""
        sq = np.empty(self.N, np.float32)
        cnt = np.empty(self.N, np.float32)
        self._ck(self.lib.pmf_link_col_sqerr(self.h, fptr(sq), fptr(cnt)))
        return sq, cnt

    def batch_stats(self):
        """Per batched view the (n_b, N_v) tables ba_map(isfinite, theta, data) and
        ba_map(sqerr_func, theta, model, data) (src/batch_array.jl:320-334, src/fit.jl:332,355-356)."""
        ba = self.model.matfac.col_transform.unwrapped(3).theta if self.n_views else None
        if ba is None:
            return [], []
        shapes = [v.shape for v in ba.values]                      # (n_b, N_v)
        cnt = [np.empty((nv, nb), np.float32) for nb, nv in shapes]  # device layout: [column][n_b]
        sq = [np.empty((nv, nb), np.float32) for nb, nv in shapes]
        pc = (c_float_p * len(cnt))(*[fptr(a) for a in cnt])
        ps = (c_float_p * len(sq))(*[fptr(a) for a in sq])
        self._ck(self.lib.pmf_batch_stats(self.h, len(cnt), pc, ps))
        return [a.T.copy() for a in cnt], [a.T.copy() for a in sq]


# ---- device placement (the reference's gpu(model) / cpu(model)) -----------------------------------

def gpu(model, device: Optional[int] = None):
    """``gpu(model)``.  The device ordinal is remembered on the model (``model._device``): the transient handles of the
    staging passes (statistics, M-estimates, EM) are created on the same GPU."""
    if device is not None:
        model._device = int(device)
    if model._engine is None:
        model._engine = Engine(model, device=getattr(model, "_device", 0))
    return model


def cpu(model):
    if model._engine is not None:
        model._engine.pull_params()
        model._engine.close()
        model._engine = None
    return model


def plan_batch_orders(M, N, col_ranges, n_batches, batch_of_sample) -> Dict:
    """Host-side preview of the batch-layer layout of the tcgen05 data pass (``pmf_plan_batch_orders``; no device
    needed): sample orders, passes and per-chunk batches.  ``col_ranges``: [(start, stop)] 0-based per batched
    view, ``batch_of_sample``: [n_views][M] int."""
    lib = _lib.load()
    V = len(col_ranges)
    cs = np.array([r[0] for r in col_ranges], np.int32)
    ce = np.array([r[1] for r in col_ranges], np.int32)
    nb = np.asarray(n_batches, np.int32)
    bos = np.ascontiguousarray(np.asarray(batch_of_sample, np.int32).reshape(V, M))
    n = (C.c_int32 * 3)()
    p0, p1, p2 = (C.cast(C.byref(n, 4 * k), _lib.c_int32_p) for k in range(3))

    def call(*bufs):
        rc = lib.pmf_plan_batch_orders(M, N, V, iptr(cs), iptr(ce), iptr(nb), iptr(bos), p0, p1, p2, *bufs)
        if rc != 0:
            msg = lib.pmf_last_error(None)
            raise _lib.PmfError(msg.decode() if msg else f"pmf_plan_batch_orders failed ({rc})")
    call(None, None, 0, None, None, 0, None, 0)
    n_orders, n_pass, n_pos = int(n[0]), int(n[1]), int(n[2])
    view_order = np.zeros(V, np.int32)
    perm = np.zeros((n_orders, n_pos), np.int32)
    pass_feat0 = np.zeros(n_pass, np.int32)
    pass_order = np.zeros(n_pass, np.int32)
    chunk_batch = np.zeros((V, n_pos // 16), np.uint16)
    call(iptr(view_order), iptr(perm), perm.size, iptr(pass_feat0), iptr(pass_order), n_pass,
         chunk_batch.ctypes.data_as(C.POINTER(C.c_uint16)), chunk_batch.size)
    return {"n_orders": n_orders, "n_pass": n_pass, "n_pos": n_pos, "view_order": view_order, "perm": perm,
            "pass_feat0": pass_feat0, "pass_order": pass_order, "chunk_batch": chunk_batch}


# ---- staging passes that bracket the hot loop -----------------------------------

def _with_zero_factors(model, fn):
    """Run ``fn(engine)`` with X and Y temporarily set to zero on the device, as init_logsigma! /
    reweight_col_losses! do (src/fit.jl:133-136,160-163); the host model is not touched."""
    transient = model._engine is None
    eng = Engine(model, device=getattr(model, "_device", 0)) if transient else model._engine
    try:
        if not transient:
            eng.push_structure()
            eng.push_params()
        mf = model.matfac
        zx = np.zeros((eng.M, eng.K), np.float32)
        zy = np.zeros((eng.N, eng.K), np.float32)
        eng._ck(eng.lib.pmf_set_factors(eng.h, fptr(zx), fptr(zy)))
        out = fn(eng)
        if not transient:
            eng.push_params()          # restore X, Y on the device
        return out
    finally:
        if transient:
            eng.close()


def compute_M_estimates(model, lr=0.1, max_epochs=500, rel_tol=1e-5, abs_tol=1e-3, device=None):
    """``MF.compute_M_estimates`` as ``init_mu!`` calls it (src/fit.jl:88-92): per column, the shift that
    minimises the column's noise-model loss.  It is the fit loop on a model reduced to its ColShift layer:
    X = Y = 0, sigma = 1, batch parameters zero, every other layer and every penalty frozen, mu started at 0
    -- the same fused data pass, 2 bytes of parameter traffic per column.  Returns (M_estimates, history);
    the host model is not touched."""
    transient = model._engine is None
    eng = Engine(model, device=device if device is not None else getattr(model, "_device", 0)) if transient else model._engine
    try:
        if not transient:
            eng.push_structure()
        zx = np.zeros((eng.M, eng.K), np.float32)
        zy = np.zeros((eng.N, eng.K), np.float32)
        zn = np.zeros(eng.N, np.float32)
        eng._ck(eng.lib.pmf_set_factors(eng.h, fptr(zx), fptr(zy)))
        eng._ck(eng.lib.pmf_set_col_params(eng.h, fptr(zn), fptr(zn)))
        ld, th = eng._batch_arrays()
        for v in range(eng.n_views):
            z = np.zeros((ld if ld is not None else th).values[v].shape[::-1], np.float32)
            eng._ck(eng.lib.pmf_set_batch_values(eng.h, v, fptr(z), fptr(z)))
        eng._ck(eng.lib.pmf_set_frozen(eng.h, 0b1011, 0xFFFFFFFF))     # only ColShift (slot 3 of 4) trains
        eng.reset_opt_state(1e-8)
        o = eng.make_opts(max_epochs=int(max_epochs), epoch=1, lr=float(lr), adagrad_eps=1e-8, rel_tol=float(rel_tol),
                          abs_tol=float(abs_tol), update_X=0, update_Y=0, update_col_layers=1)
        h = eng.fit(o)
        est = np.empty(eng.N, np.float32)
        eng._ck(eng.lib.pmf_get_col_params(eng.h, None, fptr(est)))
        if not transient:
            eng.push_structure()       # frozen masks, penalties
            eng.push_params()
            eng.reset_opt_state(1e-8)
        return est, h
    finally:
        if transient:
            eng.close()


def init_mu(model, lr_mu=0.1, max_epochs=500, history=None, **kwargs):
    """``init_mu!`` (src/fit.jl:82-104)."""
    est, h = compute_M_estimates(model, lr=lr_mu, max_epochs=max_epochs, **kwargs)
    if history is not None:
        h["name"] = "init_mu"
        history.append(h)
    model.matfac.col_transform.unwrapped(2).mu[...] = est
    if model._engine is not None:
        model._engine.push_params()
    return h


def init_logsigma(model):
    """``init_logsigma!`` (src/fit.jl:125-148): logsigma_j = log sqrt(mean_i (D_ij - forward_ij)^2) with
    X = Y = 0, i.e. the spread of each column around its shift; one streaming pass on the device."""
    sq, cnt = _with_zero_factors(model, lambda eng: eng.link_col_sqerr())
    with np.errstate(divide="ignore", invalid="ignore"):
        col_vars = sq / cnt
        model.matfac.col_transform.unwrapped(0).logsigma[...] = np.log(np.sqrt(col_vars)).astype(np.float32)
    if model._engine is not None:
        model._engine.push_params()


def reweight_col_losses(model):
    """``reweight_col_losses!`` (src/fit.jl:151-187): weights = 1 / (rms_i(dl/dz) * sigma_j), rms over ALL M
    samples (missing entries count as zero gradient, :169-172), non-finite weights -> 1."""
    nm = model.matfac.noise_model
    N = model.matfac.Y.shape[1]
    nm.set_weight(np.ones(N, np.float32))
    ssq, _ = _with_zero_factors(model, lambda eng: eng.column_stats())
    M = model.matfac.X.shape[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        rms = np.sqrt(ssq / np.float32(M)) * np.exp(model.matfac.col_transform.unwrapped(0).logsigma)
        w = (1.0 / rms).astype(np.float32)
    w[~np.isfinite(w)] = 1.0
    nm.set_weight(w)
    if model._engine is not None:
        model._engine.push_structure()


def theta_mom(theta_values):
    """src/fit.jl:297-301: per-batch mean / (corrected) variance across the view's columns."""
    return ([v.mean(axis=1, keepdims=True) for v in theta_values],
            [v.var(axis=1, ddof=1, keepdims=True) for v in theta_values])


def delta2_mom(delta2_values):
    """src/fit.jl:303-311: inverse-gamma (alpha, beta) by the method of moments."""
    f32 = np.float32
    mean = [v.mean(axis=1, keepdims=True) for v in delta2_values]
    var = [v.var(axis=1, ddof=1, keepdims=True) for v in delta2_values]
    alpha = [f32(2) + (m * m) / (v + f32(1e-9)) for m, v in zip(mean, var)]
    beta = [m * (a - f32(1)) for m, a in zip(mean, alpha)]
    return alpha, beta


def _nans_to_val(arrs, val):
    for a in arrs:
        a[~np.isfinite(a)] = val


def theta_delta_em(model, delta2, sigma2, update_priors=True, batch_em_max_iter=100, batch_em_rtol=1e-8):
    """``theta_delta_em`` (src/fit.jl:326-375): empirical-Bayes EM for the batch shifts theta and the batch
    scales delta^2.  The two O(MN) quantities of every iteration -- ba_map(isfinite) once and
    ba_map(sqerr_func) per iteration -- are the device statistics pass (pmf_batch_stats); everything else is
    n_b x N_v arithmetic.  Returns (theta values, delta2) and leaves theta in the model like the reference."""
    f32 = np.float32
    theta = model.matfac.col_transform.unwrapped(3).theta
    sigma2 = np.asarray(sigma2, dtype=f32)
    delta2 = [np.asarray(d, dtype=f32).copy() for d in delta2]
    theta_lsq = [v.astype(f32).copy() for v in theta.values]
    transient = model._engine is None
    eng = Engine(model, device=getattr(model, "_device", 0)) if transient else model._engine
    diffs = []
    try:
        if not transient:
            eng.push_structure()
            eng.push_params()
        batch_sizes, _ = eng.batch_stats()
        for it in range(batch_em_max_iter):
            if update_priors or it == 0:
                theta_mean, theta_var = theta_mom(theta.values)
                alpha, beta = delta2_mom(delta2)
            theta_old = [v.copy() for v in theta.values]
            with np.errstate(divide="ignore", invalid="ignore"):
                new_vals = [((e * d2 * sigma2[None, cr.start:cr.stop] + tl * bs * vt)
                             / (sigma2[None, cr.start:cr.stop] * d2 + bs * vt)).astype(f32)
                            for e, vt, d2, tl, bs, cr in zip(theta_mean, theta_var, delta2, theta_lsq, batch_sizes,
                                                             theta.col_ranges)]
            _nans_to_val(new_vals, f32(0))
            for v, nv in zip(theta.values, new_vals):
                v[...] = nv
            eng.push_params()                                   # the squared errors use the new theta
            _, sqerr = eng.batch_stats()
            _nans_to_val(sqerr, f32(0))
            with np.errstate(divide="ignore", invalid="ignore"):
                delta2 = [((b + f32(0.5) * (sq / sigma2[None, cr.start:cr.stop])) / (a + f32(0.5) * bs - f32(1))).astype(f32)
                          for a, b, sq, bs, cr in zip(alpha, beta, sqerr, batch_sizes, theta.col_ranges)]
            _nans_to_val(delta2, f32(1))
            num = sum(float(((v - o) ** 2).sum()) for v, o in zip(theta.values, theta_old))
            den = sum(float((v * v).sum()) for v in theta.values)
            diffs.append(num / den if den > 0 else float("nan"))
            if diffs[-1] < batch_em_rtol:
                break
    finally:
        if transient:
            eng.close()
    return theta.values, delta2, diffs


# ---- the boundary ----------------------------------------------------------------------------------

def mf_fit(model, *, scale_column_losses=False, update_X=False, update_Y=False, update_row_layers=False,
           update_col_layers=False, update_noise_models=True, reg_relative_weighting=False,
           update_X_reg=False, update_Y_reg=False, update_row_layers_reg=False, update_col_layers_reg=False,
           keep_history=True, opt: Optional[AdaGrad] = None, lr=1.0, max_epochs=1000, epoch=1,
           rel_tol=1e-5, abs_tol=1e-5, capacity=None, verbosity=1, print_prefix="", print_iter=10,
           kernel=KERNEL_AUTO, precision=0, check_every=8, device=None, alternating=False, **kwargs) -> Dict:
    """``mf_fit!`` (src/fit.jl:9-38): one ``MF.fit!`` call on the model.

    Same keyword surface and defaults.  ``capacity`` is accepted and ignored (the fused pass
    never materialises Z).  ``update_noise_models`` (true by default, as in every call of the reference) trains
    the interior thresholds of the ordinal noise models.  ``alternating`` selects the other
    reading of MatFac.jl's epoch (App. D1): column-side step, then the row-side step from a second pass.  A model
    that is not device-resident (``gpu(model)`` not called) is uploaded, fitted and written back within the call --
    the end-to-end path."""
    if scale_column_losses:
        raise NotImplementedError("scale_column_losses=true is never used by the reference (src/fit.jl:9)")
    transient = model._engine is None
    eng = Engine(model, device=device if device is not None else getattr(model, "_device", 0)) if transient else model._engine
    try:
        if not transient:
            eng.push_regs()      # stages swap regularisers / freeze layers between calls
        if opt is None:
            opt = AdaGrad(lr)
        if opt._owner is not eng:
            eng.reset_opt_state(opt.epsilon)
            opt._owner = eng
        t0 = time.time()
        o = eng.make_opts(max_epochs=int(max_epochs), epoch=int(epoch), lr=float(opt.eta),
                          adagrad_eps=float(opt.epsilon), rel_tol=float(rel_tol), abs_tol=float(abs_tol),
                          update_X=int(update_X), update_Y=int(update_Y),
                          update_col_layers=int(update_col_layers), kernel=int(kernel),
                          precision=int(precision), check_every=int(check_every),
                          update_noise_models=int(bool(update_noise_models)), alternating=int(bool(alternating)))
        h = eng.fit(o)
        h["time"] = time.time() - t0
        h["lr"] = opt.eta
        if verbosity > 1:
            print(f"{print_prefix}mf_fit: epochs={h['epochs']} term={h['term_code']} loss={h['loss'][-1] if h['loss'] else None}")
        # The host copy of the parameters is current after every fit, resident model or not: the steps between fits
        # (reweight_eb!, whiten!, the statistics passes' save / restore of X and Y) work on it.
        eng.pull_params()
        if transient:
            opt._owner = None
        h["h2d_bytes"], h["d2h_bytes"] = eng.h2d_bytes, eng.d2h_bytes
        return h
    finally:
        if transient:
            eng.close()


def mf_fit_adapt_lr(model, *, lr=1.0, min_lr=0.001, max_epochs=1000, history=None, keep_history=True,
                    verbosity=1, print_prefix="", **kwargs) -> List[Dict]:
    """``mf_fit_adapt_lr!`` (src/fit.jl:46-75): on "loss_increase" halve the learning rate
    (AdaGrad state kept) and resume at h["epochs"]; stop below ``min_lr``."""
    made_resident = model._engine is None
    if made_resident:
        gpu(model, device=kwargs.get("device"))
    try:
        opt = AdaGrad(lr)
        epoch = 1
        hs = [] if history is None else history
        while epoch <= max_epochs:
            h = mf_fit(model, opt=opt, max_epochs=max_epochs, epoch=epoch, keep_history=True,
                       print_prefix=print_prefix, verbosity=verbosity, **kwargs)
            h["name"] = f"mf_fit_lr={opt.eta}"
            hs.append(h)
            if h["term_code"] == "loss_increase":
                opt.eta *= 0.5
                if opt.eta < min_lr:
                    break
                if verbosity > 0:
                    print(f"{print_prefix}Resuming with smaller learning rate ({opt.eta})")
                epoch = h["epochs"]
            else:
                break
        return hs
    finally:
        if made_resident:
            cpu(model)
