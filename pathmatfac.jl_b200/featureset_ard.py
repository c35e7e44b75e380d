"""Feature-set ARD outer step: host wrappers of ``update_A!`` / ``update_lambda!``
(src/featureset_ard.jl:189-294).  The ISTA loop itself runs on the device
(csrc/fsard.cu) against the handle's current Y."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import c_double_p, c_int32_p, fptr, iptr
from .regularizers import FeatureSetARDReg


def update_lambda(reg: FeatureSetARDReg, Y: np.ndarray):
    """update_lambda! (src/featureset_ard.jl:189-209); tiny, host side."""
    for cr, S, A, opt in zip(reg.col_ranges, reg.S, reg.A, reg.A_opts):
        Yv = Y[:, cr.start:cr.stop].astype(np.float32)
        ms = np.mean(Yv * Yv, axis=1)
        floor = min(ms.min(), reg.v0)
        opt.lam[...] = (A.shape[0] * np.float32(S.sum() / (S.shape[0] * S.shape[1]))) / (ms - floor + np.float32(1e-3))


def update_A(reg: FeatureSetARDReg, model, max_epochs=1000, term_iter=20, atol=1e-5, verbosity=0,
             print_prefix="", print_iter=100):
    """update_A! (src/featureset_ard.jl:278-294) on a device-resident model (``gpu(model)``).
    Returns [(best_loss, epochs)] per view; ``reg.A``, ``reg.beta`` and the optimiser state
    are updated in place, and the device copy of beta is refreshed for the next fit stage."""
    eng = model._engine
    if eng is None:
        raise RuntimeError("update_A needs a device-resident model: call gpu(model) first")
    eng.push_regs()       # installs reg.alpha / reg.beta as the handle's FSARD regulariser
    out = []
    for cr, A, S, opt in zip(reg.col_ranges, reg.A, reg.S, reg.A_opts):
        S = S.tocsr()
        S.sort_indices()
        rp = np.ascontiguousarray(S.indptr.astype(np.int32))
        ci = np.ascontiguousarray(S.indices.astype(np.int32))
        va = np.ascontiguousarray(S.data.astype(np.float32))
        A_buf = np.zeros(A.shape, np.float32)           # [L][K]
        ssq = np.ascontiguousarray(opt.ssq_grad.astype(np.float32))
        lam = np.ascontiguousarray(opt.lam.astype(np.float32))
        best = C.c_double(0.0)
        epochs = C.c_int32(0)
        eng._ck(eng.lib.pmf_fsard_update_A(eng.h, cr.start, cr.stop, A.shape[0], iptr(rp), iptr(ci), fptr(va),
                                           fptr(A_buf), fptr(ssq), fptr(lam), float(opt.lr), float(reg.alpha0),
                                           float(reg.v0), int(max_epochs), int(term_iter), float(atol),
                                           C.byref(best), C.byref(epochs)))
        A[...] = A_buf
        opt.ssq_grad[...] = ssq
        out.append((best.value, int(epochs.value)))
        if verbosity > 0:
            print(f"{print_prefix}    View: final loss {best.value} after {epochs.value} epochs")
    beta = np.empty((eng.N, eng.K), np.float32)
    eng._ck(eng.lib.pmf_get_fsard_beta(eng.h, fptr(beta)))
    reg.beta[...] = beta.T
    return out
