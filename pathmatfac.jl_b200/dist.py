"""Sample-sharded multi-GPU fit: one process per GPU, each rank owns a
contiguous block of samples (its rows of the data, its columns of X, its AdaGrad state for X);
Y, the column / batch parameters and the Y-side regulariser data are replicated.  Per epoch
there is exactly one exchange step: an all-reduce (sum) of the shared gradient buffer
[dY | dlogsigma | dmu | dlogdelta | dtheta] and of the two rank-local loss scalars.  On GPUs the
library issues it itself (``NcclFit``: ncclAllReduce over NVLink on the handle's stream, inside
``pmf_fit``, so the epoch loop never returns to the host; ``torch.distributed`` only carries the
128-byte NCCL id at start-up).  ``ShardedFit`` is the same step driven from the host through
``torch.distributed`` collectives (any backend; gloo in the CPU tests of the sharding logic).
Every rank then applies the identical update, so the replicas never diverge.

This mirrors the reference's only spelled-out sharding pattern: row blocks with one Y-gradient
buffer per worker, summed (src/fit_lbfgs.jl:14-54)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from ._lib import TERM_CODES, c_double_p, pmf_history


def shard_rows(M: int, rank: int, world: int) -> range:
    """Contiguous block of samples owned by ``rank`` (balanced to within one sample)."""
    lo = (M * rank) // world
    hi = (M * (rank + 1)) // world
    return range(lo, hi)


def shard_plan(M: int, world: int) -> List[range]:
    return [shard_rows(M, r, world) for r in range(world)]


class ShardedFit:
    """Drives pmf_fit_start / pmf_epoch_begin / all-reduce / pmf_epoch_end on one rank.  The engine supplies
    ``shared_buffers()`` (torch tensors aliasing the handle's shared gradient buffer and its rank-local loss
    scalars) and ``torch_stream()`` (the stream the library runs on, made current around the loop so that every
    collective is ordered with the kernels; None on a CPU engine); the CPU tests drive the same sequence through a
    stub engine over gloo."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.eng = engine
        self.group = group
        self.grads, self.scalars = engine.shared_buffers()
        self.stream = engine.torch_stream()

    def fit(self, opts) -> Dict:
        if self.stream is None:
            return self._fit(opts)
        import torch
        with torch.cuda.stream(self.stream):
            return self._fit(opts)

    def _fit(self, opts) -> Dict:
        eng, lib = self.eng, self.eng.lib
        eng._ck(lib.pmf_fit_start(eng.h, C.byref(opts)))
        check = opts.check_every if opts.check_every > 0 else 8
        stopped = C.c_int32(0)
        since = 0
        for e in range(opts.epoch, opts.max_epochs + 1):
            eng._ck(lib.pmf_epoch_begin(eng.h, C.byref(opts)))
            self.dist.all_reduce(self.grads, op=self.dist.ReduceOp.SUM, group=self.group)
            self.dist.all_reduce(self.scalars, op=self.dist.ReduceOp.SUM, group=self.group)
            eng._ck(lib.pmf_epoch_end(eng.h, C.byref(opts)))
            since += 1
            if since >= check and e < opts.max_epochs:
                since = 0
                eng._ck(lib.pmf_fit_poll(eng.h, None, C.byref(stopped)))
                if stopped.value:
                    break
        cap = max(1, opts.max_epochs - opts.epoch + 1)
        arrs = [np.zeros(cap, np.float64) for _ in range(5)]
        hist = pmf_history()
        hist.capacity = cap
        (hist.loss_total, hist.loss_data, hist.loss_x_reg, hist.loss_y_reg, hist.loss_layer_reg) = [
            a.ctypes.data_as(c_double_p) for a in arrs]
        eng._ck(lib.pmf_fit_poll(eng.h, C.byref(hist), C.byref(stopped)))
        n = hist.n_recorded
        return {"term_code": TERM_CODES[hist.term_code], "epochs": int(hist.epochs), "loss": arrs[0][:n].tolist(),
                "data_loss": arrs[1][:n].tolist(), "X_reg": arrs[2][:n].tolist(), "Y_reg": arrs[3][:n].tolist(),
                "layer_reg": arrs[4][:n].tolist(), "kernel_launches": int(hist.kernel_launches)}


class NcclFit:
    """Attach an NCCL communicator to the engine's handle (one rank per process / GPU); afterwards
    ``engine.fit`` runs the sharded epoch loop entirely inside libpmf.  ``torch.distributed`` (any
    backend) is used once, to broadcast rank 0's ncclUniqueId."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.eng = engine
        lib = engine.lib
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            engine._ck(lib.pmf_comm_unique_id(ident))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        engine._ck(lib.pmf_comm_init_rank(engine.h, world, rank, ident))

    def fit(self, opts) -> Dict:
        return self.eng.fit(opts)

    def close(self):
        self.eng._ck(self.eng.lib.pmf_comm_destroy(self.eng.h))


def allreduce_plan_check(per_rank: List[Dict[str, np.ndarray]], full: Dict[str, np.ndarray]) -> bool:
    """Host-side statement of the exchange step, used by the gloo tests: the shared gradients
    of the row shards must sum to the full-batch gradients, dX must concatenate."""
    ok = True
    for k in ("dY", "dmu", "dlogsigma"):
        ok &= np.allclose(sum(r[k] for r in per_rank), full[k], rtol=1e-5, atol=1e-6)
    ok &= np.allclose(np.concatenate([r["dX"] for r in per_rank], axis=1), full["dX"], rtol=1e-5, atol=1e-6)
    return bool(ok)
