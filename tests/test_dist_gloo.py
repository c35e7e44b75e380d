"""World-size-2 gloo test (CPU) of the sample-sharded decomposition: each rank owns a
contiguous block of samples, computes its shard's loss and gradients (here with the CPU oracle,
since there is no GPU), all-reduces the SHARED buffer [dY | dlogsigma | dmu | dtheta | dlogdelta]
and the two rank-local loss scalars, then applies the identical AdaGrad update.  The result must
equal the single-process full-batch step; X and dX stay rank-local."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pmf_oracle as O
from pathmatfac_b200.dist import allreduce_plan_check, shard_plan, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    m, D, meta = O.simulate_model(37, {"mutation": ("bernoulli", 9), "methylation": ("normal", 14)}, K=4, seed=40,
                                  batch_views=["methylation"], n_batches=3, missing=0.25)
    m.X_reg = O.L2Regularizer(4, 1.0)
    m.Y_reg = O.GroupRegularizer(meta["views"], K=4)
    return m, D


def _shard(m, D, rows):
    import copy
    s = copy.deepcopy(m)
    s.X = m.X[:, rows.start:rows.stop].copy()
    for ba in (s.logdelta, s.theta):
        ba.row_batches = [rb[rows.start:rows.stop, :] for rb in ba.row_batches]
    return s, D[rows.start:rows.stop, :]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, D = _problem()
    rows = shard_rows(D.shape[0], rank, world)
    s, Ds = _shard(m, D, rows)
    g = O.data_loss_grads(s, Ds)
    xreg_local = s.X_reg.value(s.X)
    shared = np.concatenate([g["dY"].ravel(), g["dlogsigma"], g["dmu"], g["dtheta"][0].ravel(), g["dlogdelta"][0].ravel()])
    t = torch.from_numpy(shared)
    dist.all_reduce(t)                                 # the one exchange step per epoch
    sc = torch.tensor([g["loss"], xreg_local], dtype=torch.float64)
    dist.all_reduce(sc)
    K, N = m.Y.shape
    dY = t[:K * N].numpy().reshape(K, N) + m.Y_reg.grad(m.Y)     # Y-side penalty applied once, after the reduce
    opt = O.AdaGrad(0.5)
    Y = m.Y.copy()
    opt.apply("Y", Y, dY)
    Xs = s.X.copy()
    opt.apply("X", Xs, g["dX"] + s.X_reg.grad(s.X))                # rank-local
    out[rank] = dict(Y=Y, X=Xs, rows=(rows.start, rows.stop), loss=float(sc[0] + sc[1] + m.Y_reg.value(m.Y)),
                     dX=g["dX"], dY=g["dY"], dmu=g["dmu"], dlogsigma=g["dlogsigma"])
    dist.destroy_process_group()


def test_shard_plan_is_a_partition():
    for M, W in [(10, 3), (7, 8), (10000, 8), (80000, 4)]:
        plan = shard_plan(M, W)
        assert plan[0].start == 0 and plan[-1].stop == M
        assert all(a.stop == b.start for a, b in zip(plan, plan[1:]))
        assert max(len(r) for r in plan) - min(len(r) for r in plan) <= 1


@pytest.mark.timeout(180)
def test_two_rank_sharded_step_equals_full_batch():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    m, D = _problem()
    full = O.data_loss_grads(m, D)
    loss_full = full["loss"] + m.X_reg.value(m.X) + m.Y_reg.value(m.Y)
    opt = O.AdaGrad(0.5)
    Y = m.Y.copy()
    opt.apply("Y", Y, full["dY"] + m.Y_reg.grad(m.Y))
    X = m.X.copy()
    opt.apply("X", X, full["dX"] + m.X_reg.grad(m.X))
    per_rank = [out[r] for r in range(world)]
    assert allreduce_plan_check(per_rank, full)
    for r in per_rank:
        assert np.allclose(r["Y"], Y, rtol=1e-10, atol=1e-12)               # replicas stay identical
        assert np.allclose(r["X"], X[:, r["rows"][0]:r["rows"][1]], rtol=1e-10, atol=1e-12)
        assert np.isclose(r["loss"], loss_full, rtol=1e-12)
