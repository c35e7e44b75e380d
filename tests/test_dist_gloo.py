"""World-size-2 and -4 gloo tests (CPU) of the sample-sharded decomposition: each rank owns a
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
@pytest.mark.parametrize("world", [2, 4])           # 37 samples: uneven shards at both sizes (18 + 19; 9 + 9 + 9 + 10)
def test_two_rank_sharded_step_equals_full_batch(world):
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


# ---- the product's ShardedFit driven end to end over gloo through a stub engine --------------------------
class _StubLib:
    """The five entry points ShardedFit calls (include/pmf.h: pmf_fit_start, pmf_epoch_begin, pmf_epoch_end,
    pmf_fit_poll), served by the CPU oracle on this rank's row shard: begin = shard data loss + gradients + rank-local
    X penalty into the shared buffers, end = replicated penalties, termination test, AdaGrad step.  ShardedFit itself
    (begin -> all-reduce of both buffers -> end, polling, history) is the product code under test."""

    def __init__(self, shard, Ds, grads, scalars):
        self.m, self.D, self.grads, self.scalars = shard, Ds, grads, scalars
        self.K, self.N = shard.Y.shape
        self.opt = O.AdaGrad(1.0)
        self.hist, self.stop, self.term, self.prev, self.epochs = [], 0, 0, None, 0

    def pmf_fit_start(self, h, o):
        self.hist, self.stop, self.term, self.prev = [], 0, 0, None
        return 0

    def pmf_epoch_begin(self, h, o):
        if self.stop:
            return 0
        m = self.m
        g = O.data_loss_grads(m, self.D)
        self.dX = g["dX"] + m.X_reg.grad(m.X)
        shared = np.concatenate([g["dY"].ravel(), g["dlogsigma"], g["dmu"], g["dtheta"][0].ravel(), g["dlogdelta"][0].ravel()])
        self.grads.copy_(torch.from_numpy(shared))
        self.scalars[0], self.scalars[1] = g["loss"], m.X_reg.value(m.X)
        return 0

    def pmf_epoch_end(self, h, o):
        if self.stop:
            return 0
        opts, m, K, N = o._obj, self.m, self.K, self.N
        t = self.grads.numpy()
        dY = t[:K * N].reshape(K, N) + m.Y_reg.grad(m.Y)             # Y-side penalty once, after the reduce
        dls, dmu = t[K * N:K * N + N], t[K * N + N:K * N + 2 * N]
        nb = m.theta.values[0].size
        dth = t[K * N + 2 * N:K * N + 2 * N + nb].reshape(m.theta.values[0].shape)
        dld = t[K * N + 2 * N + nb:].reshape(m.theta.values[0].shape)
        loss = float(self.scalars[0] + self.scalars[1]) + m.Y_reg.value(m.Y)
        self.hist.append(loss)
        self.epochs = opts.epoch + len(self.hist) - 1
        if self.prev is not None:
            d = self.prev - loss
            code = 3 if d < 0 else 1 if abs(d) < opts.abs_tol else 2 if abs(d / loss) < opts.rel_tol else -1
            if code >= 0:
                self.stop, self.term = 1, code
                return 0
        self.prev = loss
        self.opt.eta = opts.lr
        self.opt.apply("X", m.X, self.dX)
        self.opt.apply("Y", m.Y, dY)
        self.opt.apply("logsigma", m.logsigma, dls)
        self.opt.apply("mu", m.mu, dmu)
        self.opt.apply("logdelta0", m.logdelta.values[0], dld)
        self.opt.apply("theta0", m.theta.values[0], dth)
        return 0

    def pmf_fit_poll(self, h, hist, stopped):
        stopped._obj.value = self.stop
        if hist is not None:
            ho = hist._obj
            ho.term_code, ho.epochs, ho.n_recorded, ho.kernel_launches = self.term, self.epochs, len(self.hist), 0
            for i, v in enumerate(self.hist):
                ho.loss_total[i] = v
        return 0


class _StubEngine:
    def __init__(self, shard, Ds):
        K, N = shard.Y.shape
        self._g = torch.zeros(K * N + 2 * N + 2 * shard.theta.values[0].size, dtype=torch.float64)
        self._s = torch.zeros(2, dtype=torch.float64)
        self.lib, self.h = _StubLib(shard, Ds, self._g, self._s), None
        self.stream_calls = 0

    def shared_buffers(self):
        return self._g, self._s

    def torch_stream(self):
        self.stream_calls += 1
        return None

    def _ck(self, rc):
        assert rc == 0


def _no_layer_regs(m):
    m.layer_regs = [O.ZeroReg()] * 4
    return m


def _sharded_worker(rank, world, port, out, epochs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pathmatfac_b200._lib import pmf_fit_opts
    from pathmatfac_b200.dist import ShardedFit
    m, D = _problem()
    _no_layer_regs(m)
    rows = shard_rows(D.shape[0], rank, world)
    s, Ds = _shard(m, D, rows)
    eng = _StubEngine(s, Ds)
    o = pmf_fit_opts()
    o.max_epochs, o.epoch, o.lr, o.rel_tol, o.abs_tol, o.check_every = epochs, 1, 0.25, 1e-12, 1e-12, 2
    h = ShardedFit(eng).fit(o)
    out[rank] = dict(h=h, Y=s.Y, X=s.X, mu=s.mu, theta=s.theta.values[0], rows=(rows.start, rows.stop), stream_calls=eng.stream_calls)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_fit_class_over_gloo_matches_full_batch_fit(world):
    """dist.ShardedFit (the product's host-driven exchange loop) on two gloo ranks, each over a stub engine that
    serves the ABI calls with the oracle on its row shard, against the single-process full-batch oracle fit: same
    loss curve, same term code, replicated parameters identical on both ranks, X equal to the matching columns."""
    epochs = 7
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, _free_port(), out, epochs), nprocs=world, join=True)
    m, D = _problem()
    _no_layer_regs(m)
    href = O.mf_fit(m, D, O.AdaGrad(0.25), max_epochs=epochs, rel_tol=1e-12, abs_tol=1e-12, update_X=True, update_Y=True,
                    update_col_layers=True)
    ranks = [out[r] for r in range(world)]
    for r in ranks:
        assert r["stream_calls"] == 1
        assert r["h"]["term_code"] == href["term_code"] and r["h"]["epochs"] == href["epochs"]
        assert np.allclose(r["h"]["loss"], href["loss"], rtol=1e-10)
        assert np.allclose(r["Y"], m.Y, rtol=1e-9, atol=1e-11) and np.allclose(r["mu"], m.mu, rtol=1e-9, atol=1e-11)
        assert np.allclose(r["theta"], m.theta.values[0], rtol=1e-9, atol=1e-11)
        assert np.allclose(r["X"], m.X[:, r["rows"][0]:r["rows"][1]], rtol=1e-9, atol=1e-11)
    assert all(np.array_equal(ranks[0]["Y"], r["Y"]) for r in ranks[1:])          # replicas never diverge
    assert [r["rows"] for r in ranks] == [(p.start, p.stop) for p in shard_plan(D.shape[0], world)]
