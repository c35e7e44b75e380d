#!/usr/bin/env python
"""Race and deadlock check of the barrier choreography of the tcgen05 kernels -- the fused data pass (csrc/fused_tc.cu) and
the two K > 64 kernels of csrc/wide_tc.cu (`--model zlink`, `--model grad_gemm`) -- at the level of their PROTOCOL.  For
the fused data pass: the 24 warps of a CTA as actors, the 16 mbarrier families with their arrival counts and the parity form of
`mbarrier.try_wait`, the asynchronous engines (TMA loads with complete_tx, tcgen05.mma with tcgen05.commit -- in order per
issuing thread --, TMA reduce-add bulk groups with wait_group.read) and every buffer the roles hand to one another (shared
memory rings, TMEM accumulators, staging buffers), sliced the way the warps slice them.

The pool's compute-sanitizer is closed (profiles/r2_compute_sanitizer_closed.txt), so the kernel itself cannot be run under
racecheck; what can be checked without a GPU is that the ORDER the barriers impose is sufficient:

  * no data race: every pair of conflicting accesses to a buffer slice (write/write, read/write) is ordered by
    happens-before, computed with vector clocks over program order, barrier arrive -> completed phase -> wait, issue ->
    asynchronous operation -> completion (commit / complete_tx / wait_group);
  * no deadlock and no lost phase: under randomly chosen interleavings (and randomly delayed asynchronous completions) every
    actor runs to the end; a waiter uses the real parity test, so a barrier that ran two phases ahead of a waiter shows up.

Each role below is a transcription of its loop in fused_tc.cu (line references in the docstrings); `--mutate NAME` removes
one wait (or one ring stage) to show that the checker notices.  tests/test_tc_protocol.py runs it over item / tile
configurations and seeds, and checks that every mutation is caught.

    python scripts/tc_protocol_check.py [--items 5,1,7,2] [--seeds 200] [--mutate no_z_empty]"""
from __future__ import annotations

import argparse
import random
import sys
from collections import defaultdict, deque

NEPI, NDRAIN = 16, 4
SA, SXK, SXM, SZ, SDX = 3, 3, 1, 4, 2          # fused_tc.cu:62-79 (PMF_SA = 3)
LA = SZ - 1
RLAG = 5 if SDX == 2 else 3                     # :344


class Ring:                                      # struct Ring, fused_tc.cu:184-188
    def __init__(self):
        self.s, self.ph = 0, 0

    def next(self, n):
        self.s += 1
        if self.s == n:
            self.s, self.ph = 0, self.ph ^ 1


class Race(Exception):
    pass


class Deadlock(Exception):
    pass


def join(a, b):
    for k, v in b.items():
        if a.get(k, 0) < v:
            a[k] = v


class Barrier:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.pending = count        # arrivals still missing in the current phase
        self.tx = 0                 # outstanding transaction groups (TMA loads) of the current phase
        self.completed = 0          # phases completed
        self.acc = {}               # clocks of the arrivals of the current phase
        self.clock = {}             # what a waiter of a completed phase acquires

    def parity_passes(self, p):     # mbarrier.try_wait.parity: true once the phase of parity p has completed
        return (self.completed & 1) != p

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            join(self.clock, self.acc)
            self.acc = {}
            self.pending = self.count
            self.completed += 1

    def arrive(self, clock, expect_tx=0):
        if self.pending == 0:
            raise Race(f"{self.name}: arrival on a phase that already has all its arrivals")
        join(self.acc, clock)
        self.pending -= 1
        self.tx += expect_tx
        self._maybe_complete()

    def complete_tx(self, clock):
        join(self.acc, clock)
        self.tx -= 1
        self._maybe_complete()


class Sim:
    def __init__(self, items, seed, mutate=None, ordered_loads=False):
        self.rng = random.Random(seed)
        self.items, self.mutate, self.ordered_loads = items, mutate, ordered_loads
        self.bars = {}
        self.vc = defaultdict(dict)                 # actor -> vector clock
        self.mem = defaultdict(lambda: {"w": None, "r": []})
        self.mma_q = defaultdict(deque)              # tcgen05 pipe, in order per issuing thread
        self.pipe_clock = defaultdict(dict)
        self.loads = []                              # TMA loads in flight (complete in any order)
        self.bulk = defaultdict(deque)               # bulk async-groups per issuing thread (TMA reduce-add / store; reads complete in order)
        self.bulk_done_clock = defaultdict(dict)
        self.n_async = 0
        # odd seeds: every actor and every asynchronous engine gets its own speed (1/50 to 5 times the others), so that
        # schedules in which one role is starved for many tiles, or an engine lags far behind, are explored too
        self.weights = {} if seed & 1 else None
        self.setup()
        for fam, n, cnt in self.barrier_table():
            for i in range(n):
                self.bars[(fam, i)] = Barrier(f"{fam}[{i}]", cnt)

    # ---- happens-before bookkeeping ----------------------------------------------------------------------------------
    def tick(self, actor):
        c = self.vc[actor]
        c[actor] = c.get(actor, 0) + 1
        return c

    def access(self, res, slices, kind, clock, who):
        """`who` = (component, value): the access is identified by that entry of its own clock."""
        for sl in slices:
            m = self.mem[(res, sl)]
            w = m["w"]
            if w is not None and clock.get(w[0][0], 0) < w[0][1]:
                raise Race(f"{kind} of {res}[{sl}] by {who[0]} is not ordered after the write by {w[0][0]}")
            if kind == "write":
                for r in m["r"]:
                    if clock.get(r[0], 0) < r[1]:
                        raise Race(f"write of {res}[{sl}] by {who[0]} is not ordered after the read by {r[0]}")
                m["w"], m["r"] = (who, None), []
            else:
                m["r"].append(who)

    def sync_access(self, actor, res, slices, kind):
        c = self.tick(actor)
        self.access(res, slices, kind, c, (actor, c[actor]))

    # ---- asynchronous engines ------------------------------------------------------------------------------------------
    def issue_mma(self, thread, accesses):
        c = dict(self.tick(thread))
        self.mma_q[thread].append(("op", accesses, c))

    def commit(self, thread, bar):
        self.mma_q[thread].append(("commit", bar, dict(self.tick(thread))))

    def step_pipe(self, thread):
        kind, x, c = self.mma_q[thread].popleft()
        pc = self.pipe_clock[thread]
        join(pc, c)
        name = "pipe:" + thread
        pc[name] = pc.get(name, 0) + 1
        if kind == "op":
            for res, slices, k in x:
                self.access(res, slices, k, pc, (name, pc[name]))
        else:
            self.bars[x].arrive(pc)

    def issue_load(self, actor, bar, res, slices):
        c = dict(self.tick(actor))
        self.bars[bar].arrive(c, expect_tx=1)                 # mbarrier.arrive.expect_tx by the producer thread
        self.n_async += 1
        self.loads.append((bar, res, slices, c, f"tma:{self.n_async}", actor))

    def step_load(self, i):
        bar, res, slices, c, name, _actor = self.loads.pop(i)
        c = dict(c)
        c[name] = 1
        self.access(res, slices, "write", c, (name, 1))
        self.bars[bar].complete_tx(c)

    def issue_bulk(self, actor, res, slices):
        """A bulk async-group that READS a shared-memory buffer (TMA reduce-add / TMA store)."""
        c = dict(self.tick(actor))
        self.n_async += 1
        self.bulk[actor].append((res, slices, c, f"bulk:{self.n_async}"))

    def step_bulk(self, actor):
        res, slices, c, name = self.bulk[actor].popleft()
        c = dict(c)
        c[name] = 1
        self.access(res, slices, "read", c, (name, 1))
        join(self.bulk_done_clock[actor], c)

    # ---- what a model provides ---------------------------------------------------------------------------------------
    CONST = dict(SA=SA, SXK=SXK, SXM=SXM, SZ=SZ, SDX=SDX)     # the default build (PMF_SA = 3)

    def setup(self):
        c = self.CONST
        self.SA, self.SXK, self.SXM, self.SZ, self.SDX = c["SA"], c["SXK"], c["SXM"], c["SZ"], c["SDX"]
        self.LA, self.RLAG = self.SZ - 1, (5 if self.SDX == 2 else 3)
        # the kernel's static_assert(SXK >= LA): "with fewer XK stages than the MMA1 look-ahead the X producer deadlocks at
        # item boundaries" -- the XM copy of a tile is issued LA tiles after its XK pair
        if self.mutate == "two_xk_stages":
            self.SXK = 2

    def barrier_table(self):                        # mbar_init counts, fused_tc.cu:226-235
        # variant "per_group_full_a" (the proposed fix of the finding in tests/test_tc_protocol.py): one FULL_A barrier per
        # (stage, epilogue group), so that a group sees CONSECUTIVE phases of the barrier it waits on
        n_full_a = 2 * self.SA if self.mutate == "per_group_full_a" else self.SA
        return (("FULL_XK", self.SXK, 1), ("EMPTY_XK", self.SXK, 1), ("FULL_XM", self.SXM, 1), ("EMPTY_XM", self.SXM, 1), ("FULL_A", n_full_a, 1),
                ("EMPTY_AG", self.SA, 1), ("Z_FULL", self.SZ, 1), ("G_READY", self.SZ, NEPI // 2), ("DX_FULL", 2, 1), ("DX_EMPTY", 2, NDRAIN),
                ("Y_READY", 1, NEPI), ("DY_FULL", 1, 1), ("DY_EMPTY", 1, NEPI), ("DXS_FULL", self.SDX, NDRAIN), ("DXS_DONE", self.SDX, 1),
                ("Z_EMPTY", self.SZ, 1))

    def actors(self):
        a = {"TMA_A": self.tma_a(), "TMA_X": self.tma_x(), "MMA1": self.mma1(), "MMA": self.mma()}
        for d in range(NDRAIN):
            a["DRAIN%d" % d] = self.drain(d)
        for w in range(NEPI):
            a["EPI%d" % w] = self.epi(w)
        return a

    # ---- the roles (generators yield ("wait", bar, parity) / ("wait_reduce", n) when they may block) --------------------
    def tiles(self):
        for q, n in enumerate(self.items):
            yield q, n

    def tma_a(self):
        """fused_tc.cu:257-280."""
        r = Ring()
        g = 0
        for _q, n in self.tiles():
            for _ in range(n):
                if self.mutate != "no_empty_ag":
                    yield ("wait", ("EMPTY_AG", r.s), r.ph ^ 1)
                full = r.s + self.SA * (g & 1) if self.mutate == "per_group_full_a" else r.s
                self.issue_load("TMA_A", ("FULL_A", full), "AG%d" % r.s, range(8))
                r.next(self.SA)
                g += 1

    def tma_x(self):
        """fused_tc.cu:281-321."""
        rk, rm = Ring(), Ring()
        pend = [-1] * self.LA

        def load_xm():
            yield ("wait", ("EMPTY_XM", rm.s), rm.ph ^ 1)
            self.issue_load("TMA_X", ("FULL_XM", rm.s), "XM%d" % rm.s, [0])
            rm.next(self.SXM)
        for _q, n in self.tiles():
            for t in range(n):
                if self.mutate != "no_empty_xk":
                    yield ("wait", ("EMPTY_XK", rk.s), rk.ph ^ 1)
                self.issue_load("TMA_X", ("FULL_XK", rk.s), "XK%d" % rk.s, [0])
                rk.next(self.SXK)
                if pend[0] >= 0:
                    yield from load_xm()
                pend = pend[1:] + [t]
        for k in range(self.LA):
            if pend[k] >= 0:
                yield from load_xm()

    def mma1(self):
        """fused_tc.cu:322-403: Z = Y X' up to self.SZ tiles ahead, and the TMA reduce-adds of the staged dX tiles."""
        rx1, rz1 = Ring(), Ring()
        gr = 0

        def reduce_tile():
            nonlocal gr
            sb = gr % self.SDX
            yield ("wait", ("DXS_FULL", sb), (gr // self.SDX) & 1)
            self.issue_bulk("MMA1", "DXS%d" % sb, range(NDRAIN))
            if self.SDX == 2:
                yield ("wait_reduce", 1)                                 # cp.async.bulk.wait_group.read 1
                if gr > 0:
                    self.bars[("DXS_DONE", sb ^ 1)].arrive(self.tick("MMA1"))
            else:
                yield ("wait_reduce", 0)
                self.bars[("DXS_DONE", 0)].arrive(self.tick("MMA1"))
            gr += 1
        for q, n in self.tiles():
            if self.mutate != "no_y_ready":
                yield ("wait", ("Y_READY", 0), q & 1)
            it_r = 0
            for it in range(n):
                if self.mutate != "no_z_empty":
                    yield ("wait", ("Z_EMPTY", rz1.s), rz1.ph ^ 1)
                yield ("wait", ("FULL_XK", rx1.s), rx1.ph)
                self.issue_mma("MMA1", [("YB", range(NEPI), "read"), ("XK%d" % rx1.s, [0], "read"), ("Z%d" % rz1.s, range(8), "write")])
                self.commit("MMA1", ("EMPTY_XK", rx1.s))
                self.commit("MMA1", ("Z_FULL", rz1.s))
                rx1.next(self.SXK)
                rz1.next(self.SZ)
                if it - it_r >= self.RLAG:
                    yield from reduce_tile()
                    it_r += 1
            while it_r < n:
                yield from reduce_tile()
                it_r += 1
        yield ("wait_reduce", 0)
        if self.SDX == 2 and gr > 0:
            self.bars[("DXS_DONE", (gr - 1) & 1)].arrive(self.tick("MMA1"))

    def mma(self):
        """fused_tc.cu:405-466: dX = G' Ys (MMA2), dY += G X (MMA3)."""
        rx3, ra, rz = Ring(), Ring(), Ring()
        g = 0
        for q, n in self.tiles():
            yield ("wait", ("Y_READY", 0), q & 1)
            if self.mutate != "no_dy_empty":
                yield ("wait", ("DY_EMPTY", 0), (q & 1) ^ 1)
            for _ in range(n):
                b, ph = g & 1, (g >> 1) & 1
                if self.mutate != "no_g_ready":
                    yield ("wait", ("G_READY", rz.s), rz.ph)
                if self.mutate != "no_dx_empty":
                    yield ("wait", ("DX_EMPTY", b), ph ^ 1)
                self.issue_mma("MMA", [("AG%d" % ra.s, range(8), "read"), ("YS", range(NEPI), "read"), ("DX%d" % b, range(NDRAIN), "write")])
                self.commit("MMA", ("EMPTY_AG", ra.s))
                self.commit("MMA", ("DX_FULL", b))
                ra.next(self.SA)
                if self.mutate != "no_full_xm":
                    yield ("wait", ("FULL_XM", rx3.s), rx3.ph)
                self.issue_mma("MMA", [("Z%d" % rz.s, range(8), "read"), ("XM%d" % rx3.s, [0], "read"), ("DY", range(NEPI), "write")])
                self.commit("MMA", ("EMPTY_XM", rx3.s))
                self.commit("MMA", ("Z_EMPTY", rz.s))
                rx3.next(self.SXM)
                rz.next(self.SZ)
                g += 1
            self.commit("MMA", ("DY_FULL", 0))

    def drain(self, d):
        """fused_tc.cu:467-529: dX accumulator -> staging buffer."""
        me = "DRAIN%d" % d
        g = 0
        for _q, n in self.tiles():
            for _ in range(n):
                b, sb = g & 1, g % self.SDX
                yield ("wait", ("DX_FULL", b), (g >> 1) & 1)
                self.sync_access(me, "DX%d" % b, [d], "read")
                self.bars[("DX_EMPTY", b)].arrive(self.tick(me))
                if self.mutate != "no_dxs_done":
                    yield ("wait", ("DXS_DONE", sb), ((g // self.SDX) - 1) & 1)
                self.sync_access(me, "DXS%d" % sb, [d], "write")
                self.bars[("DXS_FULL", sb)].arrive(self.tick(me))
                g += 1

    def epi(self, w):
        """fused_tc.cu:531-955: warp w of 16; group = w // 8 takes every other tile."""
        me = "EPI%d" % w
        grp, sl = w // 8, w % 8
        ra, rz = Ring(), Ring()
        if grp == 1:
            ra.next(self.SA)
            rz.next(self.SZ)
        g = 0
        for q, n in self.tiles():
            # item prologue: this warp's part of the Y operands (TMEM [Yh|Yl], shared-memory Ys), :693-717
            self.sync_access(me, "YS", [w], "write")
            self.sync_access(me, "YB", [w], "write")
            self.bars[("Y_READY", 0)].arrive(self.tick(me))
            for _ in range(n):
                mine = (g & 1) == grp
                g += 1
                if not mine:
                    continue
                if self.mutate != "no_z_full":
                    yield ("wait", ("Z_FULL", rz.s), rz.ph)
                if self.mutate == "per_group_full_a":       # this group's k-th tile on this stage: k = (g - 1) // (2 self.SA)
                    yield ("wait", ("FULL_A", ra.s + self.SA * grp), ((g - 1) // (2 * self.SA)) & 1)
                elif self.mutate != "no_full_a":
                    yield ("wait", ("FULL_A", ra.s), ra.ph)
                self.sync_access(me, "Z%d" % rz.s, [sl], "read")
                self.sync_access(me, "AG%d" % ra.s, [sl], "read")
                self.sync_access(me, "Z%d" % rz.s, [sl], "write")        # dloss/dz in place of Z (A operand of MMA3)
                self.sync_access(me, "AG%d" % ra.s, [sl], "write")       # and over the A values (A operand of MMA2)
                self.bars[("G_READY", rz.s)].arrive(self.tick(me))
                ra.next(self.SA); ra.next(self.SA)
                rz.next(self.SZ); rz.next(self.SZ)
            if self.mutate != "no_dy_full":
                yield ("wait", ("DY_FULL", 0), q & 1)
            self.sync_access(me, "DY", [w], "read")
            self.bars[("DY_EMPTY", 0)].arrive(self.tick(me))

    def weight(self, choice):
        key = choice[1] if choice[0] in ("actor", "pipe", "bulk") else "tma-engine"
        key = (choice[0], key)
        if key not in self.weights:
            self.weights[key] = self.rng.choice([0.02, 0.2, 1.0, 1.0, 5.0])
        return self.weights[key]

    # ---- scheduler -----------------------------------------------------------------------------------------------------------
    def run(self):
        actors = self.actors()
        blocked = {}                                   # actor -> the op it is waiting on
        for a in list(actors):
            try:
                blocked[a] = next(actors[a])
            except StopIteration:
                del actors[a]

        def enabled(a):
            op = blocked[a]
            if op[0] == "wait":
                return self.bars[op[1]].parity_passes(op[2])
            return len(self.bulk[a]) <= op[1]          # cp.async.bulk.wait_group.read n

        steps = 0
        while actors or self.loads or any(self.bulk.values()) or any(self.mma_q.values()):
            choices = [("actor", a) for a in actors if enabled(a)]
            choices += [("pipe", t) for t, q in self.mma_q.items() if q]
            if self.ordered_loads:                        # TMA loads of one issuing thread complete in issue order
                seen = set()
                for i, l in enumerate(self.loads):
                    if l[5] not in seen:
                        seen.add(l[5])
                        choices.append(("load", i))
            else:                                         # the PTX model: no order among bulk asynchronous copies
                choices += [("load", i) for i in range(len(self.loads))]
            choices += [("bulk", t) for t, q in self.bulk.items() if q]
            if not choices:
                raise Deadlock("; ".join(f"{a} waits for {blocked[a][1]}" + (f" parity {blocked[a][2]}" if blocked[a][0] == "wait" else "")
                                          for a in sorted(actors)))
            if self.weights is None:
                kind, x = self.rng.choice(choices)
            else:                                         # starve some actors / engines, rush others
                ws = [self.weight(c) for c in choices]
                kind, x = self.rng.choices(choices, weights=ws)[0]
            steps += 1
            if kind == "pipe":
                self.step_pipe(x)
            elif kind == "load":
                self.step_load(x)
            elif kind == "bulk":
                self.step_bulk(x)
            else:
                op = blocked[x]
                if op[0] == "wait":
                    join(self.vc[x], self.bars[op[1]].clock)      # acquire
                else:
                    join(self.vc[x], self.bulk_done_clock[x])
                try:
                    blocked[x] = next(actors[x])
                except StopIteration:
                    del actors[x]
                    del blocked[x]
        return steps


# every one of these removes an ordering the kernel needs (a race, or for two_xk_stages the deadlock its static_assert names)
class Zlink(Sim):
    """zlink_kernel of csrc/wide_tc.cu (K > 64: Z contraction + link epilogue, :150-400).  `items` = 128-sample tiles per
    range of one CTA; every tile runs `nks` 64-factor slabs through the operand ring and hands two 64-sample sub-tiles of
    data in / G' out through the A/G ring."""
    ZS, ZSZ, ZSA, NE = 2, 4, 3, 16                    # wide_tc.cu:130, :135
    nks = 2                                           # K = 128

    def setup(self):
        pass

    def barrier_table(self):                          # mbar_init, wide_tc.cu:168-173: ZEMPTY and G_READY take one arrive per epilogue warp
        return (("FULL", self.ZS, 1), ("EMPTY", self.ZS, 1), ("ZFULL", self.ZSZ, 1), ("ZEMPTY", self.ZSZ, self.NE),
                ("AG_FULL", self.ZSA, 1), ("AG_EMPTY", self.ZSA, 1), ("G_READY", self.ZSA, self.NE))

    def actors(self):
        a = {"TMA": self.tma(), "TMA_A": self.tma_a(), "GST": self.gst(), "MMA": self.mma()}
        for w in range(self.NE):
            a["EPI%d" % w] = self.epi(w)
        return a

    def tma(self):
        """Operand producer, wide_tc.cu:184-204."""
        r = Ring()
        for _q, n in self.tiles():
            for _ in range(n):
                for _ks in range(self.nks):
                    if self.mutate != "no_empty":
                        yield ("wait", ("EMPTY", r.s), r.ph ^ 1)
                    self.issue_load("TMA", ("FULL", r.s), "ST%d" % r.s, [0])
                    r.next(self.ZS)

    def tma_a(self):
        """Data producer, :205-224."""
        r = Ring()
        for _q, n in self.tiles():
            for _ in range(2 * n):
                if self.mutate != "no_ag_empty":
                    yield ("wait", ("AG_EMPTY", r.s), r.ph ^ 1)
                self.issue_load("TMA_A", ("AG_FULL", r.s), "AG%d" % r.s, range(self.NE))
                r.next(self.ZSA)

    def gst(self):
        """G' store issuer, :225-249: a buffer returns to the data producer once the store engine has READ it."""
        r = Ring()
        prev = -1
        for _q, n in self.tiles():
            for _ in range(2 * n):
                if self.mutate != "no_g_ready":
                    yield ("wait", ("G_READY", r.s), r.ph)
                self.issue_bulk("GST", "AG%d" % r.s, range(self.NE))
                if self.mutate != "no_store_wait":
                    yield ("wait_reduce", 1)
                if prev >= 0:
                    self.bars[("AG_EMPTY", prev)].arrive(self.tick("GST"))
                prev = r.s
                r.next(self.ZSA)
        yield ("wait_reduce", 0)
        if prev >= 0:
            self.bars[("AG_EMPTY", prev)].arrive(self.tick("GST"))

    def mma(self):
        """:250-282."""
        r, rz = Ring(), Ring()
        for _q, n in self.tiles():
            for _ in range(n):
                if self.mutate != "no_zempty":
                    yield ("wait", ("ZEMPTY", rz.s), rz.ph ^ 1)
                for _ks in range(self.nks):
                    if self.mutate != "no_full":
                        yield ("wait", ("FULL", r.s), r.ph)
                    self.issue_mma("MMA", [("ST%d" % r.s, [0], "read"), ("ZACC%d" % rz.s, range(self.NE), "write")])
                    self.commit("MMA", ("EMPTY", r.s))
                    r.next(self.ZS)
                self.commit("MMA", ("ZFULL", rz.s))
                rz.next(self.ZSZ)

    def epi(self, w):
        """:283-400: warp = (TMEM lane quarter, 16-sample chunk of the 64-sample sub-tile)."""
        me = "EPI%d" % w
        ra, rz = Ring(), Ring()
        for _q, n in self.tiles():
            for _ in range(n):
                if self.mutate != "no_zfull":
                    yield ("wait", ("ZFULL", rz.s), rz.ph)
                for h in range(2):
                    self.sync_access(me, "ZACC%d" % rz.s, [w], "read")
                    if self.mutate != "no_ag_full":
                        yield ("wait", ("AG_FULL", ra.s), ra.ph)
                    self.sync_access(me, "AG%d" % ra.s, [w], "read")
                    if h == 1:
                        self.bars[("ZEMPTY", rz.s)].arrive(self.tick(me))
                    self.sync_access(me, "AG%d" % ra.s, [w], "write")
                    self.bars[("G_READY", ra.s)].arrive(self.tick(me))
                    ra.next(self.ZSA)
                rz.next(self.ZSZ)


class GradGemm(Sim):
    """grad_gemm_kernel<A_MN> of csrc/wide_tc.cu (:440-545): `items` = contraction steps per output group of one CTA."""
    GS, NE = 3, 4                                     # wide_tc.cu:418-420

    def setup(self):
        pass

    def barrier_table(self):                          # :447
        return (("FULL", self.GS, 1), ("EMPTY", self.GS, 1), ("ACC_FULL", 1, 1), ("ACC_EMPTY", 1, self.NE))

    def actors(self):
        a = {"TMA": self.tma(), "MMA": self.mma()}
        for w in range(self.NE):
            a["EPI%d" % w] = self.epi(w)
        return a

    def tma(self):
        r = Ring()
        for _q, n in self.tiles():
            for _ in range(n):
                if self.mutate != "no_empty":
                    yield ("wait", ("EMPTY", r.s), r.ph ^ 1)
                self.issue_load("TMA", ("FULL", r.s), "ST%d" % r.s, [0])
                r.next(self.GS)

    def mma(self):
        r = Ring()
        for q, n in self.tiles():
            if self.mutate != "no_acc_empty":
                yield ("wait", ("ACC_EMPTY", 0), (q & 1) ^ 1)
            for _ in range(n):
                if self.mutate != "no_full":
                    yield ("wait", ("FULL", r.s), r.ph)
                self.issue_mma("MMA", [("ST%d" % r.s, [0], "read"), ("ACC", range(self.NE), "write")])
                self.commit("MMA", ("EMPTY", r.s))
                r.next(self.GS)
            self.commit("MMA", ("ACC_FULL", 0))

    def epi(self, w):
        me = "EPI%d" % w
        for q, _n in self.tiles():
            if self.mutate != "no_acc_full":
                yield ("wait", ("ACC_FULL", 0), q & 1)
            self.sync_access(me, "ACC", [w], "read")
            self.bars[("ACC_EMPTY", 0)].arrive(self.tick(me))


class FusedSA4(Sim):
    """The -DPMF_SA=4 build of fused_tc.cu (:66-79): four A/G stages, two Xb stages, three Z accumulators, one dX staging
    buffer.  With an even ring every A/G stage belongs to ONE epilogue group, which then sees consecutive phases."""
    CONST = dict(SA=4, SXK=2, SXM=1, SZ=3, SDX=1)


MODELS = {"fused": Sim, "fused_sa4": FusedSA4, "zlink": Zlink, "grad_gemm": GradGemm}
MODEL_MUTATIONS = {"zlink": ["no_empty", "no_ag_empty", "no_g_ready", "no_store_wait", "no_zempty", "no_full", "no_zfull", "no_ag_full"],
                   "grad_gemm": ["no_empty", "no_acc_empty", "no_full", "no_acc_full"]}
MUTATIONS = ["no_z_empty", "no_dx_empty", "no_dxs_done", "no_full_a", "no_dy_full", "no_empty_ag", "no_empty_xk", "no_y_ready",
             "no_g_ready", "no_full_xm", "no_z_full", "two_xk_stages"]
# removing this wait changes nothing: Y_READY of the next item already implies that every epilogue warp has read the dY
# tile of this one (a warp arrives on Y_READY after its item epilogue, in program order)
IMPLIED = ["no_dy_empty"]


def check(items, seeds, mutate=None, first_seed=0, model="fused", ordered_loads=False):
    """(number of clean runs, first failure or None)."""
    for s in range(first_seed, first_seed + seeds):
        try:
            MODELS[model](items, s, mutate, ordered_loads).run()
        except (Race, Deadlock) as e:
            return s - first_seed, f"{type(e).__name__}: {e} (items {items}, seed {s})"
    return seeds, None


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--items", default="5,1,7,2", help="tiles per work item of one CTA")
    ap.add_argument("--seeds", type=int, default=200)
    ap.add_argument("--mutate", default=None)
    ap.add_argument("--model", default="fused", choices=sorted(MODELS), help="fused_tc.cu's data pass, or one of wide_tc.cu's kernels")
    ap.add_argument("--ordered-loads", action="store_true", help="TMA loads of one issuing thread complete in issue order")
    a = ap.parse_args(argv)
    items = [int(x) for x in a.items.split(",")]
    ok, fail = check(items, a.seeds, a.mutate, model=a.model, ordered_loads=a.ordered_loads)
    print(f"{a.model}: items {items}, mutation {a.mutate}: {ok} interleavings clean" + (f"; then {fail}" if fail else ""))
    return 1 if fail else 0


if __name__ == "__main__":
    sys.exit(main())
