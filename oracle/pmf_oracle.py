"""CPU oracle: a NumPy restatement of the PathMatFac.jl fit-loop hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pathmatfac.jl_b200/`` may import this
file; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker.

Parity status
-------------
* Layer / BatchArray / regulariser / FSARD / bookkeeping algebra: PINNED against
  the reference's own known-answer tests (``/root/reference/test/runtests.jl``,
  ported to ``tests/golden/*.json``; see ``tests/test_oracle_golden.py``).
* Noise-model losses and the epoch loop of ``MF.fit!`` live in the un-vendored
  dependency MatFac.jl (dpmerrell/MatFac.jl, v0.1.0, git-tree-sha1
  37d124152593f8e04a24d210d3054a98a2a7bbc9, ``Manifest.toml:558-564``).  Its
  source is not in the container and Julia is not installed, so that part is
  restated from its published algorithm and from the call sites in the reference
  tree: **parity unpinned** for the noise losses, the epoch ordering and the
  termination test (SURVEY.md Appendix D lists every assumed default; each is an
  explicit option here).

Conventions: arrays follow the reference's Julia shapes -- ``X`` is K x M, ``Y``
is K x N, data ``D`` is M x N with NaN for missing.  Ranges are Python
``range(start, stop)`` (0-based, half-open); the golden tests convert from the
Julia 1-based inclusive ranges.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

###############################################################################
# Index / ID bookkeeping          (reference: src/util.jl:140-262)
###############################################################################


def unique_in_order(vec) -> list:
    """Julia ``unique``: first-appearance order."""
    seen = set()
    out = []
    for v in vec:
        if v not in seen:
            seen.add(v)
            out.append(v)
    return out


def is_contiguous(vec) -> bool:
    """src/util.jl:140-156."""
    past = set()
    vec = list(vec)
    for i in range(len(vec) - 1):
        nxt = vec[i + 1]
        if nxt in past:
            return False
        if nxt != vec[i]:
            past.add(vec[i])
    return True


def ids_to_ranges(id_vec) -> List[range]:
    """src/util.jl:187-197 (returns 0-based half-open ranges)."""
    id_vec = list(id_vec)
    assert is_contiguous(id_vec), "IDs in id_vec need to appear in contiguous chunks."
    out = []
    n = len(id_vec)
    for u in unique_in_order(id_vec):
        first = id_vec.index(u)
        last = n - 1 - id_vec[::-1].index(u)
        out.append(range(first, last + 1))
    return out


def ids_to_ind_mat(id_vec) -> np.ndarray:
    """src/util.jl:200-210: Bool indicator, columns in ``unique`` order."""
    id_vec = list(id_vec)
    unq = unique_in_order(id_vec)
    ind = np.zeros((len(id_vec), len(unq)), dtype=bool)
    for c, name in enumerate(unq):
        ind[:, c] = [v == name for v in id_vec]
    return ind


def subset_ranges(ranges: Sequence[range], rng: range):
    """src/util.jl:214-253.  Returns (new_ranges, r_min_idx, r_max_idx) with the
    two indices 0-based *inclusive* (Julia's minus one); empty -> ([], 0, -1)."""
    ranges = list(ranges)
    if len(ranges) == 0 or len(rng) == 0:
        return [], 0, -1
    r_min = max(rng.start, ranges[0].start)
    r_max = min(rng.stop, ranges[-1].stop) - 1  # inclusive
    if r_min > r_max:
        return [], 0, -1
    # last range whose start <= r_min
    lo = max(i for i, r in enumerate(ranges) if r.start <= r_min)
    if r_min > ranges[lo].stop - 1:
        lo += 1
        r_min = ranges[lo].start
    # first range whose (inclusive) stop >= r_max
    hi = min(i for i, r in enumerate(ranges) if r.stop - 1 >= r_max)
    if r_max < ranges[hi].start:
        hi -= 1
        r_max = ranges[hi].stop - 1
    if lo > hi:
        return [], 0, -1
    new = [range(r.start, r.stop) for r in ranges[lo:hi + 1]]
    new[0] = range(r_min, new[0].stop)
    new[-1] = range(new[-1].start, r_max + 1)
    return new, lo, hi


def value_to_idx(values) -> dict:
    return {v: i for i, v in enumerate(values)}


def keymatch(l_keys, r_keys):
    """src/util.jl:169-184 (0-based)."""
    r = value_to_idx(r_keys)
    li, ri = [], []
    for i, k in enumerate(l_keys):
        if k in r:
            li.append(i)
            ri.append(r[k])
    return li, ri


def nansum(x):
    x = np.asarray(x, dtype=float)
    return float(np.sum(x[~np.isnan(x)]))


def nanmean(x):
    x = np.asarray(x, dtype=float)
    return float(np.mean(x[~np.isnan(x)]))


def nanvar(x):
    x = np.asarray(x, dtype=float)
    return float(np.var(x[~np.isnan(x)], ddof=1))


###############################################################################
# Sparse-matrix bookkeeping        (reference: src/util.jl:269-314, 453-511)
###############################################################################


def edgelist_to_spmat(edgelist, node_to_idx: dict, epsilon: float = 0.0) -> sp.csc_matrix:
    """Signed-graph Laplacian, src/util.jl:269-314.  Duplicate (unordered) edges
    keep the latest value; diag = epsilon + sum |w|; off-diag = -w."""
    N = len(node_to_idx)
    edge_dict: Dict[Tuple[int, int], float] = {}
    for e in edgelist:
        e1, e2 = node_to_idx[e[0]], node_to_idx[e[1]]
        edge_dict[(max(e1, e2), min(e1, e2))] = float(e[2])
    diag = np.full(N, float(epsilon))
    I, J, V = [], [], []
    for (i, j), v in edge_dict.items():
        I += [i, j]
        J += [j, i]
        V += [-v, -v]
        diag[i] += abs(v)
        diag[j] += abs(v)
    I = list(range(N)) + I
    J = list(range(N)) + J
    V = list(diag) + V
    # Julia's sparse(I,J,V) sums duplicates (a self-loop would add onto the diagonal)
    return sp.coo_matrix((V, (I, J)), shape=(N, N)).tocsc()


def csc_select(A: sp.spmatrix, rng1: range, rng2: range) -> sp.csc_matrix:
    """src/util.jl:494-511."""
    return sp.csc_matrix(A.tocsc()[rng1.start:rng1.stop, rng2.start:rng2.stop])


def featuresets_to_csc(feature_ids, feature_sets) -> sp.csc_matrix:
    """src/util.jl:453-477: L x N, each row scaled 1/sqrt(set size); Float32 values."""
    f_to_j = value_to_idx(list(feature_ids))
    I, J, V = [], [], []
    for i, fs in enumerate(feature_sets):
        fs = list(fs)
        scale = np.float32(1.0 / math.sqrt(len(fs)))
        for f in fs:
            I.append(i)
            J.append(f_to_j[f])
            V.append(scale)
    return sp.coo_matrix((np.asarray(V, dtype=np.float32), (I, J)),
                         shape=(len(feature_sets), len(f_to_j))).tocsc()


def get_all_nodes(edgelist) -> set:
    s = set()
    for e in edgelist:
        s.add(e[0])
        s.add(e[1])
    return s


def compute_nongraph_nodes(feature_ids, edgelists):
    allf = set(feature_ids)
    return [allf - get_all_nodes(el) for el in edgelists]


###############################################################################
# BatchArray                        (reference: src/batch_array.jl)
###############################################################################


class BatchArray:
    """src/batch_array.jl:5-15.  ``row_batches[v]`` is the M x n_b Bool indicator;
    ``values[v]`` is n_b x N_v."""

    def __init__(self, col_ranges, col_range_ids, row_batches, row_batch_ids, values):
        self.col_ranges = list(col_ranges)
        self.col_range_ids = list(col_range_ids)
        self.row_batches = list(row_batches)
        self.row_batch_ids = list(row_batch_ids)
        self.values = list(values)

    @classmethod
    def construct(cls, feature_views, row_batch_dict, value_dicts):
        """src/batch_array.jl:47-79: views missing from the dict are skipped."""
        unq_views = unique_in_order(feature_views)
        kept = [i for i, v in enumerate(unq_views) if v in row_batch_dict]
        kept_views = [unq_views[i] for i in kept]
        row_batch_ids = [list(row_batch_dict[v]) for v in kept_views]
        unq_rb = [unique_in_order(r) for r in row_batch_ids]
        col_ranges = ids_to_ranges(feature_views)
        kept_ranges = [col_ranges[i] for i in kept]
        kept_vd = [value_dicts[i] for i in kept]
        values = []
        for urb, cr, vd in zip(unq_rb, kept_ranges, kept_vd):
            v = np.zeros((len(urb), len(cr)))
            for i, rb in enumerate(urb):
                v[i, :] = vd[rb]
            values.append(v)
        row_batches = [ids_to_ind_mat(r) for r in row_batch_ids]
        return cls(kept_ranges, kept_views, row_batches, unq_rb, values)

    # -- helpers ---------------------------------------------------------
    def batch_index(self, v) -> np.ndarray:
        """ordinal (0-based) of each sample's batch in view v; -1 if none."""
        rb = self.row_batches[v]
        idx = np.full(rb.shape[0], -1, dtype=np.int32)
        r, c = np.nonzero(rb)
        idx[r] = c
        return idx

    def zero(self):
        return BatchArray(self.col_ranges, self.col_range_ids, self.row_batches,
                          self.row_batch_ids, [np.zeros_like(v) for v in self.values])

    def copy(self):
        return BatchArray(self.col_ranges, self.col_range_ids, self.row_batches,
                          self.row_batch_ids, [v.copy() for v in self.values])

    def exp(self):
        """src/batch_array.jl:230-237."""
        return BatchArray(self.col_ranges, self.col_range_ids, self.row_batches,
                          self.row_batch_ids, [np.exp(v) for v in self.values])

    def view(self, rows, cols: range):
        """src/batch_array.jl:83-106.  ``rows`` is a range or an int array."""
        new_ranges, lo, hi = subset_ranges(self.col_ranges, cols)
        rows = np.arange(rows.start, rows.stop) if isinstance(rows, range) else np.asarray(rows)
        if lo > hi:
            return BatchArray([], [], [], [], [])
        shifted = [range(r.start - cols.start, r.stop - cols.start) for r in new_ranges]
        vals = []
        for r_new, r_old, v in zip(new_ranges, self.col_ranges[lo:hi + 1], self.values[lo:hi + 1]):
            vals.append(v[:, r_new.start - r_old.start: r_new.stop - r_old.start])
        return BatchArray(shifted, self.col_range_ids[lo:hi + 1],
                          [rb[rows, :] for rb in self.row_batches[lo:hi + 1]],
                          self.row_batch_ids[lo:hi + 1], vals)

    # -- arithmetic against a dense matrix ------------------------------------
    def expand(self, v) -> np.ndarray:
        """row_batches[v] * values[v]  (the M x N_v dense buffer)."""
        return self.row_batches[v].astype(self.values[v].dtype) @ self.values[v]

    def add_to(self, A):
        """``A + B``  src/batch_array.jl:121-129."""
        out = A.copy()
        for v, cr in enumerate(self.col_ranges):
            out[:, cr.start:cr.stop] += self.expand(v)
        return out

    def mul_to(self, A):
        """``A * B``  src/batch_array.jl:174-181."""
        out = A.copy()
        for v, cr in enumerate(self.col_ranges):
            out[:, cr.start:cr.stop] *= self.expand(v)
        return out

    def add_pullback(self, result_bar):
        """src/batch_array.jl:132-150 -> (A_bar, values_bar)."""
        vb = []
        for v, cr in enumerate(self.col_ranges):
            rb = self.row_batches[v].astype(result_bar.dtype)
            vb.append(rb.T @ result_bar[:, cr.start:cr.stop])
        return result_bar.copy(), vb

    def mul_pullback(self, A, result_bar):
        """src/batch_array.jl:184-212 -> (A_bar, values_bar); A is the forward input."""
        A_bar = result_bar.copy()
        vb = []
        for v, cr in enumerate(self.col_ranges):
            A_bar[:, cr.start:cr.stop] *= self.expand(v)
            rb = self.row_batches[v].astype(result_bar.dtype)
            vb.append(rb.T @ (A[:, cr.start:cr.stop] * result_bar[:, cr.start:cr.stop]))
        return A_bar, vb


def ba_map(map_func: Callable, template: BatchArray, *args, capacity=10 ** 8):
    """src/batch_array.jl:320-334: segmented (per batch) column sums of map_func(args)."""
    M, N = args[0].shape
    row_batch_size = max(1, capacity // N)
    result = template.zero()
    for r0 in range(0, M, row_batch_size):
        rows = range(r0, min(M, r0 + row_batch_size))
        q = map_func(*[a[rows.start:rows.stop, :] for a in args])
        for v, (rb, cr) in enumerate(zip(result.row_batches, result.col_ranges)):
            rbv = rb[rows.start:rows.stop, :].astype(q.dtype)
            result.values[v] += rbv.T @ q[:, cr.start:cr.stop]
    return result.values


###############################################################################
# Noise models  (MatFac.jl, EXTERNAL -- restated)
###############################################################################

DIST_CODES = {"normal": 0, "bernoulli": 1, "poisson": 2, "ordinal3": 3,
              "bernoulli_sq_hinge": 4, "ordinal_sq_hinge3": 5}
VALID_LOSSES = list(DIST_CODES)  # src/util.jl:128
ORDINAL_EPS = 1e-10      # guard inside the ordinal log (assumed; option)
SQ_HINGE_MARGIN = 1.0    # margin of the ordinal squared hinge (assumed; option)


def _sigmoid(z):
    return 0.5 * (1.0 + np.tanh(0.5 * z))


def _softplus(z):
    return np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))


def noise_loss_grad(dist: str, z, a, thresholds=None):
    """Per-entry (loss, dloss/dz) at finite entries of ``a``; NaN entries give 0.
    ``thresholds`` = ext_thresholds [-inf, t1, .., +inf] for ordinal types."""
    mask = np.isfinite(a)
    a0 = np.where(mask, a, 0.0).astype(z.dtype)
    if dist == "normal":
        g = z - a0
        l = 0.5 * g * g
    elif dist == "bernoulli":
        l = _softplus(z) - a0 * z
        g = _sigmoid(z) - a0
    elif dist == "poisson":
        ez = np.exp(z)
        l = ez - a0 * z
        g = ez - a0
    elif dist == "ordinal3":
        t = np.asarray(thresholds, dtype=z.dtype)
        cat = np.where(mask, a0, 1).astype(np.int64)  # 1-based category
        lo, hi = t[cat - 1], t[cat]
        sr = _sigmoid(hi - z)   # sigmoid(+inf)=1
        sl = _sigmoid(lo - z)   # sigmoid(-inf)=0
        l = -np.log(sr - sl + z.dtype.type(ORDINAL_EPS))
        g = 1.0 - sr - sl
    elif dist == "bernoulli_sq_hinge":
        y = 2.0 * a0 - 1.0
        h = np.maximum(0.0, 1.0 - y * z)
        l = h * h
        g = -2.0 * y * h
    elif dist == "ordinal_sq_hinge3":
        t = np.asarray(thresholds, dtype=z.dtype)
        cat = np.where(mask, a0, 1).astype(np.int64)
        lo, hi = t[cat - 1], t[cat]
        m = z.dtype.type(SQ_HINGE_MARGIN)
        with np.errstate(invalid="ignore"):
            hl = np.where(np.isfinite(lo), np.maximum(0.0, lo - z + m), 0.0)
            hr = np.where(np.isfinite(hi), np.maximum(0.0, z - hi + m), 0.0)
        l = hl * hl + hr * hr
        g = -2.0 * hl + 2.0 * hr
    else:
        raise ValueError(dist)
    return np.where(mask, l, 0.0), np.where(mask, g, 0.0)


def noise_threshold_grads(dist: str, z, a, thresholds):
    """Per-entry derivatives of the ordinal losses with respect to the two INTERIOR thresholds t1, t2
    (``ext_thresholds`` = [-inf, t1, t2, +inf]; the outer two are fixed, src/fit.jl:228-242).  With lo = t[c-1],
    hi = t[c] for category c:  ordinal3  l = -log(s(hi-z) - s(lo-z) + eps):  dl/dhi = -s'(hi-z)/p, dl/dlo = s'(lo-z)/p;
    ordinal_sq_hinge3  l = max(0, lo-z+m)^2 + max(0, z-hi+m)^2:  dl/dlo = 2 hl, dl/dhi = -2 hr.
    Returns (d/dt1, d/dt2) arrays shaped like z; zero at missing entries.  SURVEY App. D7: which noise parameters
    MatFac.jl trains is INFERRED -- thresholds are the only ones the reference ever touches."""
    mask = np.isfinite(a)
    a0 = np.where(mask, a, 1.0)
    cat = a0.astype(np.int64)
    t = np.asarray(thresholds, dtype=z.dtype)
    lo, hi = t[cat - 1], t[cat]
    if dist == "ordinal3":
        sr, sl = _sigmoid(hi - z), _sigmoid(lo - z)
        p = sr - sl + z.dtype.type(ORDINAL_EPS)
        dhi = -sr * (1.0 - sr) / p
        dlo = sl * (1.0 - sl) / p
    elif dist == "ordinal_sq_hinge3":
        m = z.dtype.type(SQ_HINGE_MARGIN)
        with np.errstate(invalid="ignore"):
            hl = np.where(np.isfinite(lo), np.maximum(0.0, lo - z + m), 0.0)
            hr = np.where(np.isfinite(hi), np.maximum(0.0, z - hi + m), 0.0)
        dlo, dhi = 2.0 * hl, -2.0 * hr
    else:
        raise ValueError(dist)
    # category c touches thresholds c-1 (as lo) and c (as hi); interior ones are indices 1 and 2
    g1 = np.where(cat == 2, dlo, 0.0) + np.where(cat == 1, dhi, 0.0)
    g2 = np.where(cat == 3, dlo, 0.0) + np.where(cat == 2, dhi, 0.0)
    return np.where(mask, g1, 0.0), np.where(mask, g2, 0.0)


@dataclass
class NoiseModel:
    """CompositeNoise stand-in: contiguous column ranges, one distribution each,
    per-column weights (MF.set_weight!, src/fit.jl:157,180)."""
    col_ranges: List[range]
    dists: List[str]
    weights: np.ndarray                       # length N
    thresholds: List[Optional[np.ndarray]]    # per range (ordinal types) or None

    @classmethod
    def from_distributions(cls, feature_distributions):
        ranges = ids_to_ranges(feature_distributions)
        dists = unique_in_order(feature_distributions)
        th = [np.array([-np.inf, -1.0, 1.0, np.inf]) if d.startswith("ordinal") else None
              for d in dists]
        return cls(ranges, dists, np.ones(len(feature_distributions)), th)


###############################################################################
# Regularisers                     (reference: src/regularizers.jl)
###############################################################################


class ZeroReg:
    """the ``x->0`` closures (src/regularizers.jl:674,715; src/fit.jl:415,769)."""

    def value(self, X):
        return 0.0

    def grad(self, X):
        return np.zeros_like(X)


class L2Regularizer:
    """src/regularizers.jl:11-55."""

    def __init__(self, K, w):
        self.weights = np.full(K, float(w))

    def value(self, X):
        return 0.5 * float(np.sum(self.weights[:, None] * X * X))

    def grad(self, X):
        return self.weights[:, None] * X


class GroupRegularizer:
    """src/regularizers.jl:345-456 (value :423-428, rrule :431-446)."""

    def __init__(self, group_labels, weight=1.0, K=1, group_idx=None, group_weights=None):
        self.group_labels = unique_in_order(group_labels)
        self.group_idx = ids_to_ranges(group_labels) if group_idx is None else list(group_idx)
        self.group_weights = ([np.full(K, float(weight)) for _ in self.group_idx]
                              if group_weights is None else [np.asarray(w, float) for w in group_weights])

    def value(self, X):
        return 0.5 * float(sum(np.sum(w[:, None] * X[:, r.start:r.stop] ** 2)
                               for w, r in zip(self.group_weights, self.group_idx)))

    def grad(self, X):
        d = np.zeros_like(X)
        for w, r in zip(self.group_weights, self.group_idx):
            d[:, r.start:r.stop] = w[:, None] * X[:, r.start:r.stop]
        return d


class SelectiveL1Reg:
    """src/regularizers.jl:106-163."""

    def __init__(self, feature_ids, edgelists, weight=1.0):
        l1_features = compute_nongraph_nodes(feature_ids, edgelists)
        self.l1_idx = np.array([[f in s for f in feature_ids] for s in l1_features], dtype=bool)
        self.weight = np.full(len(l1_features), float(weight))

    def value(self, X):
        return float(np.sum(self.weight[:, None] * np.abs(self.l1_idx * X)))

    def grad(self, X):
        return self.weight[:, None] * np.sign(self.l1_idx * X)


def krylov_cg(A, b, x0, atol=None, rtol=None, itmax=0):
    """Krylov.jl 0.9 ``cg(A, b, x0)`` restated: warm start solves A dx = b - A x0,
    tolerance atol + rtol*||r0||, itmax = 2n; atol = rtol = sqrt(eps(T))."""
    n = b.shape[0]
    if n == 0:
        return x0.copy()
    T = b.dtype.type
    eps = np.sqrt(np.finfo(b.dtype).eps)
    atol = eps if atol is None else atol
    rtol = eps if rtol is None else rtol
    itmax = 2 * n if itmax == 0 else itmax
    x = x0.astype(b.dtype).copy()
    r = b - A @ x
    p = r.copy()
    gamma = float(r @ r)
    rnorm = math.sqrt(gamma)
    tol = atol + rtol * rnorm
    it = 0
    while rnorm > tol and it < itmax:
        Ap = A @ p
        pAp = float(p @ Ap)
        if pAp <= 0:
            break
        alpha = gamma / pAp
        x += T(alpha) * p
        r -= T(alpha) * Ap
        gamma_next = float(r @ r)
        rnorm = math.sqrt(gamma_next)
        beta = gamma_next / gamma
        gamma = gamma_next
        p = r + T(beta) * p
        it += 1
    return x


class NetworkRegularizer:
    """src/regularizers.jl:169-338.  ``cg_tol``: the reference uses Krylov's
    defaults (sqrt(eps) ~ 1.5e-8 in Float64); the oracle keeps that."""

    def __init__(self, feature_ids, edgelists, epsilon=0.1, weight=1.0):
        feature_ids = list(feature_ids)
        N = len(feature_ids)
        self.AA, self.AB, self.BB, self.x_virtual = [], [], [], []
        for el in edgelists:
            net_nodes = get_all_nodes(el)
            virt = sorted(net_nodes - set(feature_ids))
            all_nodes = feature_ids + virt
            spmat = edgelist_to_spmat(el, value_to_idx(all_nodes), epsilon=epsilon) * float(weight)
            Nt = spmat.shape[1]
            self.AA.append(csc_select(spmat, range(0, N), range(0, N)))
            self.AB.append(csc_select(spmat, range(0, N), range(N, Nt)))
            self.BB.append(csc_select(spmat, range(N, Nt), range(N, Nt)))
            self.x_virtual.append(np.zeros(Nt - N))
        self.cur_weights = np.full(len(edgelists), float(weight))

    def _solve(self, k, xk):
        xAB = np.asarray(self.AB[k].T @ xk).ravel()
        # quirk (vi): the stored (sign-flipped) vector is fed back as the warm start
        self.x_virtual[k] = -krylov_cg(self.BB[k], xAB, self.x_virtual[k])
        return xAB

    def value(self, X):
        """src/regularizers.jl:249-266."""
        loss = 0.0
        for k in range(X.shape[0]):
            xk = X[k, :]
            xAB = self._solve(k, xk)
            u = self.x_virtual[k]
            nl = float(xk @ (self.AA[k] @ xk)) + 2.0 * float(xAB @ u) + float(u @ (self.BB[k] @ u))
            loss += 0.5 * nl
        return loss

    def value_grad(self, X):
        """src/regularizers.jl:269-306."""
        loss = 0.0
        G = np.zeros_like(X)
        for k in range(X.shape[0]):
            xk = X[k, :]
            xAA = np.asarray(self.AA[k].T @ xk).ravel()
            self._solve(k, xk)
            u = self.x_virtual[k]
            ABu = np.asarray(self.AB[k] @ u).ravel()
            BBu = np.asarray(self.BB[k] @ u).ravel()
            loss += 0.5 * float(xAA @ xk) + float(xk @ ABu) + 0.5 * float(u @ BBu)
            G[k, :] = xAA + ABu
        return loss, G

    def grad(self, X):
        return self.value_grad(X)[1]


class ARDRegularizer:
    """src/regularizers.jl:526-609."""

    def __init__(self, column_groups, alpha=np.float32(1.001), beta=np.float32(0.001), weight=1.0):
        self.col_ranges = ids_to_ranges(column_groups)
        self.alpha = [float(alpha)] * len(self.col_ranges)
        self.beta = [float(beta)] * len(self.col_ranges)

    def value(self, X):
        tot = 0.0
        for a, b, r in zip(self.alpha, self.beta, self.col_ranges):
            Xv = X[:, r.start:r.stop]
            tot += (0.5 + a) * float(np.sum(np.log(1.0 + (0.5 / b) * Xv * Xv)))
        return tot

    def grad(self, X):
        G = np.zeros_like(X)
        for a, b, r in zip(self.alpha, self.beta, self.col_ranges):
            Xv = X[:, r.start:r.stop]
            G[:, r.start:r.stop] = (1.0 / b) * (0.5 + a) * Xv / (1.0 + (0.5 / b) * Xv * Xv)
        return G


class CompositeRegularizer:
    """src/regularizers.jl:616-649: sum_s p_s * reg_s(x)."""

    def __init__(self, regs, mixture_p):
        self.regularizers = list(regs)
        self.mixture_p = [float(p) for p in mixture_p]

    def value(self, X):
        return float(sum(p * r.value(X) for r, p in zip(self.regularizers, self.mixture_p)))

    def grad(self, X):
        G = np.zeros_like(X)
        for r, p in zip(self.regularizers, self.mixture_p):
            if p != 0.0:
                G = G + p * r.grad(X)
        return G


def construct_X_reg(K, M, sample_ids, sample_conditions, sample_graphs, lambda_X_l2,
                    lambda_X_condition, lambda_X_graph, Y_ard, Y_geneset_ard):
    """src/regularizers.jl:655-689."""
    if Y_ard or Y_geneset_ard:
        if sample_conditions is not None:
            return GroupRegularizer(sample_conditions, weight=1.0, K=K)
        return L2Regularizer(K, 1.0)
    regs = [ZeroReg(), ZeroReg(), ZeroReg()]
    p = np.zeros(3)
    if lambda_X_l2 is not None:
        regs[0] = L2Regularizer(K, lambda_X_l2)
        p[0] = 1
    if sample_conditions is not None:
        regs[1] = GroupRegularizer(sample_conditions, weight=lambda_X_condition, K=K)
        p[1] = 1
    if sample_graphs is not None:
        regs[2] = NetworkRegularizer(sample_ids, sample_graphs, weight=lambda_X_graph)
        p[2] = 1
    # the reference divides unguarded here (0/0 = NaN mixture when nothing is enabled, which its
    # fit! stages never evaluate because they install their own X_reg); guard like construct_Y_reg
    s = p.sum()
    return CompositeRegularizer(regs, p / (s if s > 0 else 1))


def construct_Y_reg(K, N, feature_ids, feature_views, feature_sets_dict, feature_graphs,
                    lambda_Y_l2, lambda_Y_selective_l1, lambda_Y_graph, Y_ard, Y_geneset_ard,
                    featureset_names, alpha0, v0):
    """src/regularizers.jl:696-739."""
    if Y_geneset_ard:
        return construct_featureset_ard(K, feature_ids, feature_views, feature_sets_dict,
                                        featureset_ids=featureset_names, alpha0=alpha0, v0=v0)
    if Y_ard:
        return ARDRegularizer(feature_views)
    regs = [ZeroReg(), ZeroReg(), ZeroReg()]
    p = np.zeros(3)
    if lambda_Y_l2 is not None:
        regs[0] = GroupRegularizer(feature_views, K=K, weight=lambda_Y_l2)
        p[0] = 1
    if feature_ids is not None and feature_graphs is not None:
        if lambda_Y_selective_l1 is not None:
            regs[1] = SelectiveL1Reg(feature_ids, feature_graphs, weight=lambda_Y_selective_l1)
            p[1] = 1
        if lambda_Y_graph is not None:
            regs[2] = NetworkRegularizer(feature_ids, feature_graphs, weight=lambda_Y_graph)
            p[2] = 1
    s = p.sum()
    if s == 0:
        s = 1
    return CompositeRegularizer(regs, p / s)


class ColParamReg:
    """src/regularizers.jl:462-519."""

    def __init__(self, feature_views, weight=1.0, center=0.0):
        self.col_ranges = ids_to_ranges(feature_views)
        self.weights = [float(weight)] * len(self.col_ranges)
        self.centers = [float(center)] * len(self.col_ranges)

    def value(self, v):
        return 0.5 * float(sum(w * np.sum((v[r.start:r.stop] - c) ** 2)
                               for c, w, r in zip(self.centers, self.weights, self.col_ranges)))

    def grad(self, v):
        g = np.zeros_like(v)
        for c, w, r in zip(self.centers, self.weights, self.col_ranges):
            g[r.start:r.stop] = w * (v[r.start:r.stop] - c)
        return g


class BatchArrayReg:
    """src/regularizers.jl:781-889: per-batch centres / weights."""

    def __init__(self, ba: BatchArray, center=0.0, weight=1.0):
        self.centers = [np.full(v.shape[0], float(center)) for v in ba.values]
        self.weights = [np.full(v.shape[0], float(weight)) for v in ba.values]

    def value(self, ba: BatchArray):
        return 0.5 * float(sum(np.sum(w[:, None] * (v - c[:, None]) ** 2)
                               for v, c, w in zip(ba.values, self.centers, self.weights)))

    def grad(self, ba: BatchArray):
        return [w[:, None] * (v - c[:, None]) for v, c, w in zip(ba.values, self.centers, self.weights)]


###############################################################################
# Feature-set ARD                  (reference: src/featureset_ard.jl, src/optimizers.jl)
###############################################################################


class ISTAOptimiser:
    """src/optimizers.jl:26-62."""

    def __init__(self, target, lr, l1_lambda):
        self.lr = np.float32(lr)
        self.ssq_grad = np.zeros_like(target) + target.dtype.type(1e-8)
        self.lam = np.asarray(l1_lambda, dtype=target.dtype)

    def update(self, p, g):
        self.ssq_grad += g * g
        eta = self.lr / np.sqrt(self.ssq_grad)
        p -= eta * g
        np.maximum(p, 0, out=p)
        p[...] = np.maximum(np.abs(p) - self.lam[None, :] * eta, 0)


class FeatureSetARDReg:
    """src/featureset_ard.jl:19-65.  A[v]: L_v x K, S[v]: L_v x N_v sparse."""

    def __init__(self, K, feature_views, S_vec, featureset_ids_vec, alpha0=1.01, v0=0.8, lr=0.05,
                 dtype=np.float32):
        N = len(feature_views)
        self.col_ranges = ids_to_ranges(feature_views)
        self.S = [sp.csr_matrix(S, dtype=dtype) for S in S_vec]
        self.A = [np.zeros((S.shape[0], K), dtype=dtype) for S in self.S]
        self.alpha0 = np.float32(alpha0)
        self.v0 = np.float32(v0)
        self.featureset_ids = [list(f) for f in featureset_ids_vec]
        self.alpha = np.full(N, self.alpha0, dtype=dtype)
        self.beta = np.full((K, N), self.alpha0 - np.float32(1), dtype=dtype)
        self.A_opts = [ISTAOptimiser(A, lr, np.ones(K, dtype=dtype)) for A in self.A]
        self.dtype = dtype

    def value(self, Y):
        """src/featureset_ard.jl:135-138."""
        b = 1.0 + (0.5 / self.beta) * (Y * Y)
        return float(np.sum((0.5 + self.alpha)[None, :] * np.sum(np.log(b), axis=0)))

    def grad(self, Y):
        """src/featureset_ard.jl:141-150 (loss_bar = 1)."""
        b = 1.0 + (0.5 / self.beta) * (Y * Y)
        return (self.alpha + 0.5)[None, :] * Y / (b * self.beta)


def construct_featureset_ard(K, feature_ids, feature_views, feature_sets_dict, featureset_ids=None,
                             alpha0=np.float32(1.001), v0=np.float32(0.8), lr=np.float32(0.05),
                             dtype=np.float32):
    """src/featureset_ard.jl:111-132.  ``feature_sets_dict`` may be a dict keyed by
    view or (as in the reference's own test, runtests.jl:823) a list indexed by view."""
    col_ranges = ids_to_ranges(feature_views)
    unq_views = unique_in_order(feature_views)
    feature_ids = list(feature_ids)

    def _get(container, uv, pos):
        if isinstance(container, dict):
            return container[uv]
        return container[uv - 1] if isinstance(uv, (int, np.integer)) else container[pos]

    S_vec, fs_vec = [], []
    for pos, (cr, uv) in enumerate(zip(col_ranges, unq_views)):
        fsets = _get(feature_sets_dict, uv, pos)
        S_vec.append(featuresets_to_csc(feature_ids[cr.start:cr.stop], fsets))
        if featureset_ids is None:
            fs_vec.append(list(range(1, len(fsets) + 1)))
        else:
            fs_vec.append(_get(featureset_ids, uv, pos))
    return FeatureSetARDReg(K, feature_views, S_vec, fs_vec, alpha0=alpha0, v0=v0, lr=lr, dtype=dtype)


def gamma_normal_loss(A, S, alpha, alpha0, v0, Y):
    """src/featureset_ard.jl:154-162."""
    beta0 = alpha0 - 1
    beta = beta0 * (v0 + np.asarray((S.T @ A).T))
    ap5 = alpha + A.dtype.type(0.5)
    lss = -np.sum(alpha[None, :] * np.log(beta)) + np.sum(ap5[None, :] * np.log(beta + A.dtype.type(0.5) * Y * Y))
    lss -= np.sum((ap5 * np.log(ap5) - alpha * np.log(alpha))[None, :]
                  + np.sum(np.log(np.abs(Y) + A.dtype.type(1e-9)), axis=0, keepdims=True))
    return float(lss)


def gamma_normal_grad_A(A, S, alpha, alpha0, v0, Y):
    """src/featureset_ard.jl:164-186 (the pullback)."""
    beta0 = alpha0 - 1
    beta = beta0 * (v0 + np.asarray((S.T @ A).T))
    ap5 = alpha + A.dtype.type(0.5)
    grad_AtS = beta0 * ((-alpha[None, :] / beta) + ap5[None, :] / (beta + A.dtype.type(0.5) * Y * Y))
    return np.asarray(S @ grad_AtS.T)


def update_lambda(reg: FeatureSetARDReg, Y):
    """src/featureset_ard.jl:189-209."""
    for cr, S, A, opt in zip(reg.col_ranges, reg.S, reg.A, reg.A_opts):
        Yv = Y[:, cr.start:cr.stop]
        Y_ms = np.mean(Yv * Yv, axis=1)
        min_ms = min(Y_ms.min(), reg.v0)
        den = Y_ms - min_ms + np.float32(1e-3)
        S_mean = S.sum() / (S.shape[0] * S.shape[1])
        opt.lam[...] = (A.shape[0] * S_mean) / den


def update_A_inner(A, S, Y, alpha, alpha0, v0, A_opt, max_epochs=1000, term_iter=20, atol=1e-5):
    """src/featureset_ard.jl:214-276.  Returns (best_loss, epochs_run)."""
    A_loss = lambda A_: gamma_normal_loss(A_, S, alpha, alpha0, v0, Y)
    reg_loss = lambda A_: float(np.sum(A_opt.lam[None, :] * np.abs(A_)))
    term_count = 0
    best_loss = A_loss(A) + reg_loss(A)
    A_best = A.copy()
    epochs = 0
    for epoch in range(1, max_epochs + 1):
        epochs = epoch
        g = gamma_normal_grad_A(A, S, alpha, alpha0, v0, Y).astype(A.dtype)
        A_opt.update(A, g)
        new_loss = A_loss(A) + reg_loss(A)
        if new_loss < best_loss:
            loss_diff = best_loss - new_loss
            best_loss = new_loss
            A_best[...] = A
            term_count = 0 if loss_diff > atol else term_count + 1
        else:
            term_count += 1
        if term_count >= term_iter:
            break
    A[...] = A_best
    return best_loss, epochs


def update_A(reg: FeatureSetARDReg, Y, max_epochs=1000, term_iter=20, atol=1e-5):
    """src/featureset_ard.jl:278-294."""
    beta0 = reg.alpha0 - np.float32(1)
    out = []
    for cr, A, S, opt in zip(reg.col_ranges, reg.A, reg.S, reg.A_opts):
        Yv = Y[:, cr.start:cr.stop].astype(A.dtype)
        A[...] = 0
        out.append(update_A_inner(A, S, Yv, reg.alpha[cr.start:cr.stop], reg.alpha0, reg.v0, opt,
                                  max_epochs=max_epochs, term_iter=term_iter, atol=atol))
        reg.beta[:, cr.start:cr.stop] = beta0 * (reg.v0 + np.asarray((S.T @ A).T))
    return out


###############################################################################
# The model container and the data pass
###############################################################################


@dataclass
class OracleModel:
    """The state MF.fit! sees (MatFacModel fields used by the reference,
    SURVEY.md Appendix A) plus the PathMatFac layers."""
    X: np.ndarray                      # K x M
    Y: np.ndarray                      # K x N
    logsigma: np.ndarray               # N   (ColScale, layers.jl:9-48)
    mu: np.ndarray                     # N   (ColShift, layers.jl:53-90)
    logdelta: Optional[BatchArray]     # BatchScale (layers.jl:95-152) or None (x->x)
    theta: Optional[BatchArray]        # BatchShift (layers.jl:158-214) or None
    noise: NoiseModel
    X_reg: object = field(default_factory=ZeroReg)
    Y_reg: object = field(default_factory=ZeroReg)
    layer_regs: list = field(default_factory=lambda: [ZeroReg(), ZeroReg(), ZeroReg(), ZeroReg()])
    frozen: list = field(default_factory=lambda: [False, False, False, False])  # FrozenLayer per slot


def forward(m: OracleModel, rows: Optional[range] = None) -> np.ndarray:
    """col_transform(X'Y) in the fixed order ColScale, BatchScale, ColShift,
    BatchShift (layers.jl:221-253)."""
    M = m.X.shape[1]
    rows = range(0, M) if rows is None else rows
    N = m.Y.shape[1]
    Z = m.X[:, rows.start:rows.stop].T @ m.Y
    Z = Z * np.exp(m.logsigma)[None, :]
    if m.logdelta is not None:
        Z = m.logdelta.view(rows, range(0, N)).exp().mul_to(Z)
    Z = Z + m.mu[None, :]
    if m.theta is not None:
        Z = m.theta.view(rows, range(0, N)).add_to(Z)
    return Z


def data_loss_grads(m: OracleModel, D: np.ndarray, capacity: int = 10 ** 8, want_grads=True):
    """Full-batch data loss and every gradient, accumulated over row minibatches of
    ``capacity // N`` rows exactly like the reference (Z materialised per block).
    Follows the rrules layer by layer, including the ColScale quirk
    (layers.jl:39-44: logsigma_bar = sum_i sigma_j * Gbar_ij, no Z factor)."""
    K, M = m.X.shape
    N = m.Y.shape[1]
    dt = m.X.dtype
    out = {"loss": 0.0}
    if want_grads:
        out.update(dX=np.zeros_like(m.X), dY=np.zeros_like(m.Y), dlogsigma=np.zeros(N, dt),
                   dmu=np.zeros(N, dt),
                   dlogdelta=[np.zeros_like(v) for v in m.logdelta.values] if m.logdelta is not None else None,
                   dtheta=[np.zeros_like(v) for v in m.theta.values] if m.theta is not None else None,
                   dthresholds=np.zeros((len(m.noise.dists), 2), np.float64))
    sigma = np.exp(m.logsigma)
    block = max(1, capacity // N)
    for r0 in range(0, M, block):
        rows = range(r0, min(M, r0 + block))
        Xv = m.X[:, rows.start:rows.stop]
        z0 = Xv.T @ m.Y
        z1 = z0 * sigma[None, :]
        if m.logdelta is not None:
            ld_v = m.logdelta.view(rows, range(0, N))
            ed_v = ld_v.exp()
            z2 = ed_v.mul_to(z1)
        else:
            z2 = z1
        z3 = z2 + m.mu[None, :]
        if m.theta is not None:
            th_v = m.theta.view(rows, range(0, N))
            z4 = th_v.add_to(z3)
        else:
            z4 = z3
        Dv = D[rows.start:rows.stop, :]
        G = np.zeros_like(z4)
        for cr, dist, th in zip(m.noise.col_ranges, m.noise.dists, m.noise.thresholds):
            sl = slice(cr.start, cr.stop)
            l, g = noise_loss_grad(dist, z4[:, sl], Dv[:, sl], th)
            w = m.noise.weights[sl].astype(dt)[None, :]
            out["loss"] += float(np.sum(w * l, dtype=np.float64))
            G[:, sl] = w * g
            if want_grads and dist.startswith("ordinal"):
                g1, g2 = noise_threshold_grads(dist, z4[:, sl], Dv[:, sl], th)
                r = m.noise.dists.index(dist)
                out["dthresholds"][r, 0] += float(np.sum(w * g1, dtype=np.float64))
                out["dthresholds"][r, 1] += float(np.sum(w * g2, dtype=np.float64))
        if not want_grads:
            continue
        # BatchShift pullback (batch_array.jl:132-150)
        if m.theta is not None:
            G3, vb = th_v.add_pullback(G)
            for acc, b in zip(out["dtheta"], vb):
                acc += b
        else:
            G3 = G
        # ColShift pullback (layers.jl:78-90)
        out["dmu"] += G3.sum(axis=0)
        G2 = G3
        # BatchScale pullback: '*' (batch_array.jl:184-212) then exp (:246-256)
        if m.logdelta is not None:
            G1, vb = ed_v.mul_pullback(z1, G2)
            for acc, b, ev in zip(out["dlogdelta"], vb, ed_v.values):
                acc += b * ev
        else:
            G1 = G2
        # ColScale pullback (layers.jl:34-48) -- quirk kept
        G0 = sigma[None, :] * G1
        out["dlogsigma"] += G0.sum(axis=0)
        # GEMM pullbacks
        out["dX"][:, rows.start:rows.stop] = m.Y @ G0.T
        out["dY"] += Xv @ G0
    return out


def layer_reg_value(m: OracleModel) -> float:
    """SequenceReg (regularizers.jl:896-938); frozen slot => 0 (FrozenRegularizer :950-1004)."""
    params = [m.logsigma, m.logdelta, m.mu, m.theta]
    tot = 0.0
    for slot, (reg, p) in enumerate(zip(m.layer_regs, params)):
        if p is None or isinstance(reg, ZeroReg) or m.frozen[slot]:
            continue
        tot += reg.value(p)
    return tot


def total_loss_grads(m: OracleModel, D, capacity=10 ** 8):
    """loss = data + X_reg(X) + Y_reg(Y) + layer_reg; gradients add the penalties'
    pullbacks (SURVEY Appendix B tail, D9).  Frozen layers get no gradient."""
    out = data_loss_grads(m, D, capacity)
    comp = {"data": out["loss"]}
    if isinstance(m.X_reg, NetworkRegularizer):
        comp["X_reg"], gx = m.X_reg.value_grad(m.X)
    else:
        comp["X_reg"], gx = m.X_reg.value(m.X), m.X_reg.grad(m.X)
    comp["Y_reg"], gy = _value_grad(m.Y_reg, m.Y)
    out["dX"] = out["dX"] + gx
    out["dY"] = out["dY"] + gy
    comp["layer_reg"] = layer_reg_value(m)
    if not isinstance(m.layer_regs[0], ZeroReg) and not m.frozen[0]:
        out["dlogsigma"] = out["dlogsigma"] + m.layer_regs[0].grad(m.logsigma)
    if not isinstance(m.layer_regs[2], ZeroReg) and not m.frozen[2]:
        out["dmu"] = out["dmu"] + m.layer_regs[2].grad(m.mu)
    if m.logdelta is not None and not isinstance(m.layer_regs[1], ZeroReg) and not m.frozen[1]:
        out["dlogdelta"] = [a + b for a, b in zip(out["dlogdelta"], m.layer_regs[1].grad(m.logdelta))]
    if m.theta is not None and not isinstance(m.layer_regs[3], ZeroReg) and not m.frozen[3]:
        out["dtheta"] = [a + b for a, b in zip(out["dtheta"], m.layer_regs[3].grad(m.theta))]
    out["components"] = comp
    out["loss"] = comp["data"] + comp["X_reg"] + comp["Y_reg"] + comp["layer_reg"]
    return out


def _value_grad(reg, P):
    if isinstance(reg, CompositeRegularizer):
        val, G = 0.0, np.zeros_like(P)
        for r, p in zip(reg.regularizers, reg.mixture_p):
            if p == 0.0 or isinstance(r, ZeroReg):
                continue
            v, g = _value_grad(r, P)
            val += p * v
            G = G + p * g
        return val, G
    if isinstance(reg, NetworkRegularizer):
        return reg.value_grad(P)
    return reg.value(P), reg.grad(P)


###############################################################################
# AdaGrad + epoch loop (MatFac.jl fit!, EXTERNAL -- restated)
###############################################################################


class AdaGrad:
    """Flux 0.13.13 ``AdaGrad`` with the view shim of src/optimizers.jl:6-13:
    acc starts at epsilon; acc += g^2; p -= eta * g / (sqrt(acc) + epsilon).
    State is keyed by parameter name and persists across LR-halving restarts."""

    def __init__(self, eta=1.0, epsilon=1e-8):
        self.eta = float(eta)
        self.epsilon = float(epsilon)
        self.acc: Dict[str, np.ndarray] = {}

    def apply(self, name, p, g):
        acc = self.acc.get(name)
        if acc is None:
            acc = self.acc[name] = np.full_like(p, self.epsilon)
        acc += g * g
        p -= p.dtype.type(self.eta) * g / (np.sqrt(acc) + p.dtype.type(self.epsilon))


###############################################################################
# Staging passes that bracket the hot loop   (src/fit.jl:125-187)
###############################################################################


def link_col_sqerr(m: OracleModel, D: np.ndarray) -> np.ndarray:
    """MF.link_col_sqerr as used at src/fit.jl:138-139: per column sum_i (D_ij - forward_ij)^2 over finite
    entries (EXTERNAL; identity link assumed for every noise model)."""
    Z = forward(m)
    return np.where(np.isfinite(D), (D - Z) ** 2, 0.0).sum(axis=0)


def column_ssq_grads(m: OracleModel, D: np.ndarray) -> np.ndarray:
    """MF.batched_column_ssq_grads as used at src/fit.jl:166-168: per column sum_i (dl_ij/dz_ij)^2."""
    Z = forward(m)
    out = np.zeros(D.shape[1])
    w = m.noise.weights
    for cr, dist, th in zip(m.noise.col_ranges, m.noise.dists, m.noise.thresholds):
        sl = slice(cr.start, cr.stop)
        _, g = noise_loss_grad(dist, Z[:, sl], D[:, sl], th)
        out[sl] = ((g * w[sl][None, :]) ** 2).sum(axis=0)
    return out


###############################################################################
# Between hot-loop calls: empirical-Bayes re-weighting, factor re-ordering, post-processing
# (reference: src/regularizers.jl reweight_eb! / reorder_reg! methods, src/fit.jl:504-555)
###############################################################################

def reweight_eb(reg, x, mixture_p=1.0):
    """``reweight_eb!``.  x: K x n matrix, vector, BatchArray, or the list of the four layer parameters
    [logsigma, logdelta, mu, theta] for the layer SequenceReg (a plain list of regs here)."""
    if isinstance(reg, ZeroReg):                                            # regularizers.jl:941
        return
    if isinstance(reg, L2Regularizer):                                      # :39-51  (largest singular value)^2
        X = np.atleast_2d(np.asarray(x, float))
        reg.weights = np.full(len(reg.weights), mixture_p / np.linalg.svd(X, compute_uv=False)[0] ** 2)
    elif isinstance(reg, SelectiveL1Reg):                                   # :149-159
        sel = reg.l1_idx * np.asarray(x, float)
        var_x = (sel ** 2).mean(axis=1) - sel.mean(axis=1) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            w = mixture_p * np.sqrt(2.0 / var_x)
        w[~np.isfinite(w)] = 1.0
        reg.weight = w
    elif isinstance(reg, NetworkRegularizer):                               # :313-328
        row_precs = mixture_p / np.var(np.asarray(x, float), axis=1, ddof=1)
        for k, ratio in enumerate(row_precs / reg.cur_weights):
            reg.AA[k] = reg.AA[k] * ratio
            reg.AB[k] = reg.AB[k] * ratio
            reg.BB[k] = reg.BB[k] * ratio
        reg.cur_weights = row_precs
    elif isinstance(reg, GroupRegularizer):                                 # :406-420
        X = np.asarray(x, float)
        reg.group_weights = [np.full(X.shape[0], mixture_p / np.linalg.svd(X[:, r.start:r.stop], compute_uv=False)[0] ** 2)
                             for r in reg.group_idx]
    elif isinstance(reg, ColParamReg):                                      # :490-497
        v = np.asarray(x, float)
        reg.centers = [float(np.mean(v[r.start:r.stop])) for r in reg.col_ranges]
        reg.weights = [float(mixture_p * np.float32(1e-1 + 0.5) / (np.float32(1e-1) + np.float32(0.5) * np.var(v[r.start:r.stop], ddof=1)))
                       for r in reg.col_ranges]
    elif isinstance(reg, ARDRegularizer):                                   # :588-609
        reg.alpha = [0.001] * len(reg.alpha)
        reg.beta = [0.001] * len(reg.beta)
    elif isinstance(reg, CompositeRegularizer):                             # :634-638
        for r, p in zip(reg.regularizers, reg.mixture_p):
            reweight_eb(r, x, mixture_p=p * mixture_p)
    elif isinstance(reg, BatchArrayReg):                                    # :818-840
        reg.centers = [v.mean(axis=1) for v in x.values]
        with np.errstate(divide="ignore", invalid="ignore"):
            reg.weights = [mixture_p / np.var(v, axis=1, ddof=1) for v in x.values]
        for w, cr in zip(reg.weights, x.col_ranges):
            w[~np.isfinite(w)] = 1.0 + 0.5 * len(cr)
    elif isinstance(reg, list):                                             # SequenceReg, :928-932
        for r, param in zip(reg, x):
            reweight_eb(r, param, mixture_p=mixture_p)
    else:
        raise TypeError(type(reg).__name__)


def reorder_reg(reg, perm):
    """``reorder_reg!`` methods (regularizers.jl:5,53,161,330,449,645; featureset_ard.jl:68); perm 0-based."""
    perm = list(perm)
    if isinstance(reg, L2Regularizer):
        reg.weights = reg.weights[perm]
    elif isinstance(reg, SelectiveL1Reg):
        reg.l1_idx = reg.l1_idx[perm, :]
    elif isinstance(reg, NetworkRegularizer):
        reg.AA = [reg.AA[k] for k in perm]
        reg.AB = [reg.AB[k] for k in perm]
        reg.BB = [reg.BB[k] for k in perm]
        reg.x_virtual = [reg.x_virtual[k] for k in perm]
        reg.cur_weights = reg.cur_weights[perm]
    elif isinstance(reg, GroupRegularizer):
        reg.group_weights = [w[perm] for w in reg.group_weights]
    elif isinstance(reg, CompositeRegularizer):
        for r in reg.regularizers:
            reorder_reg(r, perm)
    elif isinstance(reg, FeatureSetARDReg):
        reg.beta = reg.beta[perm, :]
        reg.A = [A[:, perm] for A in reg.A]
        for opt in reg.A_opts:
            opt.ssq_grad = opt.ssq_grad[:, perm]
            opt.lam = opt.lam[perm]


def init_ordinal_thresholds(m: OracleModel, D: np.ndarray) -> None:
    """``init_ordinal_thresholds!`` (src/fit.jl:222-246) with ``rec_set_thresholds`` (:190-219) and ``inv_logistic``
    (src/util.jl:8-10).  ``p_vec[2:1-end]`` in the reference is an empty range, so only the outermost pair of level
    probabilities is ever used -- which is all a three-level model has."""
    def inv_logistic(x):
        return math.log(0.5 + 0.99 * (x / (1.0 - x) - 0.5))
    for r, dist, th in zip(m.noise.col_ranges, m.noise.dists, m.noise.thresholds):
        if dist != "ordinal3":
            continue
        levels = len(th) - 1
        block = D[:, r.start:r.stop]
        p = np.array([float(np.sum(block == k)) + 1.0 for k in range(1, levels + 1)])
        p /= p.sum()
        if len(p) >= 2:
            th[1:-1] = [inv_logistic(p[0]), inv_logistic(1.0 - p[-1])]


def whiten(m: OracleModel, feature_views) -> None:
    """``whiten!`` (src/fit.jl:504-528)."""
    x_rms = np.sqrt(np.mean(m.X * m.X, axis=1, keepdims=True))
    m.X = m.X / x_rms
    m.Y = m.Y * x_rms
    for cr in ids_to_ranges(list(feature_views)):
        y_rms_max = np.sqrt(np.mean(m.Y[:, cr.start:cr.stop] ** 2, axis=1)).max()
        if y_rms_max > 0:
            m.Y[:, cr.start:cr.stop] /= y_rms_max
            m.logsigma[cr.start:cr.stop] += math.log(y_rms_max)
        else:
            m.Y[:, cr.start:cr.stop] = 0.0
            m.logsigma[cr.start:cr.stop] = np.float32(-1e9)


def rotate_by_svd(m: OracleModel) -> None:
    """``rotate_by_svd!`` (src/fit.jl:531-544)."""
    U, s, Vt = np.linalg.svd(m.Y, full_matrices=False)
    m.Y = s[:, None] * Vt
    m.X = (m.X.T @ U).T


def reorder_by_importance(m: OracleModel):
    """``reorder_by_importance!`` (src/fit.jl:547-555): sortperm(Y_ssq, rev=true) is stable."""
    y_ssq = np.sum(m.Y * m.Y, axis=1)
    idx = sorted(range(len(y_ssq)), key=lambda k: -y_ssq[k])
    m.X = m.X[idx, :]
    m.Y = m.Y[idx, :]
    reorder_reg(m.Y_reg, idx)
    reorder_reg(m.X_reg, idx)
    return idx


def compute_M_estimates(m: OracleModel, D: np.ndarray, lr=0.1, max_epochs=500, rel_tol=1e-5, abs_tol=1e-3):
    """``MF.compute_M_estimates`` as ``init_mu!`` calls it (src/fit.jl:82-104) [EXTERNAL, INFERRED; parity
    unpinned]: per column, the shift minimising the column's noise-model loss.  Restated as the full-batch
    AdaGrad loop of ``mf_fit`` on a model reduced to its ColShift layer (Z = 0, sigma = 1, no batch layers,
    no penalties), started at mu = 0.  Returns (M_estimates, history)."""
    K, M = m.X.shape
    N = m.Y.shape[1]
    mm = OracleModel(X=np.zeros((K, M)), Y=np.zeros((K, N)), logsigma=np.zeros(N), mu=np.zeros(N),
                     logdelta=None, theta=None, noise=m.noise)
    mm.frozen = [True, True, False, True]
    h = mf_fit(mm, D, AdaGrad(lr), max_epochs=max_epochs, rel_tol=rel_tol, abs_tol=abs_tol, update_col_layers=True)
    return mm.mu, h


def init_mu(m: OracleModel, D: np.ndarray, lr_mu=0.1, max_epochs=500):
    """``init_mu!`` (src/fit.jl:82-104): mu <- M-estimates (rel_tol 1e-5, abs_tol 1e-3 as hard-coded there)."""
    est, h = compute_M_estimates(m, D, lr=lr_mu, max_epochs=max_epochs)
    m.mu = est.astype(m.mu.dtype)
    return h


def init_logsigma(m: OracleModel, D: np.ndarray) -> None:
    """src/fit.jl:125-148: with X = Y = 0, logsigma = log sqrt(col sq. error / number of finite entries)."""
    X0, Y0 = m.X, m.Y
    m.X, m.Y = np.zeros_like(X0), np.zeros_like(Y0)
    with np.errstate(divide="ignore", invalid="ignore"):
        col_vars = link_col_sqerr(m, D) / np.isfinite(D).sum(axis=0)
        m.X, m.Y = X0, Y0
        m.logsigma = np.log(np.sqrt(col_vars)).astype(m.logsigma.dtype)


def reweight_col_losses(m: OracleModel, D: np.ndarray) -> None:
    """src/fit.jl:151-187: weights = 1 / (sqrt(ssq_grads / M) * exp(logsigma)), non-finite -> 1."""
    M, N = D.shape
    m.noise.weights = np.ones(N, dtype=m.noise.weights.dtype)
    X0, Y0 = m.X, m.Y
    m.X, m.Y = np.zeros_like(X0), np.zeros_like(Y0)
    ssq = column_ssq_grads(m, D)
    m.X, m.Y = X0, Y0
    with np.errstate(divide="ignore", invalid="ignore"):
        w = 1.0 / (np.sqrt(ssq / M) * np.exp(m.logsigma))
    w[~np.isfinite(w)] = 1.0
    m.noise.weights = w.astype(m.noise.weights.dtype)


def theta_mom(theta_values):
    """src/fit.jl:297-301."""
    return ([v.mean(axis=1, keepdims=True) for v in theta_values],
            [v.var(axis=1, ddof=1, keepdims=True) for v in theta_values])


def delta2_mom(delta2_values):
    """src/fit.jl:303-311."""
    mean = [v.mean(axis=1, keepdims=True) for v in delta2_values]
    var = [v.var(axis=1, ddof=1, keepdims=True) for v in delta2_values]
    alpha = [2.0 + (mm * mm) / (vv + 1e-9) for mm, vv in zip(mean, var)]
    beta = [mm * (a - 1.0) for mm, a in zip(mean, alpha)]
    return alpha, beta


def theta_delta_em(m: OracleModel, delta2, sigma2, D, update_priors=True, batch_em_max_iter=100, batch_em_rtol=1e-8):
    """src/fit.jl:326-375 (sqerr_func: squared error in link space, identity link, over finite entries)."""
    theta = m.theta
    delta2 = [np.array(d, dtype=np.float64) for d in delta2]
    theta_lsq = [v.copy() for v in theta.values]
    batch_sizes = ba_map(lambda d: np.isfinite(d).astype(np.float64), theta, D)

    def nans_to(arrs, val):
        for a in arrs:
            a[~np.isfinite(a)] = val

    diffs = []
    for it in range(batch_em_max_iter):
        if update_priors or it == 0:
            theta_mean, theta_var = theta_mom(theta.values)
            alpha, beta = delta2_mom(delta2)
        theta_old = [v.copy() for v in theta.values]
        with np.errstate(divide="ignore", invalid="ignore"):
            theta.values = [(e * d2 * sigma2[None, cr.start:cr.stop] + tl * bs * vt) / (sigma2[None, cr.start:cr.stop] * d2 + bs * vt)
                            for e, vt, d2, tl, bs, cr in zip(theta_mean, theta_var, delta2, theta_lsq, batch_sizes, theta.col_ranges)]
        nans_to(theta.values, 0.0)
        Z = forward(m)
        sqerr = ba_map(lambda z, d: np.where(np.isfinite(d), (d - z) ** 2, 0.0), theta, Z, D)
        nans_to(sqerr, 0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            delta2 = [(b + 0.5 * (sq / sigma2[None, cr.start:cr.stop])) / (a + 0.5 * bs - 1.0)
                      for a, b, sq, bs, cr in zip(alpha, beta, sqerr, batch_sizes, theta.col_ranges)]
        nans_to(delta2, 1.0)
        num = sum(((v - o) ** 2).sum() for v, o in zip(theta.values, theta_old))
        den = sum((v * v).sum() for v in theta.values)
        diffs.append(num / den)
        if diffs[-1] < batch_em_rtol:
            break
    return theta.values, delta2, diffs


def _apply_updates(m, opt, out, update_X, update_Y, update_col_layers, update_noise_models):
    if update_X:
        opt.apply("X", m.X, out["dX"])
    if update_Y:
        opt.apply("Y", m.Y, out["dY"])
    if update_col_layers:
        if not m.frozen[0]:
            opt.apply("logsigma", m.logsigma, out["dlogsigma"])
        if not m.frozen[2]:
            opt.apply("mu", m.mu, out["dmu"])
        if m.logdelta is not None and not m.frozen[1]:
            for v, (p, g) in enumerate(zip(m.logdelta.values, out["dlogdelta"])):
                opt.apply(f"logdelta{v}", p, g)
        if m.theta is not None and not m.frozen[3]:
            for v, (p, g) in enumerate(zip(m.theta.values, out["dtheta"])):
                opt.apply(f"theta{v}", p, g)
    if update_noise_models:
        # D7: the interior ordinal thresholds are the trainable noise-model parameters
        for r, th in enumerate(m.noise.thresholds):
            if th is not None:
                inner = th[1:3].copy()
                opt.apply(f"thresholds{r}", inner, out["dthresholds"][r].astype(inner.dtype))
                th[1:3] = inner


def mf_fit(m: OracleModel, D, opt: AdaGrad, max_epochs=1000, epoch=1, rel_tol=1e-5, abs_tol=1e-5,
           update_X=False, update_Y=False, update_col_layers=False, capacity=10 ** 8,
           callback=None, update_noise_models=False, alternating=False):
    """One ``MF.fit!`` call as wrapped by ``mf_fit!`` (src/fit.jl:9-38).

    ``update_noise_models`` (src/fit.jl:14, true in every call of the reference; D7): train the interior ordinal
    thresholds with the same optimiser.  ``alternating`` (D1): instead of one simultaneous step from one pass, the
    epoch takes the column-side step (Y, column / batch layers, thresholds) from the first pass and the row-side step
    (X) from a SECOND pass at the new column-side parameters; the recorded loss is the first pass'.

    Assumed semantics: D1 simultaneous gradients from one pass;
    D2 the loss of epoch t is the one evaluated in that epoch's gradient pass
    (pre-update parameters); D3/D4 terminate when loss increases
    ("loss_increase"), |dloss| < abs_tol ("abs_tol"), |dloss/loss| < rel_tol
    ("rel_tol"), or the epoch counter passes max_epochs ("max_epochs"); a
    non-finite loss gives "nonfinite".  Returns the history Dict with
    ``term_code``, ``epochs`` and per-epoch loss components."""
    h = {"term_code": "max_epochs", "epochs": epoch, "loss": [], "components": []}
    prev = None
    while epoch <= max_epochs:
        out = total_loss_grads(m, D, capacity)
        loss = out["loss"]
        h["loss"].append(loss)
        h["components"].append(out["components"])
        h["epochs"] = epoch
        if callback is not None:
            callback(epoch, out)
        if not math.isfinite(loss):
            h["term_code"] = "nonfinite"
            break
        if prev is not None:
            d = prev - loss
            if d < 0:
                h["term_code"] = "loss_increase"
                break
            if abs(d) < abs_tol:
                h["term_code"] = "abs_tol"
                break
            if abs(d / loss) < rel_tol:
                h["term_code"] = "rel_tol"
                break
        prev = loss
        if alternating:
            _apply_updates(m, opt, out, False, update_Y, update_col_layers, update_noise_models)
            if update_X:
                out2 = total_loss_grads(m, D, capacity)
                _apply_updates(m, opt, out2, True, False, False, False)
        else:
            _apply_updates(m, opt, out, update_X, update_Y, update_col_layers, update_noise_models)
        epoch += 1
    return h


def mf_fit_adapt_lr(m: OracleModel, D, lr=1.0, min_lr=0.001, max_epochs=1000, **kw):
    """src/fit.jl:46-75: on "loss_increase" halve eta (AdaGrad state kept) and
    resume at h["epochs"]; stop when eta < min_lr or any other term code."""
    opt = AdaGrad(lr)
    epoch = 1
    history = []
    while epoch <= max_epochs:
        h = mf_fit(m, D, opt, max_epochs=max_epochs, epoch=epoch, **kw)
        h["lr"] = opt.eta
        history.append(h)
        if h["term_code"] == "loss_increase":
            opt.eta *= 0.5
            if opt.eta < min_lr:
                break
            epoch = h["epochs"]
        else:
            break
    return history


###############################################################################
# Synthetic inputs  (restating src/simulate_params.jl; numpy RNG, not Xoshiro)
###############################################################################

VIEW_MU_MEAN = {"mrnaseq": 10.0, "methylation": 0.0, "cna": 0.0, "mutation": -1.5, "counts": 1.0}
VIEW_MU_STD = {"mrnaseq": 2.0, "methylation": 0.1, "cna": 0.1, "mutation": 0.1, "counts": 0.5}
VIEW_LOGSIGMA_MEAN = {"mrnaseq": 0.5, "methylation": 0.1, "cna": math.log(2.0), "mutation": 0.0,
                      "counts": -1.0}
VIEW_LOGSIGMA_STD = {"mrnaseq": 0.1, "methylation": 0.1, "cna": 0.001, "mutation": 0.001,
                     "counts": 0.1}
# ("counts" is this repo's poisson view: the reference has no poisson sampler)


def simulate_model(M, N_per_view: Dict[str, Tuple[str, int]], K, seed, batch_views: Sequence[str] = (),
                   n_batches=0, n_conditions=0, missing=0.0, dtype=np.float64, noise=0.1, sort_batches=False):
    """Seeded synthetic inputs following simulate_params!/simulate_data!
    (src/simulate_params.jl:193-255).  ``N_per_view``: view -> (distribution, n cols),
    already listed in the constructor's sorted (distribution, view) order.
    Returns (OracleModel, D, meta)."""
    rng = np.random.default_rng(seed)
    views, dists = [], []
    for v, (d, n) in N_per_view.items():
        views += [v] * n
        dists += [d] * n
    order = sorted(range(len(views)), key=lambda i: (dists[i], views[i]))   # model.jl:50
    assert order == list(range(len(views))), "list views in (distribution, view) sorted order"
    N = len(views)
    X = rng.standard_normal((K, M))
    Y = rng.standard_normal((K, N))
    logsigma = np.zeros(N)
    mu = np.zeros(N)
    for cr, v in zip(ids_to_ranges(views), unique_in_order(views)):
        n = len(cr)
        logsigma[cr.start:cr.stop] = rng.standard_normal(n) * VIEW_LOGSIGMA_STD[v] + VIEW_LOGSIGMA_MEAN[v]
        mu[cr.start:cr.stop] = rng.standard_normal(n) * VIEW_MU_STD[v] + VIEW_MU_MEAN[v]
    logsigma -= math.log(math.sqrt(K))                                       # simulate_params.jl:212
    logdelta = theta = None
    batch_dict = None
    if batch_views:
        batch_dict = {v: [int(b) for b in rng.integers(0, n_batches, size=M)] for v in batch_views}
        if sort_batches:      # samples grouped by batch, as plate / centre ids of a sorted cohort are
            batch_dict = {v: sorted(b) for v, b in batch_dict.items()}
        col_ranges = ids_to_ranges(views)
        unq = unique_in_order(views)

        def mk():
            vds = []
            for v, cr in zip(unq, col_ranges):
                vds.append({b: np.zeros(len(cr)) for b in unique_in_order(batch_dict[v])} if v in batch_dict else {})
            ba = BatchArray.construct(views, batch_dict, vds)
            for i, val in enumerate(ba.values):                              # simulate_params.jl:96-101
                centers = rng.standard_normal(val.shape[0]) * 0.25
                ba.values[i] = centers[:, None] + rng.standard_normal(val.shape) * 0.25
            return ba
        logdelta, theta = mk(), mk()
    nm = NoiseModel.from_distributions(dists)
    m = OracleModel(X=X, Y=Y, logsigma=logsigma, mu=mu, logdelta=logdelta, theta=theta, noise=nm)
    D = forward(m)
    for cr, d in zip(nm.col_ranges, nm.dists):
        sl = slice(cr.start, cr.stop)
        if d == "normal":
            D[:, sl] += rng.standard_normal(D[:, sl].shape) * noise          # simulate_params.jl:240-242
        elif d == "bernoulli":
            D[:, sl] = (rng.random(D[:, sl].shape) < _sigmoid(D[:, sl])).astype(float)
        elif d == "poisson":
            D[:, sl] = rng.poisson(np.exp(np.minimum(D[:, sl], 10.0))).astype(float)
        elif d == "bernoulli_sq_hinge":
            D[:, sl] = (D[:, sl] > 0).astype(float)                          # :224-228
        elif d in ("ordinal3", "ordinal_sq_hinge3"):
            t = nm.thresholds[nm.dists.index(d)]
            D[:, sl] = 1.0 + (D[:, sl] > t[1]) + (D[:, sl] > t[2])           # :230-238
    if missing > 0:
        D[rng.random(D.shape) < missing] = np.nan
    conditions = None
    if n_conditions:
        conditions = list(np.sort(rng.integers(0, n_conditions, size=M)))
    # perturb the parameters so the fit does not start at the generating optimum
    m.X = rng.standard_normal((K, M))
    m.Y = rng.standard_normal((K, N))
    meta = {"views": views, "dists": dists, "batch_dict": batch_dict, "conditions": conditions}
    for name in ("X", "Y", "logsigma", "mu"):
        setattr(m, name, getattr(m, name).astype(dtype))
    if logdelta is not None:
        for ba in (m.logdelta, m.theta):
            ba.values = [v.astype(dtype) for v in ba.values]
    return m, D.astype(dtype), meta
