"""Index / CSR bookkeeping of the hot path (host side; must be bit-exact with the
reference: src/util.jl:140-262, 269-314, 453-511).  Ranges are 0-based half-open
``range`` objects; ``julia_ranges`` renders them the reference's way."""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

VALID_LOSSES = ["normal", "bernoulli", "bernoulli_sq_hinge", "poisson", "ordinal3", "ordinal_sq_hinge3"]
DIST_CODE = {"normal": 0, "bernoulli": 1, "poisson": 2, "ordinal3": 3, "bernoulli_sq_hinge": 4,
             "ordinal_sq_hinge3": 5}


def unique(seq) -> list:
    """Julia ``unique``: keeps first-appearance order."""
    return list(dict.fromkeys(seq))


def is_contiguous(vec) -> bool:
    """True when equal values form unbroken runs (src/util.jl:140-156)."""
    vec = list(vec)
    closed = set()
    for prev, cur in zip(vec, vec[1:]):
        if cur in closed:
            return False
        if cur != prev:
            closed.add(prev)
    return True


def ids_to_ranges(id_vec) -> List[range]:
    """Runs of equal ids -> ranges, in order of first appearance (src/util.jl:187-197)."""
    id_vec = list(id_vec)
    if not is_contiguous(id_vec):
        raise AssertionError("IDs in id_vec need to appear in contiguous chunks.")
    ranges, start = [], 0
    for i in range(1, len(id_vec) + 1):
        if i == len(id_vec) or id_vec[i] != id_vec[start]:
            ranges.append(range(start, i))
            start = i
    return ranges


def julia_ranges(ranges: Iterable[range]) -> List[Tuple[int, int]]:
    return [(r.start + 1, r.stop) for r in ranges]


def ids_to_index(id_vec) -> Tuple[np.ndarray, list]:
    """(ordinal of each id in ``unique`` order as int32, the unique ids) -- the dense
    form of ``ids_to_ind_mat`` (src/util.jl:200-210): ind[i, idx[i]] == true."""
    unq = unique(id_vec)
    lut = {u: i for i, u in enumerate(unq)}
    return np.fromiter((lut[v] for v in id_vec), dtype=np.int32, count=len(id_vec)), unq


def ids_to_ind_mat(id_vec) -> np.ndarray:
    idx, unq = ids_to_index(id_vec)
    ind = np.zeros((len(idx), len(unq)), dtype=bool)
    ind[np.arange(len(idx)), idx] = True
    return ind


def subset_ranges(ranges: Sequence[range], rng: range):
    """Intersect sorted, disjoint ``ranges`` with ``rng`` (src/util.jl:214-253).
    Returns (clipped ranges, first kept index, last kept index); empty -> ([], 0, -1)."""
    kept = [(i, range(max(r.start, rng.start), min(r.stop, rng.stop)))
            for i, r in enumerate(ranges)]
    kept = [(i, r) for i, r in kept if len(r) > 0]
    if not kept:
        return [], 0, -1
    return [r for _, r in kept], kept[0][0], kept[-1][0]


def value_to_idx(values) -> Dict:
    return {v: i for i, v in enumerate(values)}


def edgelist_to_spmat(edgelist, node_to_idx, epsilon=0.0) -> sp.csr_matrix:
    """Signed-graph Laplacian (src/util.jl:269-314): duplicated unordered pairs keep the
    last weight; diagonal = epsilon + sum |w|; off-diagonal = -w."""
    n = len(node_to_idx)
    latest = {}
    for u, v, w in edgelist:
        a, b = node_to_idx[u], node_to_idx[v]
        latest[(a, b) if a >= b else (b, a)] = float(w)
    if latest:
        ij = np.array(list(latest.keys()), dtype=np.int64)
        w = np.array(list(latest.values()), dtype=np.float64)
    else:
        ij = np.zeros((0, 2), dtype=np.int64)
        w = np.zeros(0)
    diag = np.full(n, float(epsilon))
    np.add.at(diag, ij[:, 0], np.abs(w))
    np.add.at(diag, ij[:, 1], np.abs(w))
    rows = np.concatenate([np.arange(n), ij[:, 0], ij[:, 1]])
    cols = np.concatenate([np.arange(n), ij[:, 1], ij[:, 0]])
    vals = np.concatenate([diag, -w, -w])
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()


def csc_select(A, rng1: range, rng2: range):
    """Block A[rng1, rng2] (src/util.jl:494-511)."""
    return sp.csr_matrix(A.tocsr()[rng1.start:rng1.stop, :][:, rng2.start:rng2.stop])


def featuresets_to_csc(feature_ids, feature_sets) -> sp.csr_matrix:
    """L x N membership matrix, row l scaled by 1/sqrt(|set l|) (src/util.jl:453-477)."""
    col_of = value_to_idx(list(feature_ids))
    rows, cols, vals = [], [], []
    for l, fs in enumerate(feature_sets):
        fs = list(fs)
        s = np.float32(1.0 / math.sqrt(len(fs)))
        rows += [l] * len(fs)
        cols += [col_of[f] for f in fs]
        vals += [s] * len(fs)
    return sp.coo_matrix((np.asarray(vals, np.float32), (rows, cols)),
                         shape=(len(feature_sets), len(col_of))).tocsr()


def get_all_nodes(edgelist) -> set:
    nodes = set()
    for e in edgelist:
        nodes.update((e[0], e[1]))
    return nodes


def compute_nongraph_nodes(feature_ids, edgelists) -> List[set]:
    everything = set(feature_ids)
    return [everything - get_all_nodes(el) for el in edgelists]
