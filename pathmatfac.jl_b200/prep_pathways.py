"""Pathway inputs of the model constructor: SIF files -> ``feature_graphs`` (the per-factor edge lists behind the
NetworkRegularizer on Y, src/regularizers.jl:187-240) and ``feature_sets_dict`` (the feature sets behind the feature-set
ARD prior, src/featureset_ard.jl:111-132).  Restates src/prep_pathways.jl and the graph helpers of src/util.jl:331-428;
pinned by the reference's own ``preprocess_tests`` (test/runtests.jl:456-616), one of the three test functions its
``main()`` runs.  Host-only code on the input side of the hot path: BASELINE config 1 names it ("pathways from
test_pathway.sif").

A SIF row is ``[source, code, target]`` with a two-character code: the target's level (``a`` activation, ``d`` dna,
``t`` mrna, ``p`` protein) and the sign (``>`` = +1, ``|`` = -1).  Edges are ``[u, v, weight]`` lists as in the
reference."""
from __future__ import annotations

import copy
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

DOGMA_ORDER = ["dna", "mrna", "protein", "activation"]                         # src/util.jl:117
PWY_SIF_CODE = {"a": "activation", "d": "dna", "t": "mrna", "p": "protein", ">": 1, "|": -1}   # src/util.jl:120-126


# ---- SIF input (src/prep_pathways.jl:13-25) --------------------------------------------------------------------------

def read_sif_file(sif_file: str) -> List[List[str]]:
    """One row per line, tab separated (entity names contain spaces and slashes), no header."""
    rows = []
    with open(sif_file) as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if line.strip():
                rows.append(line.split("\t"))
    return rows


def read_all_sif_files(sif_files: Sequence[str]) -> List[List[List[str]]]:
    return [read_sif_file(p) for p in sif_files]


def _as_sif_data(pwy_sifs):
    """The reference dispatches on the element type: paths are read, row lists are taken as they are (:61-69)."""
    return [read_sif_file(s) if isinstance(s, str) else s for s in pwy_sifs]


# ---- edge lists (src/prep_pathways.jl:35-133) ------------------------------------------------------------------------

def sif_to_edgelist(pwy_sif) -> List[list]:
    """Source nodes are activations; the interaction code gives the target's level and the sign."""
    edges = []
    for u, code, v in pwy_sif:
        edges.append([f"{u}_{PWY_SIF_CODE['a']}", f"{v}_{PWY_SIF_CODE[code[0]]}", PWY_SIF_CODE[code[1]]])
    return edges


def sifs_to_edgelists(pwy_sifs) -> List[List[list]]:
    return [sif_to_edgelist(s) for s in _as_sif_data(pwy_sifs)]


def construct_dogma_edges(dogma_proteins: Iterable[str], dogma_order=DOGMA_ORDER) -> List[list]:
    """dna -> mrna -> protein -> activation for every gene, weight 1."""
    return [[f"{prot}_{a}", f"{prot}_{b}", 1.0] for prot in dogma_proteins for a, b in zip(dogma_order[:-1], dogma_order[1:])]


def construct_data_edges(feature_ids: Sequence[str], feature_weights: Sequence[float]) -> List[list]:
    """Feature ``GENE_level_id`` hangs off its dogma node ``GENE_level``."""
    return [[feat, "_".join(feat.split("_")[:2]), w] for feat, w in zip(feature_ids, feature_weights)]


def tag_nodes(edgelist, tag="_activation"):
    out = copy.deepcopy(edgelist)
    for e in out:
        e[0], e[1] = f"{e[0]}{tag}", f"{e[1]}{tag}"
    return out


# ---- graph helpers (src/util.jl:331-428) -----------------------------------------------------------------------------

def edgelist_to_dict(edgelist) -> Dict:
    g: Dict = {}
    for u, v, w in edgelist:
        g.setdefault(u, {})[v] = w
        g.setdefault(v, {})[u] = w
    return g


def dict_to_edgelist(graph: Dict) -> List[list]:
    """Every undirected edge once (the reverse entry is dropped as the walk goes, as in the reference)."""
    out = []
    for u in list(graph):
        for v, w in list(graph[u].items()):
            out.append([u, v, w])
            if v != u:
                graph[v].pop(u, None)
    return out


def prune_leaves(edgelist, except_=None) -> List[list]:
    """``prune_leaves!`` (src/util.jl:369-404): remove nodes of degree <= 1 recursively, never those in ``except_``.
    Returns the pruned edge list.  NOTE the reference's function rebinds its argument instead of mutating it, so callers
    that ignore the return value -- ``extend_pathways`` does (src/prep_pathways.jl:168) -- keep the unpruned list."""
    keep = set(except_ or ())
    g = edgelist_to_dict(edgelist)
    frontier = {n for n, nb in g.items() if len(nb) < 2 and n not in keep}
    while frontier:
        leaf = frontier.pop()
        if leaf in g and len(g[leaf]) < 2:
            for nb in list(g[leaf]):
                if nb not in keep and nb != leaf:
                    frontier.add(nb)
                g[nb].pop(leaf, None)
            del g[leaf]
    return dict_to_edgelist(g)


def get_all_nodes(edgelist) -> set:
    return {e[0] for e in edgelist} | {e[1] for e in edgelist}


def get_all_entities(edgelist) -> set:
    """Node names up to the first underscore: the genes / complexes of a pathway."""
    return {str(e[0]).split("_")[0] for e in edgelist} | {str(e[1]).split("_")[0] for e in edgelist}


# ---- pathways extended by the central dogma and the data features (src/prep_pathways.jl:136-233) ----------------------

def construct_pwy_feature_ids(feature_genes, feature_dogmas, feature_ids) -> List[str]:
    return [f"{g}_{d}_{i}" for g, d, i in zip(feature_genes, feature_dogmas, feature_ids)]


def extend_pathways(pwy_edgelists, feature_ids: Sequence[str], feature_weights: Sequence[float], prune: bool = False):
    """Per pathway: dogma edges of the genes it shares with the data, the data edges of their features, then the pathway's
    own edges.  ``prune=False`` is the reference's behaviour (its ``prune_leaves!`` call has no effect, see
    ``prune_leaves``); ``prune=True`` applies the pruning the call was meant to do."""
    feature_id_set = set(feature_ids)
    feature_genes = [fid.split("_")[0] for fid in feature_ids]
    gene_set = set(feature_genes)
    out = []
    for pwy in pwy_edgelists:
        relevant = get_all_entities(pwy) & gene_set
        idx = [i for i, g in enumerate(feature_genes) if g in relevant]
        edges = construct_dogma_edges(sorted(relevant))
        edges += construct_data_edges([feature_ids[i] for i in idx], [feature_weights[i] for i in idx])
        edges += [list(e) for e in pwy]
        out.append(prune_leaves(edges, except_=feature_id_set) if prune else edges)
    return out


def prep_pathway_graphs(pwy_sifs, feature_genes, feature_dogmas, feature_ids=None, feature_weights=None, prune=False
                        ) -> Tuple[List[List[list]], List[str]]:
    """``prep_pathway_graphs`` (src/prep_pathways.jl:205-233): (edge lists for ``feature_graphs``, the new feature ids
    ``GENE_level_id`` the graphs refer to -- pass them as the model's ``feature_ids``)."""
    N = len(feature_genes)
    assert len(feature_dogmas) == N, "`feature_genes` and `feature_dogmas` must have identical length"
    assert all(d in DOGMA_ORDER for d in feature_dogmas)
    if feature_weights is None:
        feature_weights = [1.0] * N
    if feature_ids is None:
        feature_ids = list(range(1, N + 1))
    new_ids = construct_pwy_feature_ids(feature_genes, feature_dogmas, feature_ids)
    graphs = extend_pathways(sifs_to_edgelists(pwy_sifs), new_ids, list(feature_weights), prune=prune)
    return graphs, new_ids


# ---- feature sets (src/prep_pathways.jl:238-354) ---------------------------------------------------------------------

def sif_to_nodeset(pwy_sif) -> set:
    return {row[0] for row in pwy_sif} | {row[2] for row in pwy_sif}


def sifs_to_nodesets(pwy_sifs) -> List[set]:
    return [sif_to_nodeset(s) for s in _as_sif_data(pwy_sifs)]


def prep_pathway_featuresets(pwy_sifs, feature_genes, feature_views=None, feature_ids=None, featureset_ids=None):
    """``prep_pathway_featuresets`` (both methods, src/prep_pathways.jl:272-354).

    Without ``feature_views``: (feature sets, feature ids, ids of the kept feature sets) -- a pathway that shares no gene
    with the data is dropped.  With ``feature_views``: the same per unique view, as dicts keyed by view (the
    ``feature_sets_dict`` / ``featureset_names`` of the model constructor)."""
    feature_genes = list(feature_genes)
    if feature_ids is None:
        feature_ids = [f"{g}_{i}" for i, g in enumerate(feature_genes, start=1)]
    feature_ids = list(feature_ids)
    if featureset_ids is None:
        featureset_ids = list(range(1, len(pwy_sifs) + 1))
    if feature_views is not None:
        feature_views = list(feature_views)
        sets_by_view, ids_by_view = {}, {}
        sif_data = _as_sif_data(pwy_sifs)
        for uv in dict.fromkeys(feature_views):                   # unique(), first-appearance order
            idx = [i for i, v in enumerate(feature_views) if v == uv]
            fs, _, kept = prep_pathway_featuresets(sif_data, [feature_genes[i] for i in idx],
                                                   feature_ids=[feature_ids[i] for i in idx], featureset_ids=featureset_ids)
            sets_by_view[uv], ids_by_view[uv] = fs, kept
        return sets_by_view, feature_ids, ids_by_view
    all_genes = set(feature_genes)
    gene_to_idx: Dict[str, List[int]] = {}
    for i, g in enumerate(feature_genes):
        gene_to_idx.setdefault(g, []).append(i)
    feature_sets, kept_ids = [], []
    for fid, ns in zip(featureset_ids, sifs_to_nodesets(pwy_sifs)):
        ns = ns & all_genes
        if ns:
            feature_sets.append({feature_ids[i] for node in ns for i in gene_to_idx.get(node, [])})
            kept_ids.append(fid)
    return feature_sets, feature_ids, kept_ids
