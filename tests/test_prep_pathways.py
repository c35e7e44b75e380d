"""The reference's ``preprocess_tests`` (test/runtests.jl:456-616) -- one of the three test functions its ``main()`` runs
-- and the graph-helper assertions of ``util_tests`` (:79-91), transcribed against the mirror's ``prep_pathways``; then the
"real pathway" block of ``reg_tests`` (:669-688): the NetworkRegularizer built from the prepared pathway has one block
row per data feature and one per unobserved node.  The SIF rows are the contents of the reference's
test/test_pathway.sif (written to a temporary file here)."""
import numpy as np

import pathmatfac_b200 as P
from pathmatfac_b200 import prep_pathways as PP
from oracle import pmf_oracle as O

SIF = [["MOL:GTP", "a>", "GRASP65/GM130/RAB1/GTP/PLK1"], ["METAPHASE", "a>", "PLK1"], ["PAK1", "a>", "PLK1"],
       ["RAB1A", "a>", "GRASP65/GM130/RAB1/GTP/PLK1"], ["PLK1", "a>", "SGOL1"], ["PLK1", "a>", "GRASP65/GM130/RAB1/GTP/PLK1"],
       ["PP2A-ALPHA B56", "a|", "SGOL1"]]                                                   # runtests.jl:461-468
GENES = ["PLK1", "PLK1", "PLK1", "PAK1", "PAK1", "PAK1", "SGOL1", "BRCA", "BRCA"]           # :490-493
ASSAYS = ["cna", "mutation", "mrnaseq", "rppa", "methylation", "mutation", "mrnaseq", "mrnaseq", "methylation"]
DOGMA = {"cna": "dna", "mutation": "dna", "mrnaseq": "mrna", "methylation": "mrna", "rppa": "protein"}
WEIGHT = {"cna": 1.0, "mutation": -1.0, "mrnaseq": 1.0, "methylation": -1.0, "rppa": 1.0}


def _sets(edges):
    return {frozenset(map(str, e)) for e in edges}          # Set(map(Set, edges)): direction and order do not matter


def _sif_file(tmp_path):
    p = tmp_path / "test_pathway.sif"
    p.write_text("".join("\t".join(r) + "\n" for r in SIF))
    return str(p)


def test_graph_helpers():
    """runtests.jl:70, 79-91."""
    edgelist = [["a", "b", 1], ["b", "c", 1], ["c", "d", -1], ["d", "a", -1]]
    d = PP.edgelist_to_dict(edgelist)
    assert d == {"a": {"b": 1, "d": -1}, "b": {"a": 1, "c": 1}, "c": {"b": 1, "d": -1}, "d": {"c": -1, "a": -1}}
    assert _sets(PP.dict_to_edgelist(d)) == _sets(edgelist)
    leafy = edgelist + [["d", "e", 1], ["e", "g", -1], ["c", "f", 1], ["f", "h", -1]]
    assert _sets(PP.prune_leaves(leafy, except_=["f"])) == _sets(edgelist + [["c", "f", 1]])
    assert len(leafy) == 8                                     # the argument is not modified (the reference rebinds it)


def test_prep_pathway_graphs(tmp_path):
    """runtests.jl:470-575."""
    path = _sif_file(tmp_path)
    pwy_edges = [[f"{u}_activation", f"{v}_activation", 1 if c == "a>" else -1] for u, c, v in SIF]
    dogmas = [DOGMA[a] for a in ASSAYS]
    weights = [WEIGHT[a] for a in ASSAYS]
    feature_ids = PP.construct_pwy_feature_ids(GENES, dogmas, range(1, 10))
    assert feature_ids[:3] == ["PLK1_dna_1", "PLK1_dna_2", "PLK1_mrna_3"] and feature_ids[-1] == "BRCA_mrna_9"
    relevant_genes = PP.get_all_entities(pwy_edges) & set(GENES)
    assert relevant_genes == {"PLK1", "PAK1", "SGOL1"}
    rel = [i for i, g in enumerate(GENES) if g in relevant_genes]
    dogma_edges = [[f"{g}_{a}", f"{g}_{b}", 1.0] for g in ("PLK1", "PAK1", "SGOL1")
                   for a, b in (("dna", "mrna"), ("mrna", "protein"), ("protein", "activation"))]
    data_edges = [["PLK1_dna", "PLK1_dna_1", 1.], ["PLK1_dna", "PLK1_dna_2", -1.], ["PLK1_mrna", "PLK1_mrna_3", 1.],
                  ["PAK1_protein", "PAK1_protein_4", 1.], ["PAK1_mrna", "PAK1_mrna_5", -1.], ["PAK1_dna", "PAK1_dna_6", -1.],
                  ["SGOL1_mrna", "SGOL1_mrna_7", 1.]]
    assert PP.read_sif_file(path) == SIF                                                     # :549-550
    assert _sets(PP.sif_to_edgelist(SIF)) == _sets(pwy_edges)                                # :552-553
    assert [tuple(e) for e in PP.sif_to_edgelist(SIF)] == [tuple(e) for e in pwy_edges]
    assert _sets(PP.construct_dogma_edges(relevant_genes)) == _sets(dogma_edges)             # :556-557
    assert _sets(PP.construct_data_edges([feature_ids[i] for i in rel], [weights[i] for i in rel])) == _sets(data_edges)
    # the reference's expected graph: data + dogma + pathway edges; its prune_leaves! call leaves the list as it is
    expect = _sets(data_edges + dogma_edges + pwy_edges)
    for sifs in ([SIF], [path]):                                                             # :564-575, rows or a file
        graphs, new_ids = PP.prep_pathway_graphs(sifs, GENES, dogmas, feature_weights=weights)
        assert new_ids == feature_ids and len(graphs) == 1 and _sets(graphs[0]) == expect
    pruned, _ = PP.prep_pathway_graphs([SIF], GENES, dogmas, feature_weights=weights, prune=True)
    nodes = PP.get_all_nodes(pruned[0])
    assert "MOL:GTP_activation" not in nodes and "PLK1_dna_1" in nodes and len(pruned[0]) < len(graphs[0])
    deg = {}
    for u, v, _w in pruned[0]:
        deg[u] = deg.get(u, 0) + 1
        deg[v] = deg.get(v, 0) + 1
    assert all(d >= 2 or n in set(feature_ids) for n, d in deg.items())                      # only data features may be leaves


def test_prep_pathway_featuresets(tmp_path):
    """runtests.jl:578-615."""
    path = _sif_file(tmp_path)
    nodeset = {"MOL:GTP", "GRASP65/GM130/RAB1/GTP/PLK1", "METAPHASE", "PLK1", "PAK1", "RAB1A", "SGOL1", "PP2A-ALPHA B56"}
    assert PP.sif_to_nodeset(SIF) == nodeset and PP.sifs_to_nodesets([SIF]) == [nodeset]
    ids = [f"{g}_{i}" for i, g in enumerate(GENES, start=1)]
    fs, new_ids, kept = PP.prep_pathway_featuresets([path], GENES)
    assert new_ids == ids and kept == [1]
    assert fs[0] == {"PLK1_1", "PLK1_2", "PLK1_3", "PAK1_4", "PAK1_5", "PAK1_6", "SGOL1_7"}
    fs, new_ids, _ = PP.prep_pathway_featuresets([path], GENES, feature_ids=ids)
    assert new_ids == ids
    by_view, new_ids, kept_by_view = PP.prep_pathway_featuresets([path], GENES, ASSAYS)
    assert by_view == {"cna": [{"PLK1_1"}], "mutation": [{"PLK1_2", "PAK1_6"}], "mrnaseq": [{"PLK1_3", "SGOL1_7"}],
                       "rppa": [{"PAK1_4"}], "methylation": [{"PAK1_5"}]}
    assert new_ids == ids and all(v == [1] for v in kept_by_view.values())
    # a pathway without any gene of the data is dropped, and its id with it
    other = [["AAA", "a>", "BBB"]]
    fs, _, kept = PP.prep_pathway_featuresets([other, SIF], GENES, featureset_ids=["none", "plk"])
    assert kept == ["plk"] and len(fs) == 1


def test_network_regulariser_on_the_prepared_pathway(tmp_path):
    """runtests.jl:669-688 for the mirror's and the oracle's NetworkRegularizer, then BASELINE config 1 in small: the
    constructor takes the prepared graphs as `feature_graphs` (K = number of pathways, src/model.jl:122)."""
    graphs, model_features = PP.prep_pathway_graphs([_sif_file(tmp_path)], GENES, [DOGMA[a] for a in ASSAYS])
    pwy_nodes = PP.get_all_nodes(graphs[0])
    n_obs, n_unobs = len(model_features), len(pwy_nodes - set(model_features))
    assert n_unobs > 0 and set(model_features) & pwy_nodes == set(model_features[:7])      # BRCA is not in the pathway
    for reg in (P.NetworkRegularizer(model_features, graphs), O.NetworkRegularizer(model_features, graphs)):
        assert len(reg.AA) == 1
        assert reg.AA[0].shape == (n_obs, n_obs) and reg.AB[0].shape == (n_obs, n_unobs) and reg.BB[0].shape == (n_unobs, n_unobs)
    a, b = P.NetworkRegularizer(model_features, graphs), O.NetworkRegularizer(model_features, graphs)
    for x, y in ((a.AA[0], b.AA[0]), (a.AB[0], b.AB[0]), (a.BB[0], b.BB[0])):
        assert np.allclose(np.asarray(x.todense()), np.asarray(y.todense()))
    rng = np.random.default_rng(0)
    D = rng.standard_normal((20, n_obs)).astype(np.float32)
    model = P.PathMatFacModel(D, feature_ids=model_features, feature_views=ASSAYS, feature_graphs=graphs, lambda_Y_graph=1.0)
    assert model.matfac.X.shape == (1, 20) and model.matfac.Y.shape == (1, n_obs)
    regs = getattr(model.matfac.Y_reg, "regularizers", [model.matfac.Y_reg])
    assert any(isinstance(r, P.NetworkRegularizer) for r in regs)
