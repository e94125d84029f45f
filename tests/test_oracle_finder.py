"""Pins oracle/finder_oracle.py (CriticalGraphPaths + DenseKmerFinder restatement) to the unmodified reference: ordered,
element-wise equality on every fixture of tests/golden/finder_cases.npz (the reference's own test graphs, incl. the
38-row ordered golden of tests/test_kmer_finder.py::test_case1, plus random SNP / indel / nested graphs)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import finder_oracle, numpy_oracle as no

G_KEYS = ("seq_offsets", "seq", "edge_offsets", "edges", "is_linear", "allele_frequencies", "n_in_edges", "first_node",
          "node_to_ref_offset", "chromosome_start_nodes")


def load_case(g, i):
    p = "c%d_" % i
    arrays = {k: g[p + "g_" + k] for k in G_KEYS}
    pos = g[p + "only_position"]
    store = g[p + "only_store_nodes"]
    opts = dict(k=int(g[p + "k"]), max_variant_nodes=int(g[p + "max_variant_nodes"]), only_save_one_node_per_kmer=bool(g[p + "one_node"]),
                only_store_nodes=set(int(x) for x in store) if len(store) else None, only_position=tuple(int(x) for x in pos) if len(pos) else None)
    if p + "only_follow_nodes" in g:
        opts["only_follow_nodes"] = set(int(x) for x in g[p + "only_follow_nodes"])
    ref = {k: g[p + "ref_" + k] for k in ("kmers", "nodes", "start_nodes", "start_offsets", "allele_frequencies")}
    return arrays, opts, ref, (g[p + "crit_nodes"], g[p + "crit_offsets"])


def test_finder_oracle_matches_reference():
    g = load_golden("finder_cases")
    n = int(g["n_cases"])
    assert n >= 40
    total = 0
    for i in range(n):
        arrays, opts, ref, crit = load_case(g, i)
        cn, co = finder_oracle.critical_paths(arrays, opts["k"])
        assert np.array_equal(cn, crit[0]) and np.array_equal(co, crit[1]), i
        got = finder_oracle.dense_kmer_finder(arrays, **opts)
        for key in ref:
            assert np.array_equal(got[key], ref[key]), (i, key, got[key][:10], ref[key][:10])
        total += len(ref["kmers"])
    assert total > 10000


def test_finder_oracle_only_follow_nodes_matches_reference():
    """kf:385-388 -- fixtures of tests/golden/make_golden_finder_follow.py (the unmodified reference, used the way
    unique_variant_kmers.py:90-96 does and through find())."""
    g = load_golden("finder_follow")
    n = int(g["n_cases"])
    assert n >= 20
    differs = 0
    for i in range(n):
        arrays, opts, ref, _ = load_case(g, i)
        got = finder_oracle.dense_kmer_finder(arrays, **opts)
        for key in ref:
            assert np.array_equal(got[key], ref[key]), (i, key, got[key][:10], ref[key][:10])
        plain = finder_oracle.dense_kmer_finder(arrays, **dict(opts, only_follow_nodes=None)) if opts["only_position"] is None else None
        differs += plain is not None and not np.array_equal(plain["kmers"], ref["kmers"])
    assert differs >= 5          # the option changes the walk in the fixtures


def test_reference_readme_example():
    """Readme.md:69-70 of the reference: k=5 finder output on its 4-node example."""
    from oracle.obgraph_standin import Graph
    graph = Graph.from_dicts({1: "ACTGA", 2: "A", 3: "G", 4: "AAAAA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4])
    got = finder_oracle.dense_kmer_finder(graph.to_arrays(), k=5)
    want = {no.sequence_to_kmer_hash(s) for s in ("ACTGA", "CTGAA", "TGAAA", "GAAAA", "AAAAA", "CTGAG", "TGAGA", "GAGAA", "AGAAA")}
    assert set(int(x) for x in got["kmers"]) == want
