"""Generate tests/golden/finder_cases.npz by running the UNMODIFIED reference DenseKmerFinder / CriticalGraphPaths
(kmer_finder.py, critical_graph_paths.py) on the reference's own test graphs and on random variant graphs, with
oracle/obgraph_standin.Graph standing in for obgraph.  Build container only: python tests/golden/make_golden_finder.py"""
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
from oracle import ref_shims  # noqa: E402

ref_shims.install()
logging.disable(logging.CRITICAL)
from graph_kmer_index.critical_graph_paths import CriticalGraphPaths  # noqa: E402
from graph_kmer_index.kmer_finder import DenseKmerFinder  # noqa: E402
from graph_kmer_index_b200 import synthetic  # noqa: E402
from oracle.obgraph_standin import Graph  # noqa: E402

# (node sequences, edges, linear ref, k, kwargs, only_position) -- graphs of the reference's tests/test_kmer_finder.py
REF_TESTS = [
    ({0: "AAA", 1: "C", 2: "T", 3: "AAA"}, {0: [1, 2], 2: [3], 1: [3]}, [0, 1, 3], 3, {}, None),
    ({0: "ACTGACTG", 1: "A", 2: "T", 3: "AAAAA", 4: "C", 5: "T", 6: "TGGGGG"}, {0: [1, 2], 2: [3], 1: [3], 3: [4, 5], 4: [6], 5: [6]}, [0, 1, 3, 4, 6], 3, {}, None),
    ({0: "AAA", 1: "C", 2: "T", 3: "AAAA", 4: "C", 5: "G", 6: "AAA", 7: "TTT"}, {0: [1, 2, 7], 1: [3], 2: [3], 3: [4, 5], 4: [6], 5: [6], 7: [6]}, [0, 1, 3, 4, 6], 3, {}, None),
    ({1: "ATC", 2: "AAAAAAAA", 3: "T", 4: "CTA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4], 3, {}, None),
    ({1: "ACT", 2: "C", 3: "", 4: "ACT"}, {1: [2, 3], 3: [4], 2: [4]}, [1, 2, 4], 3, {}, None),
    ({1: "AAAAA", 2: "", 3: "CCCCCC"}, {1: [2], 2: [3]}, [1, 3], 3, {}, None),
    ({1: "CC", 2: "", 3: "CCTCTG"}, {1: [2], 2: [3]}, [1, 3], 4, {}, (1, 0)),
    ({1: "AAAAA", 2: "G", 3: "", 4: "CCCCCC"}, {1: [2], 2: [3], 3: [4]}, [1, 2, 4], 3, {}, None),
    ({1: "CCCCC", 2: "G", 3: "", 4: "ACT", 5: "", 6: "GC", 7: "A", 8: "T", 9: "G", 10: "GGG"},
     {1: [2, 3], 2: [4], 3: [4], 4: [5, 6], 5: [7], 6: [7], 7: [8, 9], 8: [10], 9: [10]}, [1, 2, 4, 7, 8, 10], 3, {}, None),
    ({1: "CCCCCCCCCC", 2: "AAAA"}, {1: [2]}, [1, 2], 3, {}, None),
    ({1: "CATGCATGCCTG", 2: "CCAAG"}, {1: [2]}, [1, 2], 5, {}, None),
    ({1: "ACT", 2: "", 3: "GGG", 4: "", 5: "A", 6: "CCC"}, {1: [2, 3], 2: [4, 5], 3: [4, 5], 4: [6], 5: [6]}, [1, 5, 6], 3, {}, None),
    ({1: "ACT", 2: "", 3: "GGG", 4: "", 5: "A", 6: "CCC"}, {1: [2, 3], 2: [4, 5], 3: [4, 5], 4: [6], 5: [6]}, [1, 5, 6], 3, {"max_variant_nodes": 0}, None),
    ({1: "ACT", 2: "", 3: "GGG", 4: "", 5: "A", 6: "CCC"}, {1: [2, 3], 2: [4, 5], 3: [4, 5], 4: [6], 5: [6]}, [1, 5, 6], 3, {"max_variant_nodes": 1}, None),
    ({1: "ACTACTACTACT", 2: "G", 3: "C", 4: "GCAGCA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4], 3, {}, None),
    ({1: "G" * 100, 2: "C", 3: "T", 4: "G" * 10}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4], 31, {}, None),
    ({1: "ACTACT", 2: "G", 3: "C", 4: "GCAGCA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4], 3, {"only_store_nodes": {2, 3}}, (1, 4)),
    ({1: "ACTACT", 2: "G", 3: "C", 4: "GCAGCA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4], 5, {"only_store_nodes": {2, 3}}, (1, 5)),
    ({1: "taacccctaacccctaaccctaaccctaac", 2: "", 3: "G", 4: "ccctaaccctaaccctaacccctaacccta"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 4], 31,
     {"only_store_nodes": {2, 3}}, (1, 22)),
    ({1: "ACTGA", 2: "", 3: "C", 4: "GGGGGGGGG"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 4], 9, {"only_store_nodes": {2, 3}}, (1, 2)),
    ({1: "ACTGAACTG", 2: "A", 3: "C", 4: "GGGG", 5: "", 6: "T", 7: "CCCCCC"}, {1: [3, 2], 2: [4], 3: [4], 4: [5, 6], 5: [7], 6: [7]},
     [1, 2, 4, 6, 7], 13, {"only_store_nodes": {5, 6}, "max_variant_nodes": 5}, (1, 6)),
    ({1: "AAAAAACTG", 2: "A", 3: "G", 4: "GC", 5: "T", 6: "C", 7: "TGAGCCCCC", 8: "A", 9: "T", 10: "AAAAA"},
     {1: [2, 3], 2: [4], 3: [4], 4: [5, 6], 5: [7], 6: [7], 7: [8, 9], 9: [10], 8: [10]}, [1, 2, 4, 5, 7, 8, 10], 5, {}, None),
    ({0: "AGTAGA", 1: "G", 2: "CT", 3: "A", 4: "CTA", 5: "G", 6: "A", 7: "TCATA"}, {0: [1, 2], 1: [3], 2: [3], 3: [4], 4: [5, 6], 5: [7], 6: [7], 7: []},
     [0, 1, 3, 4, 5, 7], 3, {}, None),
    ({0: "AGTAGA", 1: "G", 2: "CT", 3: "ACTA", 5: "G", 6: "A", 7: "TCATA"}, {0: [1, 2], 1: [3], 2: [3], 3: [5, 6], 5: [7], 6: [7], 7: []},
     [0, 1, 3, 5, 7], 3, {}, None),                       # test_case1: the ordered 38-row golden
    ({0: "ACTGACTG", 1: "A", 2: "T", 3: "AAAAA", 4: "C", 5: "T", 6: "TGGGGG", 100: ""}, {0: [1, 2, 100], 2: [3], 1: [3], 3: [4, 5], 4: [6], 5: [6], 100: [6]},
     [0, 1, 3, 4, 6], 3, {}, None),                       # tests/test_critical_graph_paths.py test5
]


def run_reference(graph, k, kwargs, only_position):
    finder = DenseKmerFinder(graph, k=k, **kwargs)
    if only_position is None:
        finder.find()
    else:
        finder.find_only_kmers_starting_at_position(*only_position)
    return dict(kmers=finder._kmers.get_nparray().copy(), nodes=finder._nodes.get_nparray().copy(),
                start_nodes=finder._start_nodes.get_nparray().copy(), start_offsets=finder._start_offsets.get_nparray().copy(),
                allele_frequencies=finder._allele_frequencies.get_nparray().copy())


def main():
    cases = []
    for seqs, edges, linear, k, kwargs, pos in REF_TESTS:
        cases.append((Graph.from_dicts(seqs, edges, linear), k, kwargs, pos))
    rng = np.random.default_rng(5)
    for i in range(40):
        nv = int(rng.integers(1, 12))
        spacing = int(rng.choice([2, 3, 5, 8, 20, 60]))
        k = int(rng.choice([3, 4, 5, 7, 11, 16, 31]))
        seqs, edges, linear, af = synthetic.variant_graph(nv, spacing=spacing, seed=100 + i, p_deletion=0.3, p_insertion=0.15 if i % 3 == 0 else 0.0,
                                                          p_nested=0.3 if i % 2 else 0.0, tail=int(rng.integers(1, 50)))
        kwargs = {"max_variant_nodes": int(rng.choice([1, 2, 4, 5]))}
        if i % 4 == 0:
            kwargs["only_save_one_node_per_kmer"] = True
        cases.append((Graph.from_dicts(seqs, edges, linear, af), k, kwargs, None))
    out = {"n_cases": np.int64(len(cases))}
    kept = 0
    for graph, k, kwargs, pos in cases:
        try:
            crit = CriticalGraphPaths.from_graph(graph, k)
            ref = run_reference(graph, k, kwargs, pos)
        except Exception as e:                            # graphs the reference itself cannot process are skipped
            print("skip:", type(e).__name__, e)
            continue
        p = "c%d_" % kept
        for key, v in graph.to_arrays().items():
            out[p + "g_" + key] = v
        out[p + "k"] = np.int64(k)
        out[p + "max_variant_nodes"] = np.int64(kwargs.get("max_variant_nodes", 4))
        out[p + "one_node"] = np.int64(kwargs.get("only_save_one_node_per_kmer", False))
        out[p + "only_store_nodes"] = np.array(sorted(kwargs.get("only_store_nodes", [])), dtype=np.int64)
        out[p + "only_position"] = np.array(pos if pos is not None else [], dtype=np.int64)
        out[p + "crit_nodes"], out[p + "crit_offsets"] = crit.nodes, crit.offsets
        for key, v in ref.items():
            out[p + "ref_" + key] = v
        kept += 1
    out["n_cases"] = np.int64(kept)
    np.savez_compressed(os.path.join(HERE, "finder_cases.npz"), **out)
    print("cases:", kept, "rows:", sum(len(out["c%d_ref_kmers" % i]) for i in range(kept)), os.path.getsize(os.path.join(HERE, "finder_cases.npz")))


if __name__ == "__main__":
    main()
