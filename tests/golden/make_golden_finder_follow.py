"""Generate tests/golden/finder_follow.npz: the UNMODIFIED reference DenseKmerFinder run with `only_follow_nodes`
(kmer_finder.py:385-388) -- the way unique_variant_kmers.py:90-96 uses it (one variant node, only_store_nodes the same node,
find_only_kmers_starting_at_position a few bases ahead of it) and through find() with sets of several nodes, which also
covers nodes with more than one successor in the set (followed in the set's own iteration order) and paths that pass more
variant nodes than max_variant_nodes.  Build container only: python tests/golden/make_golden_finder_follow.py"""
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
    rng = np.random.default_rng(11)
    cases = []
    for i in range(36):
        nv = int(rng.integers(3, 40))
        spacing = int(rng.choice([2, 3, 5, 8, 20]))
        k = int(rng.choice([3, 4, 5, 7, 11, 16, 31]))
        seqs, edges, linear, af = synthetic.variant_graph(nv, spacing=spacing, seed=500 + i, p_deletion=0.3, p_insertion=0.15 if i % 3 == 0 else 0.0,
                                                          p_nested=0.3 if i % 2 else 0.0, tail=int(rng.integers(1, 50)))
        graph = Graph.from_dicts(seqs, edges, linear, af)
        a = graph.to_arrays()
        n = len(a["seq_offsets"]) - 1
        variant_nodes = np.flatnonzero(a["is_linear"] == 0)
        sources = np.searchsorted(a["edge_offsets"], np.arange(len(a["edges"])), side="right") - 1
        if i % 2 == 0:
            # unique_variant_kmers.py:90-96: one node, searched from a position ahead of it
            node = int(rng.choice(variant_nodes)) if i % 4 == 0 else int(rng.integers(1, n))
            preds = sources[a["edges"] == node]
            if len(preds) == 0:
                continue
            pred = int(preds[0])
            size = int(a["seq_offsets"][pred + 1] - a["seq_offsets"][pred])
            pos = (pred, max(0, size - int(rng.integers(1, 9))))
            kwargs = {"max_variant_nodes": int(rng.choice([0, 1, 2, 4])), "only_store_nodes": {node}, "only_follow_nodes": {node}}
        else:
            pos = None
            follow = set(int(x) for x in rng.choice(np.arange(1, n), size=min(n - 1, int(rng.integers(1, 9))), replace=False))
            if i % 3 == 1:                               # both alleles of some bubble: two successors of one node in the set
                multi = [u for u in range(n) if a["edge_offsets"][u + 1] - a["edge_offsets"][u] >= 2]
                u = int(rng.choice(multi))
                follow.update(int(x) for x in a["edges"][a["edge_offsets"][u]:a["edge_offsets"][u + 1]])
            kwargs = {"max_variant_nodes": int(rng.choice([0, 1, 2, 4])), "only_follow_nodes": follow}
            if i % 5 == 0:
                kwargs["only_save_one_node_per_kmer"] = True
        cases.append((graph, k, kwargs, pos))
    # a bubble whose alleles are nodes 7 and 8: {7, 8}.intersection([7, 8]) iterates 8 before 7
    seqs = {0: "ACGTAC", 1: "G", 2: "T", 3: "CCA", 4: "A", 5: "", 6: "GATTACA", 7: "C", 8: "G", 9: "TTGACC"}
    edges = {0: [1, 2], 1: [3], 2: [3], 3: [4, 5], 4: [6], 5: [6], 6: [7, 8], 7: [9], 8: [9]}
    cases.append((Graph.from_dicts(seqs, edges, [0, 1, 3, 4, 6, 7, 9]), 4, {"max_variant_nodes": 1, "only_follow_nodes": {7, 8, 2}}, None))
    cases.append((Graph.from_dicts(seqs, edges, [0, 1, 3, 4, 6, 7, 9]), 3, {"max_variant_nodes": 0, "only_follow_nodes": {8, 5}, "only_store_nodes": {8}}, (6, 2)))

    out = {}
    kept = reordered = 0
    for graph, k, kwargs, pos in cases:
        try:
            crit = CriticalGraphPaths.from_graph(graph, k)
            ref = run_reference(graph, k, kwargs, pos)
        except Exception as e:                            # graphs the reference itself cannot process are skipped
            print("skip:", type(e).__name__, e)
            continue
        a = graph.to_arrays()
        for u in range(len(a["seq_offsets"]) - 1):
            succ = [int(x) for x in a["edges"][a["edge_offsets"][u]:a["edge_offsets"][u + 1]]]
            inter = list(kwargs["only_follow_nodes"].intersection(succ))
            reordered += inter != [x for x in succ if x in kwargs["only_follow_nodes"]]
        p = "c%d_" % kept
        for key, v in a.items():
            out[p + "g_" + key] = v
        out[p + "k"] = np.int64(k)
        out[p + "max_variant_nodes"] = np.int64(kwargs.get("max_variant_nodes", 4))
        out[p + "one_node"] = np.int64(kwargs.get("only_save_one_node_per_kmer", False))
        out[p + "only_store_nodes"] = np.array(sorted(kwargs.get("only_store_nodes", [])), dtype=np.int64)
        out[p + "only_follow_nodes"] = np.array(sorted(kwargs["only_follow_nodes"]), dtype=np.int64)
        out[p + "only_position"] = np.array(pos if pos is not None else [], dtype=np.int64)
        out[p + "crit_nodes"], out[p + "crit_offsets"] = crit.nodes, crit.offsets
        for key, v in ref.items():
            out[p + "ref_" + key] = v
        kept += 1
    out["n_cases"] = np.int64(kept)
    np.savez_compressed(os.path.join(HERE, "finder_follow.npz"), **out)
    print("cases:", kept, "rows:", sum(len(out["c%d_ref_kmers" % i]) for i in range(kept)), "nodes followed in another order than the edges':",
          reordered, os.path.getsize(os.path.join(HERE, "finder_follow.npz")))


if __name__ == "__main__":
    main()
