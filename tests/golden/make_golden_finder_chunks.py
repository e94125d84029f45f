"""Generate tests/golden/finder_chunks.npz: the UNMODIFIED reference DenseKmerFinder run chunk by chunk over the critical paths,
the way `graph_kmer_index index -t T` does (command_line_interface.py:588-608: start/stop_at_critical_path_number per chunk,
results concatenated).  Build container only: python tests/golden/make_golden_finder_chunks.py"""
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


def chunk_list(n_paths, n_chunks):
    """command_line_interface.py:588-603."""
    n_chunks = min(n_chunks, n_paths)
    per = n_paths // n_chunks
    starts = list(range(0, n_paths, per))
    ends = starts[1:] + [n_paths]
    return list(zip(starts, ends))


def main():
    rng = np.random.default_rng(17)
    out = {}
    kept = 0
    for i in range(60):
        nv = int(rng.integers(3, 14))
        spacing = int(rng.choice([2, 3, 5, 8, 20, 60]))
        k = int(rng.choice([3, 4, 5, 7, 11, 16, 31]))
        seqs, edges, linear, af = synthetic.variant_graph(nv, spacing=spacing, seed=500 + i, p_deletion=0.3, p_insertion=0.15 if i % 3 == 0 else 0.0,
                                                          p_nested=0.3 if i % 2 else 0.0, tail=int(rng.integers(1, 50)))
        kwargs = {"max_variant_nodes": int(rng.choice([1, 2, 4, 5]))}
        graph = Graph.from_dicts(seqs, edges, linear, af)
        try:
            crit = CriticalGraphPaths.from_graph(graph, k)
            if len(crit) < 3:
                continue
            n_chunks = int(rng.choice([2, 3, 5]))
            chunks = chunk_list(len(crit), n_chunks)
            results = []
            for s, e in chunks:
                finder = DenseKmerFinder(graph, k=k, critical_graph_paths=crit, start_at_critical_path_number=s, stop_at_critical_path_number=e, **kwargs)
                finder.find()
                results.append(dict(kmers=finder._kmers.get_nparray().copy(), nodes=finder._nodes.get_nparray().copy(),
                                    start_nodes=finder._start_nodes.get_nparray().copy(), start_offsets=finder._start_offsets.get_nparray().copy(),
                                    allele_frequencies=finder._allele_frequencies.get_nparray().copy()))
            # whitelist (kf:95-104, 130-132, 362-365): only these k-mers are stored -- half of what an unrestricted run finds
            full = DenseKmerFinder(graph, k=k, critical_graph_paths=crit, **kwargs)
            full.find()
            found = np.unique(full._kmers.get_nparray())
            whitelist = found[rng.random(len(found)) < 0.5]
            wl = DenseKmerFinder(graph, k=k, critical_graph_paths=crit, whitelist=set(int(x) for x in whitelist), **kwargs)
            wl.find()
            wl_result = dict(kmers=wl._kmers.get_nparray().copy(), nodes=wl._nodes.get_nparray().copy(),
                             start_nodes=wl._start_nodes.get_nparray().copy(), start_offsets=wl._start_offsets.get_nparray().copy(),
                             allele_frequencies=wl._allele_frequencies.get_nparray().copy())
        except Exception as e:                            # graphs the reference itself cannot process are skipped
            print("skip:", type(e).__name__, e)
            continue
        p = "c%d_" % kept
        out[p + "whitelist"] = whitelist.astype(np.int64)
        for key, v in wl_result.items():
            out[p + "wl_" + key] = v
        for key, v in graph.to_arrays().items():
            out[p + "g_" + key] = v
        out[p + "k"] = np.int64(k)
        out[p + "max_variant_nodes"] = np.int64(kwargs["max_variant_nodes"])
        out[p + "crit_nodes"], out[p + "crit_offsets"] = crit.nodes, crit.offsets
        out[p + "chunks"] = np.array(chunks, dtype=np.int64)
        for j, r in enumerate(results):
            for key, v in r.items():
                out[p + "r%d_%s" % (j, key)] = v
        kept += 1
    out["n_cases"] = np.int64(kept)
    np.savez_compressed(os.path.join(HERE, "finder_chunks.npz"), **out)
    print("cases:", kept, os.path.getsize(os.path.join(HERE, "finder_chunks.npz")))


if __name__ == "__main__":
    main()
