"""Generate tests/golden/finder_c5.json: BASELINE config 5 (DenseKmerFinder k=31 on the synthetic SNP/indel graph with 100 k variants,
max_variant_nodes=5) run by the UNMODIFIED reference, chunk by chunk over the critical paths and over worker processes the way
`graph_kmer_index index -t T` does it (command_line_interface.py:588-608).  The 46.6 M rows are not stored: every chunk contributes
its row count and the SHA-256 of its ordered rows (kmers int64 | nodes int32 | start_nodes int32 | start_offsets int16 |
allele_frequencies float64, each array's bytes in order).  tests/test_gpu_finder.py::test_config5_chunk_digests compares the device
finder's chunks with them.  Build container only (needs /root/reference):  python tests/golden/make_golden_c5.py [n_variants] [n_chunks] [procs]"""
import hashlib
import json
import logging
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

K, MAX_VARIANT_NODES, SEED, SPACING, P_DELETION = 31, 5, 7, 300, 0.2
_state = {}


def chunk_list(n_paths, n_chunks):
    """command_line_interface.py:588-603."""
    n_chunks = min(n_chunks, n_paths)
    per = n_paths // n_chunks
    starts = list(range(0, n_paths, per))
    ends = starts[1:] + [n_paths]
    return list(zip(starts, ends))


def digest(rows):
    h = hashlib.sha256()
    for key, dtype in (("kmers", np.int64), ("nodes", np.int32), ("start_nodes", np.int32), ("start_offsets", np.int16), ("allele_frequencies", np.float64)):
        a = np.ascontiguousarray(rows[key])
        assert a.dtype == dtype, (key, a.dtype)
        h.update(a.tobytes())
    return h.hexdigest()


def _init(n_variants):
    from oracle import ref_shims
    ref_shims.install()
    logging.disable(logging.CRITICAL)
    from graph_kmer_index.critical_graph_paths import CriticalGraphPaths
    from graph_kmer_index_b200 import synthetic
    from oracle.obgraph_standin import Graph
    seqs, edges, linear, af = synthetic.variant_graph(n_variants, spacing=SPACING, seed=SEED, p_deletion=P_DELETION)
    graph = Graph.from_dicts(seqs, edges, linear, af)
    _state["graph"] = graph
    _state["crit"] = CriticalGraphPaths.from_graph(graph, K)


def _run(chunk):
    from graph_kmer_index.kmer_finder import DenseKmerFinder
    s, e = chunk
    finder = DenseKmerFinder(_state["graph"], k=K, critical_graph_paths=_state["crit"], max_variant_nodes=MAX_VARIANT_NODES,
                             start_at_critical_path_number=s, stop_at_critical_path_number=e)
    finder.find()
    rows = dict(kmers=finder._kmers.get_nparray(), nodes=finder._nodes.get_nparray(), start_nodes=finder._start_nodes.get_nparray(),
                start_offsets=finder._start_offsets.get_nparray(), allele_frequencies=finder._allele_frequencies.get_nparray())
    return int(len(rows["kmers"])), digest(rows)


def main():
    n_variants = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) - 1)
    t0 = time.time()
    _init(n_variants)                      # built before the fork: the workers share the graph and the critical paths
    n_paths = len(_state["crit"])
    chunks = chunk_list(n_paths, n_chunks)
    with mp.get_context("fork").Pool(procs) as pool:
        results = pool.map(_run, chunks, chunksize=1)
    out = dict(k=K, max_variant_nodes=MAX_VARIANT_NODES, n_variants=n_variants, spacing=SPACING, seed=SEED, p_deletion=P_DELETION,
               n_critical_paths=n_paths, chunks=[list(c) for c in chunks], rows=[r[0] for r in results], sha256=[r[1] for r in results],
               total_rows=int(sum(r[0] for r in results)),
               generator="tests/golden/make_golden_c5.py: unmodified reference DenseKmerFinder (oracle/ref_shims.py, oracle/obgraph_standin.py), %d worker processes" % procs,
               row_layout="sha256(kmers int64 | nodes int32 | start_nodes int32 | start_offsets int16 | allele_frequencies float64), arrays in this order, rows in the finder's order")
    name = "finder_c5.json" if n_variants == 100_000 else "finder_c5_%d.json" % n_variants
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %s: %d chunks, %d rows, %.0f s" % (name, len(chunks), out["total_rows"], time.time() - t0))


if __name__ == "__main__":
    main()
