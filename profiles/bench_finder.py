"""BASELINE config 5: DenseKmerFinder k=31 on a synthetic SNP/indel graph (default 100k variants, one every ~300 bp,
80% SNPs / 20% deletions, max_variant_nodes=5).  Times the device finder (critical paths + prepare + fill, host arrays in,
host rows out) and the oracle port (pure Python restatement of the reference's loop) on a prefix graph; checks ordered
equality on that prefix.  Usage: python profiles/bench_finder.py [n_variants] [n_variants_cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_kmer_index_b200 as gki  # noqa: E402
from graph_kmer_index_b200 import synthetic  # noqa: E402
from oracle import finder_oracle  # noqa: E402
from oracle.obgraph_standin import Graph  # noqa: E402

n_variants = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000
k = 31


def make(nv):
    seqs, edges, linear, af = synthetic.variant_graph(nv, spacing=300, seed=7, p_deletion=0.2)
    return Graph.from_dicts(seqs, edges, linear, af).to_arrays()


def run(arrays):
    t0 = time.perf_counter()
    finder = gki.DenseKmerFinder(arrays, k, max_variant_nodes=5)
    finder.find()
    return finder, time.perf_counter() - t0


small = make(n_cpu)
run(small)                                   # warm-up (library load, context)
finder, _ = run(small)
t0 = time.perf_counter()
want = finder_oracle.dense_kmer_finder(small, k, max_variant_nodes=5)
cpu_s = time.perf_counter() - t0
for key in ("kmers", "nodes", "start_nodes", "start_offsets", "allele_frequencies"):
    assert np.array_equal(finder._results[key], want[key]), key
big = make(n_variants)
finder, gpu_s = run(big)
finder, gpu_s = run(big)
rows = len(finder._results["kmers"])
print(json.dumps(dict(stage="DenseKmerFinder.find k=31 max_variant_nodes=5", n_variants=n_variants, graph_bp=int(big["seq_offsets"][-1]),
                      n_nodes=len(big["seq_offsets"]) - 1, rows=rows, gpu_seconds=gpu_s, gpu_rows_per_s=rows / gpu_s,
                      cpu_port_rows_per_s=len(want["kmers"]) / cpu_s, cpu_sample="oracle/finder_oracle.py on %d variants (%d rows), 1 thread" % (n_cpu, len(want["kmers"])),
                      parity="ordered equality on the %d-variant prefix graph" % n_cpu)))
