"""Sweep the K3 layout knobs on the c2 workload: Bloom filter budget (GKI_FILTER_MAX_MB), bits per key
(GKI_FILTER_K), raw vs canonical keys (GKI_TABLE_RAW).  Prints kernel ms per launch; checks that the node counts are
identical for every setting.  Usage: python profiles/tune_count.py [entries] [reads]
The kernel-variant knobs (GKI_HINTS, GKI_RPW, GKI_COUNT_CTAS) only exist in a library built with GKI_BUILD_EXPERIMENT_KNOBS=1
(graph_kmer_index_b200/_lib.py: -DGKI_EXPERIMENT_KNOBS, csrc/common.cuh); the shipped build ignores them."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import DeviceIndex, _lib, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
modulo, k, L, n_nodes, p_hit = 452_930_477, 31, 150, max(n // 10, 1), 100
dev = torch.device("cuda")
glen = synthetic.genome_length(n, k)
genome = torch.empty(glen, dtype=torch.uint8, device=dev)
_lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
hashes = torch.empty(n, dtype=torch.int64, device=dev)
nodes = torch.empty(n, dtype=torch.int32, device=dev)
_lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(hashes), _lib.ptr(nodes), None, None, None)
h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
s_k, s_n = torch.empty_like(hashes), torch.empty_like(nodes)
_lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, modulo, 1, _lib.ptr(h2i), _lib.ptr(nkm),
          _lib.ptr(s_k), _lib.ptr(s_n), None, None, None, None, None)
reads = torch.empty((R, L), dtype=torch.uint8, device=dev)
_lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, R, L, p_hit, 0, _lib.ptr(reads), None)
torch.cuda.synchronize()
counts = torch.zeros(n_nodes, dtype=torch.float64, device=dev)
ref_sum = None
results = []
KNOBS = ("GKI_COUNT_MINB", "GKI_COUNT_GRID_MULT", "GKI_COUNT_CTAS", "GKI_FILTER_MZ", "GKI_FILTER_MAX_MB", "GKI_FILTER_K", "GKI_TABLE_RAW", "GKI_RPW", "GKI_HINTS", "GKI_L2_FETCH")
combos = json.loads(sys.argv[3]) if len(sys.argv) > 3 else [{"GKI_FILTER_MAX_MB": mb} for mb in (0, 16, 32, 48, 64)]
for env in combos:
    for key in KNOBS:
        os.environ.pop(key, None)
    for key, v in env.items():
        os.environ[key] = str(v)
    index = DeviceIndex(h2i, nkm, s_k, s_n, modulo)
    index.prepare_counting(k)
    for _ in range(2):
        index.reset_counts()
        index.count_reads(reads, k, True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        index.reset_counts()
        a.record()
        index.count_reads(reads, k, True)
        b.record()
    index.node_counts(n_nodes, out=counts)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    total = float(counts.sum().item())
    ref_sum = total if ref_sum is None and not int(env.get("GKI_HINTS", 0)) & 16 else ref_sum
    assert total == ref_sum or int(env.get("GKI_HINTS", 0)) & 16, (env, total, ref_sum)   # bit 4 drops the survivors
    info = index.info()
    r = dict(env=env, kernel_ms=ms, gkmers_per_s=R * 240 / ms / 1e6, has_filter=info["has_filter"], device_bytes=info["device_bytes"])
    results.append(r)
    print(json.dumps(r), flush=True)
    index.close()
