"""FASTQ file -> node counts: a synthetic FASTQ of R 150 bp reads is indexed (fastx_open: host threads) and counted
(gki_count_fastx: packing lanes straight from the mapping + the PACKED count kernel) against the C2 index.
Usage: python profiles/bench_fastq.py [entries] [reads]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import DeviceIndex, _lib, synthetic  # noqa: E402
from graph_kmer_index_b200.read_kmers import FastxFile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
modulo, k, L, n_nodes = 452_930_477, 31, 150, max(n // 10, 1)
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
_lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, R, L, 100, 0, _lib.ptr(reads), None)
index = DeviceIndex(h2i, nkm, s_k, s_n, modulo)
index.prepare_counting(k)
index.count_reads(reads, k)
want = index.node_counts(n_nodes).sum()

rec = np.empty((R, 3 + L + 3 + L + 1), dtype=np.uint8)
rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
rec[:, 3:3 + L] = reads.cpu().numpy()
rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
rec[:, 6 + L:6 + 2 * L] = ord("I")
rec[:, -1] = ord("\n")
path = os.path.join(tempfile.gettempdir(), "gki_bench.fastq")
rec.tofile(path)
file_gb = rec.nbytes / 1e9
del rec
try:
    for _ in range(4):
        t = time.perf_counter()
        fx = FastxFile(path)
        t_open = time.perf_counter() - t
        index.reset_counts()
        torch.cuda.synchronize()
        t = time.perf_counter()
        n_kmers = index.count_fastx(fx, k)
        torch.cuda.synchronize()
        t_count = time.perf_counter() - t
        fx.close()
        assert index.node_counts(n_nodes).sum() == want and n_kmers == R * 240
        print(json.dumps(dict(stage="count_fastq_file", file_gb=file_gb, reads=R, open_ms=t_open * 1e3, count_ms=t_count * 1e3,
                              gkmers_per_s_count_only=n_kmers / t_count / 1e9, gkmers_per_s_open_and_count=n_kmers / (t_open + t_count) / 1e9,
                              file_gbs_open=file_gb / t_open, cpus=os.cpu_count())), flush=True)
finally:
    os.remove(path)
