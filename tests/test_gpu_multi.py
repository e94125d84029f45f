"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): reads shard over ranks, the index is replicated, one NCCL
all-reduce combines the node counts; the result must equal the oracle's count over all reads bit for bit."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import graph_kmer_index_b200 as gki
    from graph_kmer_index_b200 import distributed, synthetic
    from oracle import c_oracle
    rank, world = distributed.init_process_group("nccl")
    n, n_nodes, modulo, k = 200_000, 5_000, 1_000_003, 31
    hashes, nodes, ref, af = synthetic.flat_kmers(n, n_nodes, k)
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    reads = synthetic.reads(20_001, 150, n, k, p_hit_permille=300, n_permille=3)
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], modulo)
    got = distributed.count_reads_sharded(dev, reads, k, min_nodes=n_nodes)
    want = c_oracle.read_node_counts(idx, reads, k, n_nodes)
    assert np.array_equal(got, want), (rank, float(got.sum()), float(want.sum()))
    dist.barrier()
    if rank == 0:
        print("OK", world, int(want.sum()))
    dist.destroy_process_group()
""")


def test_two_gpu_sharded_count(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=dict(os.environ))
    assert out.returncode == 0, out.stdout[-3000:]
    assert "OK 2" in out.stdout
