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
    # hash-range partitioned build: every rank contributes a shard of the FlatKmers, the replicated result is the oracle's index
    lo, hi = distributed.shard_bounds(n, rank, world)
    t = lambda a, dt: torch.from_numpy(a[lo:hi].view(dt)).cuda()
    full = distributed.build_index_partitioned(t(hashes, np.int64), t(nodes, np.int32), t(ref, np.int64), torch.from_numpy(af[lo:hi]).cuda(),
                                               modulo, skip_frequencies=False, replicate=True)
    w2 = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=False)
    assert np.array_equal(full["hashes_to_index"].cpu().numpy(), w2["_hashes_to_index"]), rank
    assert np.array_equal(full["n_kmers"].cpu().numpy().view(np.uint32), w2["_n_kmers"]), rank
    assert np.array_equal(full["kmers"].cpu().numpy().view(np.uint64), w2["_kmers"]), rank
    assert np.array_equal(full["nodes"].cpu().numpy().view(np.uint32), w2["_nodes"]), rank
    assert np.array_equal(full["ref_offsets"].cpu().numpy().view(np.uint64), w2["_ref_offsets"]), rank
    assert np.array_equal(full["allele_frequencies"].cpu().numpy(), w2["_allele_frequencies"]), rank
    assert np.array_equal(full["frequencies"].cpu().numpy().view(np.uint16), w2["_frequencies"]), rank
    # hash-range partitioned index for counting (an index too large to replicate): entries and read k-mers are routed to the owners
    pc = distributed.PartitionedCounterIndex(t(hashes, np.int64), t(nodes, np.int32), modulo)
    my_lo, my_hi = distributed.shard_bounds(len(reads), rank, world)
    pc.count_reads(reads[my_lo:my_hi], k, chunk_reads=4096)
    got_p = pc.get_node_counts(n_nodes)
    assert np.array_equal(got_p, want), (rank, float(got_p.sum()), float(want.sum()))
    pc.reset_counts()
    pc.count_reads(torch.from_numpy(reads[my_lo:my_hi]).cuda(), k)
    assert np.array_equal(pc.get_node_counts(n_nodes), want), rank
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
