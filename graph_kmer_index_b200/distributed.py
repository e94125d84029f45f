"""Multi-GPU counting: reads shard across ranks, the index is replicated, node-count vectors are combined by one
all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  One process per GPU, launched with torchrun.

The reference's analogue is has_kmers_parallel (collision_free_kmer_index.py:222-232) /
run_numpy_based_function_in_parallel (shared_mem.py:141-176): np.linspace slices of the query array over worker
processes, results concatenated / summed on the parent."""
import os

import numpy as np


def shard_bounds(n_items, rank, world_size):
    """Contiguous, balanced [start, end) of `rank` -- the same np.linspace split as shared_mem.py:164-166."""
    edges = np.linspace(0, n_items, world_size + 1).astype(np.int64)
    return int(edges[rank]), int(edges[rank + 1])


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op when WORLD_SIZE is 1 or unset)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def allreduce_node_counts(counts):
    """Sum per-rank node counts in place.  `counts`: torch float64 tensor (device tensor for NCCL).  float64 sums of
    integer counts are exact below 2^53, so the result equals the single-GPU count bit for bit."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def count_reads_sharded(device_index, reads, k, min_nodes=0, both_strands=True, rank=None, world_size=None):
    """Count this rank's shard of `reads` (every rank holds the full array, or pass a pre-sharded array with
    rank=0, world_size=1) and all-reduce the node counts.  Returns a float64 numpy array on every rank."""
    import torch
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(reads.shape[0], rank, world_size)
    device_index.count_reads(reads[lo:hi], k, both_strands)
    n_out = max(int(min_nodes), device_index.max_node + 1)
    use_cuda = dist.is_initialized() and dist.get_backend() == "nccl"
    counts = torch.zeros(n_out, dtype=torch.float64, device="cuda" if use_cuda else "cpu")
    if use_cuda:
        device_index.node_counts(min_nodes, out=counts)
    else:
        counts.copy_(torch.from_numpy(device_index.node_counts(min_nodes)))
    allreduce_node_counts(counts)
    return counts.cpu().numpy()
