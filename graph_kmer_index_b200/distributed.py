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


class CountsComm:
    """The communicator of the path's one collective, made and used through the C ABI (include/gki.h: gki_nccl_unique_id,
    gki_nccl_comm_create, gki_allreduce_counts) -- what a non-Python host does too.  The 128-byte NCCL id travels from rank 0 over the
    torch.distributed group that torchrun set up; creation is collective (every rank constructs one)."""

    def __init__(self):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            _lib.call("gki_nccl_unique_id", ident.data_ptr())
        on_device = dist.get_backend() == "nccl"
        moved = ident.cuda() if on_device else ident
        dist.broadcast(moved, src=0)
        ident = moved.cpu().contiguous()
        self.handle = ctypes.c_void_p()
        _lib.call("gki_nccl_comm_create", ident.data_ptr(), self.rank, self.world, ctypes.byref(self.handle))

    def allreduce(self, counts):
        """in-place sum of a float64 device tensor over the ranks, on torch's current stream"""
        import torch
        from . import _lib
        assert counts.is_cuda and counts.is_contiguous() and counts.dtype == torch.float64
        _lib.call("gki_allreduce_counts", self.handle, counts.data_ptr(), counts.numel(), _lib.GKI_COUNTS_FLOAT64, _lib.current_stream())
        return counts

    def close(self):
        from . import _lib
        if getattr(self, "handle", None) is not None and self.handle.value:
            _lib.call("gki_nccl_comm_destroy", self.handle)
            self.handle = None


_counts_comm = None


def counts_comm():
    """the process's CountsComm, created on first use (collective: every rank must reach its first all-reduce)"""
    global _counts_comm
    if _counts_comm is None:
        _counts_comm = CountsComm()
    return _counts_comm


def allreduce_node_counts(counts):
    """Sum per-rank node counts in place.  `counts`: torch float64 tensor.  float64 sums of integer counts are exact below 2^53, so
    the result equals the single-GPU count bit for bit.  Device tensors under the NCCL backend go through the C ABI
    (gki_allreduce_counts on a communicator made by gki_nccl_comm_create); host tensors (gloo, the CPU tests) through torch."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if counts.is_cuda and dist.get_backend() == "nccl" and os.environ.get("GKI_TORCH_ALLREDUCE", "0") != "1":
            counts_comm().allreduce(counts)
        else:
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


def bucket_range(modulo, rank, world_size):
    """Buckets owned by `rank` in the hash-range partitioned build: [lo, hi)."""
    part = (int(modulo) + world_size - 1) // world_size
    return min(rank * part, int(modulo)), min((rank + 1) * part, int(modulo))


def build_index_partitioned(hashes, nodes, ref_offsets, allele_frequencies, modulo, skip_frequencies=True, replicate=False):
    """Hash-range partitioned index build over the ranks of the default process group (NCCL, one process per GPU).

    Every rank passes its own shard of the FlatKmers as device tensors (int64 / int32 / int64 / float32 views of the
    reference's uint64 / uint32 / uint64 / float32 columns); the global input order is rank-major.  Every rank orders its
    shard by the rank that owns the bucket range and packs it into 32-byte records in one pass (gki_partition_pack); the
    per-destination counts of all ranks travel with one all-gather, the records with ONE all-to-all; each rank builds its
    slice straight from the records it received (gki_index_build_records), and the slices concatenated in rank order are
    exactly the single-GPU index of the concatenated FlatKmers.  Returns a dict of device tensors: the local slice, or --
    with replicate=True -- the whole index on every rank (one all-gather per array)."""
    import torch
    import torch.distributed as dist
    from . import _lib
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = hashes.device
    n = int(hashes.shape[0])
    stream = _lib.current_stream()
    have = {"nodes": nodes is not None, "ref_offsets": ref_offsets is not None, "allele_frequencies": allele_frequencies is not None}
    # 1. order the local entries by owner, packed
    records = torch.empty((max(n, 1), 4), dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    _lib.call("gki_partition_pack", _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref_offsets), _lib.ptr(allele_frequencies), n, int(modulo), world,
              _lib.ptr(records), _lib.ptr(counts), stream)
    # 2. who sends how much to whom: one all-gather of the count vectors, one read-back
    matrix = torch.empty((world, world), dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(matrix, counts)
    else:
        matrix[0] = counts
    matrix = matrix.cpu()
    in_splits, out_splits = matrix[rank].tolist(), matrix[:, rank].tolist()
    totals = matrix.sum(dim=0).tolist()                      # entries owned by every rank
    n_local, offset = int(totals[rank]), int(sum(totals[:rank]))
    # 3. ONE all-to-all of the records
    recv = torch.empty((max(n_local, 1), 4), dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_to_all_single(recv[:n_local], records[:n], output_split_sizes=out_splits, input_split_sizes=in_splits)
    else:
        recv[:n_local].copy_(records[:n])
    del records
    # 4. local slice, built from the records
    lo, hi = bucket_range(modulo, rank, world)
    h2i = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    nk = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    outs = {"kmers": torch.empty(n_local, dtype=torch.int64, device=dev)}
    if have["nodes"]:
        outs["nodes"] = torch.empty(n_local, dtype=torch.int32, device=dev)
    if have["ref_offsets"]:
        outs["ref_offsets"] = torch.empty(n_local, dtype=torch.int64, device=dev)
    if have["allele_frequencies"]:
        outs["allele_frequencies"] = torch.empty(n_local, dtype=torch.float32, device=dev)
    freq = torch.empty(n_local, dtype=torch.int16, device=dev)
    if n_local:
        _lib.call("gki_index_build_records", _lib.ptr(recv), n_local, int(modulo), lo, hi, offset,
                  _lib.GKI_BUILD_SKIP_FREQUENCIES if skip_frequencies else 0, _lib.ptr(h2i), _lib.ptr(nk), _lib.ptr(outs["kmers"]),
                  _lib.ptr(outs.get("nodes")), _lib.ptr(outs.get("ref_offsets")), _lib.ptr(outs.get("allele_frequencies")), _lib.ptr(freq), stream)
    else:
        h2i.zero_()
        nk.zero_()
    del recv
    local = dict(hashes_to_index=h2i, n_kmers=nk, frequencies=freq, bucket_range=(lo, hi), position_offset=offset, n_total=int(sum(totals)), **outs)
    if not replicate or world == 1:
        return local
    # 5. replicate: one all-gather per array (slices padded to the largest, NCCL gathers equal sizes), cut back to size on arrival
    n_total = int(sum(totals))

    def gather(local_tensor, sizes, total):
        widest = max(sizes)
        padded = torch.zeros(widest, dtype=local_tensor.dtype, device=dev)
        padded[:local_tensor.shape[0]] = local_tensor
        everyone = torch.empty((world, widest), dtype=local_tensor.dtype, device=dev)
        dist.all_gather_into_tensor(everyone.view(torch.uint8).view(-1), padded.view(torch.uint8))      # byte view: NCCL has no int16
        full = torch.empty(total, dtype=local_tensor.dtype, device=dev)
        at = 0
        for r in range(world):
            full[at:at + sizes[r]] = everyone[r, :sizes[r]]
            at += sizes[r]
        return full

    table_sizes = [bucket_range(modulo, r, world)[1] - bucket_range(modulo, r, world)[0] for r in range(world)]
    full = dict(hashes_to_index=gather(h2i, table_sizes, int(modulo)), n_kmers=gather(nk, table_sizes, int(modulo)),
                frequencies=gather(freq, totals, n_total))
    for name, col in outs.items():
        full[name] = gather(col, totals, n_total)
    full.update(bucket_range=(0, int(modulo)), position_offset=0, n_total=n_total)
    return full


class PartitionedCounterIndex:
    """Counting against an index that is hash-range PARTITIONED over the ranks instead of replicated -- for indexes larger than
    one GPU's memory.  Every rank passes its shard of the FlatKmers (device tensors, as build_index_partitioned); the entries travel to
    the rank that owns their bucket range (gki_partition_pack -> one all-to-all) and every rank builds an ordinary index over the
    entries it owns.  count_reads hashes the rank's own reads (K1, both strands), routes every k-mer hash to the owner of its bucket
    (kmer % modulo, the same ranges) with one all-to-all per chunk of reads, and the owner counts what it receives
    (gki_count_kmers).  get_node_counts sums the ranks' node counts with one all-reduce: every index entry lives on exactly one rank.
    The reference's analogue of the routing is has_kmers_parallel's slicing of the queries (cfki:222-232); results equal the
    replicated path's bit for bit.  Each read k-mer crosses NVLink as 8 bytes here, where the replicated path reads 150 bytes per
    read: this is the fallback for indexes that do not fit, not the fast path."""

    def __init__(self, hashes, nodes, modulo):
        import torch
        import torch.distributed as dist
        from . import _lib
        from .collision_free_kmer_index import DeviceIndex
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.modulo = int(modulo)
        self.dev = hashes.device
        n = int(hashes.shape[0])
        stream = _lib.current_stream()
        records = torch.empty((max(n, 1), 4), dtype=torch.int64, device=self.dev)
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        _lib.call("gki_partition_pack", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, self.modulo, self.world, _lib.ptr(records), _lib.ptr(counts), stream)
        matrix = torch.empty((self.world, self.world), dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_gather_into_tensor(matrix, counts)
        else:
            matrix[0] = counts
        matrix = matrix.cpu()
        in_splits, out_splits = matrix[self.rank].tolist(), matrix[:, self.rank].tolist()
        n_local = int(sum(out_splits))
        recv = torch.empty((max(n_local, 1), 4), dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_to_all_single(recv[:n_local], records[:n], output_split_sizes=out_splits, input_split_sizes=in_splits)
        else:
            recv[:n_local].copy_(records[:n])
        del records
        # an ordinary index over the owned entries (tables span the whole modulo: buckets of other ranks stay empty)
        h2i = torch.empty(self.modulo, dtype=torch.int32, device=self.dev)
        nk = torch.empty(self.modulo, dtype=torch.int32, device=self.dev)
        kmers = torch.empty(max(n_local, 1), dtype=torch.int64, device=self.dev)
        nds = torch.empty(max(n_local, 1), dtype=torch.int32, device=self.dev)
        self.n_local = n_local
        self.index = None
        max_node = torch.zeros(1, dtype=torch.int64, device=self.dev)
        if n_local:
            _lib.call("gki_index_build_records", _lib.ptr(recv), n_local, self.modulo, 0, self.modulo, 0, _lib.GKI_BUILD_SKIP_FREQUENCIES, _lib.ptr(h2i),
                      _lib.ptr(nk), _lib.ptr(kmers), _lib.ptr(nds), None, None, None, stream)
            self.index = DeviceIndex(h2i, nk, kmers[:n_local], nds[:n_local], self.modulo)
            max_node[0] = self.index.max_node
        if self.world > 1:
            dist.all_reduce(max_node, op=dist.ReduceOp.MAX)
        self.max_node = int(max_node.item())
        del recv, h2i, nk, kmers, nds

    def reset_counts(self):
        if self.index is not None:
            self.index.reset_counts()

    def count_reads(self, reads, k, chunk_reads=1 << 19):
        """reads: this rank's (n_reads, L) uint8 ASCII rows (numpy, or a torch tensor on host / device).  Every rank must call it
        (the exchange is collective) with the same chunk_reads; ranks with fewer chunks send empty ones."""
        import numpy as np
        import torch
        import torch.distributed as dist
        from . import _lib
        from .read_kmers import hash_read_matrix
        n_reads = int(reads.shape[0])
        n_chunks = torch.tensor([(n_reads + chunk_reads - 1) // chunk_reads], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(n_chunks, op=dist.ReduceOp.MAX)
        stream = _lib.current_stream()
        for c in range(int(n_chunks.item())):
            part = reads[c * chunk_reads:(c + 1) * chunk_reads]
            if isinstance(part, np.ndarray):
                part = torch.from_numpy(np.ascontiguousarray(part))
            part = part.to(self.dev)
            m = int(part.shape[0])
            if m and part.shape[1] >= k:
                fwd, rc = hash_read_matrix(part, k)                                  # K1: both strands
                hashes = torch.cat([fwd.reshape(-1).view(torch.int64), rc.reshape(-1).view(torch.int64)])
                del fwd, rc
            else:
                hashes = torch.empty(0, dtype=torch.int64, device=self.dev)
            nq = int(hashes.shape[0])
            perm = torch.empty(max(nq, 1), dtype=torch.int32, device=self.dev)
            counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
            _lib.call("gki_partition_by_bucket_range", _lib.ptr(hashes), nq, self.modulo, self.world, _lib.ptr(perm), _lib.ptr(counts), stream)
            grouped = torch.empty(max(nq, 1), dtype=torch.int64, device=self.dev)
            if nq:
                _lib.call("gki_gather", _lib.ptr(hashes), 8, _lib.ptr(perm), nq, _lib.ptr(grouped), stream)
            matrix = torch.empty((self.world, self.world), dtype=torch.int64, device=self.dev)
            if self.world > 1:
                dist.all_gather_into_tensor(matrix, counts)
            else:
                matrix[0] = counts
            matrix = matrix.cpu()
            in_splits, out_splits = matrix[self.rank].tolist(), matrix[:, self.rank].tolist()
            n_in = int(sum(out_splits))
            recv = torch.empty(max(n_in, 1), dtype=torch.int64, device=self.dev)
            if self.world > 1:
                dist.all_to_all_single(recv[:n_in], grouped[:nq], output_split_sizes=out_splits, input_split_sizes=in_splits)
            else:
                recv[:n_in].copy_(grouped[:nq])
            if n_in and self.index is not None:
                self.index.count_kmers(recv[:n_in])
            del hashes, perm, grouped, recv

    def get_node_counts(self, min_nodes=0):
        """float64 numpy array on every rank: the node counts of the whole (partitioned) index."""
        import torch
        n_out = max(int(min_nodes), self.max_node + 1)
        counts = torch.zeros(n_out, dtype=torch.float64, device=self.dev)
        if self.index is not None:
            self.index.node_counts(n_out, out=counts)
        allreduce_node_counts(counts)
        return counts.cpu().numpy()


def critical_path_chunks(n_paths, n_chunks):
    """(start, end) chunks of the critical paths exactly as `graph_kmer_index index -t T` cuts them
    (command_line_interface.py:588-603: n_paths // n_chunks paths per chunk, the remainder in further chunks)."""
    n_chunks = max(1, min(int(n_chunks), int(n_paths)))
    per = n_paths // n_chunks
    starts = list(range(0, n_paths, per))
    ends = starts[1:] + [n_paths]
    return list(zip(starts, ends))


def find_kmers_sharded(graph, k, critical_graph_paths=None, n_chunks=None, rank=None, world_size=None, gather=True, **finder_kwargs):
    """DenseKmerFinder over all critical paths, chunked as the reference's multi-process `index` command does
    (command_line_interface.py:575-617) with the chunks dealt round-robin to the ranks (one process per GPU; the graph is
    replicated).  Every chunk is an independent DenseKmerFinder run with start/stop_at_critical_path_number; the result is the
    chunks' FlatKmers concatenated in chunk order (FlatKmers.from_multiple_flat_kmers, cli:608) -- identical to what the
    reference produces for the same n_chunks.  With gather=False every rank returns only its own chunks [(chunk_no, FlatKmers)]."""
    import torch.distributed as dist
    from .flat_kmers import FlatKmers
    from .kmer_finder import CriticalGraphPaths, DenseKmerFinder, graph_arrays
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    arrays = graph_arrays(graph)
    if critical_graph_paths is None:
        critical_graph_paths = CriticalGraphPaths.from_graph(arrays, k)
    chunks = critical_path_chunks(len(critical_graph_paths), n_chunks if n_chunks is not None else world_size * 20)
    mine = []
    for c in range(rank, len(chunks), world_size):
        finder = DenseKmerFinder(arrays, k, critical_graph_paths=critical_graph_paths, start_at_critical_path_number=chunks[c][0],
                                 stop_at_critical_path_number=chunks[c][1], **finder_kwargs)
        finder.find()
        mine.append((c, finder.get_flat_kmers(v="0")))
    if not gather:
        return mine
    if world_size > 1:
        everyone = [None] * world_size
        dist.all_gather_object(everyone, mine)
        mine = [item for part in everyone for item in part]
    mine.sort(key=lambda item: item[0])
    return FlatKmers.from_multiple_flat_kmers([flat for _, flat in mine])
