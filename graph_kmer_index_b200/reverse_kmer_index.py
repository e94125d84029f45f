"""Drop-in for graph_kmer_index/reverse_kmer_index.py: node -> (k-mers, ref positions).  `from_flat_kmers` groups the
entries by node on the device (gki_group_by_key: the K2 radix sort + run heads, keyed by node instead of bucket)."""
import logging

import numpy as np

from . import _lib

GKI_GROUP_REFERENCE_INDEX = 1


def group_by_key(keys, n_keys, reference_index=False, want_counts=True):
    """Stable grouping of `keys` (non-negative integers < n_keys <= 2**32) -> (perm u32, first u32[n_keys], counts u32[n_keys] | None)."""
    keys = np.ascontiguousarray(keys)
    if keys.dtype.itemsize not in (4, 8) or keys.dtype.kind not in "ui":
        keys = keys.astype(np.uint64)            # float64 ref_offsets of a defaulted FlatKmers (flat_kmers.py:37-40), small ints
    n = len(keys)
    perm = np.empty(n, dtype=np.uint32)
    first = np.empty(n_keys, dtype=np.uint32)
    counts = np.empty(n_keys, dtype=np.uint32) if want_counts and not reference_index else None
    _lib.call("gki_group_by_key", _lib.ptr(keys), keys.dtype.itemsize, n, int(n_keys), GKI_GROUP_REFERENCE_INDEX if reference_index else 0,
              _lib.ptr(perm), _lib.ptr(first), _lib.ptr(counts), _lib.current_stream())
    return perm, first, counts


def gather(column, perm):
    """column[perm] on the device for 1/2/4/8-byte items (any dtype)."""
    column = np.ascontiguousarray(column)
    if column.dtype.itemsize not in (1, 2, 4, 8):
        return column[perm]
    out = np.empty(len(perm), dtype=column.dtype)
    _lib.call("gki_gather", _lib.ptr(column), column.dtype.itemsize, _lib.ptr(perm), len(perm), _lib.ptr(out), _lib.current_stream())
    return out


class ReverseKmerIndex:
    """reverse_kmer_index.py:5-84: for every node the k-mers (and their reference positions) that touch it, stored as one run per
    node of `hashes` / `ref_positions`; `nodes_to_index_positions[node]` is where the run starts, `nodes_to_n_hashes[node]` its length."""
    properties = {"nodes_to_index_positions", "nodes_to_n_hashes", "hashes", "ref_positions"}
    _FIELDS = ("nodes_to_index_positions", "nodes_to_n_hashes", "hashes", "ref_positions")

    def __init__(self, nodes_to_index_positions=None, nodes_to_n_hashes=None, hashes=None, ref_positions=None):
        for name, value in zip(self._FIELDS, (nodes_to_index_positions, nodes_to_n_hashes, hashes, ref_positions)):
            setattr(self, name, value)

    def __str__(self):
        labels = ("Nodes to index positions: ", "Nodes to n hashes      : ", "Hashes:                  ", "Ref positions:                  ")
        return "".join("%s%s\n" % (label, getattr(self, name)) for label, name in zip(labels, self._FIELDS))

    def _run(self, node):
        """[start, end) of the node's run; an unknown node raises IndexError like the reference's array lookup does."""
        start = int(self.nodes_to_index_positions[node])
        return start, start + int(self.nodes_to_n_hashes[node])

    def get_node_kmers(self, node):
        """reverse_kmer_index.py:23-29 ([] for a node without k-mers)."""
        start, end = self._run(node)
        return [] if end == start else self.hashes[start:end]

    def get_node_kmers_and_ref_positions(self, node):
        """reverse_kmer_index.py:31-42 ([[], []] for a node without k-mers)."""
        try:
            start, end = self._run(node)
        except IndexError:
            logging.error("Invalid node %d" % node)
            raise
        if end == start:
            return [[], []]
        return self.hashes[start:end], self.ref_positions[start:end]

    @classmethod
    def from_file(cls, file_name):
        """reverse_kmer_index.py:44-51: `file_name`, else `file_name + ".npz"`."""
        try:
            archive = np.load(file_name)
        except FileNotFoundError:
            archive = np.load(file_name + ".npz")
        return cls(*[archive[name] for name in cls._FIELDS])

    def to_file(self, file_name):
        """reverse_kmer_index.py:53-57 (uncompressed npz, keys = attribute names)."""
        np.savez(file_name, **{name: getattr(self, name) for name in self._FIELDS})

    @classmethod
    def from_flat_kmers(cls, flat_kmers):
        """reverse_kmer_index.py:59-84.  The sort is stable (the reference's argsort is not: the order of the entries of
        one node is only defined up to a permutation there); n_kmers is uint16 like the reference's and wraps the same way."""
        nodes = np.asarray(flat_kmers._nodes)
        perm, first, counts = group_by_key(nodes, int(np.max(nodes)) + 1)
        return cls(first, counts.astype(np.uint16), gather(flat_kmers._hashes, perm), gather(flat_kmers._ref_offsets, perm))
