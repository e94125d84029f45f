"""Drop-in for graph_kmer_index/flat_kmers.py: FlatKmers / FlatKmers2 containers and their npz format."""
import logging

import numpy as np

from . import _lib
from .kmer_hashing import (kmer_hashes_to_reverse_complement_hash, letter_sequence_to_numeric,  # noqa: F401 (re-export)
                           numeric_to_letter_sequence)


class FlatKmers2:
    """flat_kmers.py:7-29."""

    def __init__(self, hashes, start_nodes, start_offsets, nodes, allele_frequencies):
        assert len(hashes) == len(nodes)
        assert len(start_nodes) == len(nodes)
        assert len(start_offsets) == len(start_nodes)
        self._hashes = hashes
        self._nodes = nodes
        self._start_nodes = start_nodes
        self._start_offsets = start_offsets
        if allele_frequencies is None:
            self._allele_frequencies = np.zeros(len(self._hashes), dtype=np.single) + 1.0
        else:
            self._allele_frequencies = allele_frequencies

    def __str__(self):
        return "\n".join([str(data) for data in [self._hashes, self._nodes]])

    __repr__ = __str__


class FlatKmers:
    """flat_kmers.py:32-131.  Column arrays are host numpy arrays exactly as in the reference."""

    def __init__(self, hashes, nodes, ref_offsets=None, allele_frequencies=None):
        assert len(hashes) == len(nodes)
        self._hashes = hashes
        self._nodes = nodes
        self._ref_offsets = np.zeros(len(self._nodes)) if ref_offsets is None else ref_offsets      # :37-40 (float64 zeros)
        if allele_frequencies is None:                                                              # :42-46
            self._allele_frequencies = np.zeros(len(self._hashes), dtype=np.single) + 1.0
        else:
            self._allele_frequencies = allele_frequencies

    def describtion(self):
        return "".join("%d: %d\n" % (kmer, node) for kmer, node in zip(self._hashes, self._nodes))

    @classmethod
    def from_file(cls, file_name):
        """flat_kmers.py:55-63."""
        try:
            data = np.load(file_name)
        except FileNotFoundError:
            data = np.load(file_name + ".npz")
        return cls(data["hashes"], data["nodes"], data["ref_offsets"], data["allele_frequencies"])

    def to_file(self, file_name):
        """flat_kmers.py:65-68: npz keys hashes, nodes, ref_offsets, allele_frequencies."""
        np.savez(file_name, hashes=self._hashes, nodes=self._nodes, ref_offsets=self._ref_offsets,
                 allele_frequencies=self._allele_frequencies)

    @classmethod
    def from_multiple_flat_kmers(cls, flat_kmers_list):
        """flat_kmers.py:71-90 (dtypes forced to uint64 / uint32 / uint64 / float32)."""
        flat_kmers_list = list(flat_kmers_list)
        cat = lambda arrays, dt: np.concatenate([np.asarray(a) for a in arrays]).astype(dt, copy=False) if arrays else np.array([], dtype=dt)
        hashes = cat([f._hashes for f in flat_kmers_list], np.uint64)
        nodes = cat([f._nodes for f in flat_kmers_list], np.uint32)
        refs = [f._ref_offsets for f in flat_kmers_list if f._ref_offsets is not None]
        ref_offsets = cat(refs, np.uint64) if sum(len(r) for r in refs) else None
        af = cat([f._allele_frequencies for f in flat_kmers_list], np.single)
        return FlatKmers(hashes, nodes, ref_offsets, af)

    def sum_of_kmer_frequencies(self, kmer_index_with_frequencies):
        if hasattr(kmer_index_with_frequencies, "get_frequencies"):       # one batched lookup instead of one per k-mer
            return int(np.maximum(1, kmer_index_with_frequencies.get_frequencies(self._hashes)).sum()) if len(self._hashes) else 0
        return sum([0] + [max(1, kmer_index_with_frequencies.get_frequency(int(kmer))) for kmer in self._hashes])

    def maximum_kmer_frequency(self, kmer_index_with_frequencies):
        if hasattr(kmer_index_with_frequencies, "get_frequencies"):
            return int(kmer_index_with_frequencies.get_frequencies(self._hashes).max()) if len(self._hashes) else 0
        return max([0] + [kmer_index_with_frequencies.get_frequency(int(kmer)) for kmer in self._hashes])

    def get_new_without_singletons(self):
        """flat_kmers.py:98-125: drop the first occurrence of every hash; the survivors keep their order.
        The first-occurrence marking runs on the device (gki_mark_non_first_occurrences)."""
        h = np.ascontiguousarray(np.asarray(self._hashes).astype(np.uint64, copy=False))
        keep = np.empty(len(h), dtype=np.uint8)
        _lib.call("gki_mark_non_first_occurrences", _lib.ptr(h), len(h), _lib.ptr(keep), _lib.current_stream())
        keep = keep.astype(bool)
        return FlatKmers(np.asarray(self._hashes)[keep], np.asarray(self._nodes)[keep],
                         np.asarray(self._ref_offsets)[keep], np.asarray(self._allele_frequencies)[keep])

    def get_reverse_complement_flat_kmers(self, k):
        """flat_kmers.py:127-131."""
        return FlatKmers(kmer_hashes_to_reverse_complement_hash(self._hashes, k), self._nodes, self._ref_offsets,
                         self._allele_frequencies)
