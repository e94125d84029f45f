"""Drop-in for graph_kmer_index/reference_kmer_index.py: k-mers ordered by reference position.  `from_flat_kmers` sorts
by ref_offset on the device (gki_group_by_key, GKI_GROUP_REFERENCE_INDEX) and `from_sequence` is K1 over a whole
chromosome (gki_hash_reads_ragged on one long read)."""
import logging

import numpy as np

from .kmer_hashing import power_array
from .read_kmers import ReadKmers
from .reverse_kmer_index import gather, group_by_key


def fill_zeros_from_end(array):
    """reference_kmer_index.py:16-21 (host helper kept for API parity; from_flat_kmers does this on the device)."""
    array = array[::-1]
    prev = np.arange(len(array))
    prev[array == 0] = 0
    prev = np.maximum.accumulate(prev)
    return array[prev][::-1]


class ReferenceKmerIndex:
    """reference_kmer_index.py:24-160: k-mers in reference order; `ref_position_to_index[p]` is the first entry at or after
    reference position p."""
    properties = {"ref_position_to_index", "kmers", "ref_positions", "nodes"}

    def __init__(self, ref_position_to_index=None, kmers=None, ref_positions=None, nodes=None):
        self.ref_position_to_index, self.kmers = ref_position_to_index, kmers
        self.ref_positions, self.nodes = ref_positions, nodes

    def _entry(self, ref_position):
        return self.ref_position_to_index[ref_position]

    def get_between(self, ref_start, ref_end):
        """:31-34 (the end is clamped to the last known position)."""
        last = len(self.ref_position_to_index) - 1
        return self.kmers[self._entry(ref_start):self._entry(min(last, ref_end))]

    def get_between_except(self, ref_start, ref_end, except_position):
        """:36-39 (linear-reference indexes only: entry i is position i)."""
        assert self.ref_positions is None
        wanted = np.arange(ref_start, ref_end)
        return self.kmers[[int(i) for i in wanted[wanted != except_position]]]

    def get_all_between(self, ref_start, ref_end):
        """:41-47."""
        if self.ref_positions is None:
            raise Exception("This index is missing reference positions and cannot be used to get all between. "
                            "Is it made from a linear reference? If so, use get_between() instead")
        span = slice(self._entry(ref_start), self._entry(ref_end))
        return self.kmers[span], self.ref_positions[span], self.nodes[span]

    @classmethod
    def from_sequence(cls, genome_sequence, k, only_store_kmers=False):
        """:50-67: K1 over the whole sequence; 32-bit k-mers when k <= 16."""
        kmers = ReadKmers.get_kmers_from_read_dynamic(genome_sequence, power_array(k)).astype(np.uint32 if k <= 16 else np.uint64)
        positions = None if only_store_kmers else np.arange(0, len(genome_sequence), dtype=np.uint32)
        return cls(positions, kmers)

    @classmethod
    def from_linear_reference(cls, fasta_file_name, reference_name="ref", k=15, only_store_kmers=False):
        """:69-74 without pyfaidx: the record named `reference_name` of a plain FASTA file."""
        pieces, inside = [], False
        with open(fasta_file_name) as fasta:
            for line in fasta:
                if line.startswith(">"):
                    name = line[1:].split()
                    inside = bool(name) and name[0] == reference_name
                elif inside:
                    pieces.append(line.strip())
        return cls.from_sequence("".join(pieces), k, only_store_kmers)

    @classmethod
    def from_flat_kmers(cls, flat_kmers):
        """:76-121 on the device (stable sort; see ReverseKmerIndex.from_flat_kmers)."""
        ref_positions = np.asarray(flat_kmers._ref_offsets)
        assert len(ref_positions) < 4294967295, "Too many kmers to store (32 bit limit reached). There are %d kmers" % len(ref_positions)
        perm, position_to_index, _ = group_by_key(ref_positions, int(np.max(ref_positions)) + 1, reference_index=True)
        kmers = gather(flat_kmers._hashes, perm)
        if np.max(kmers) < 2 ** 32:
            logging.warning("Storing kmers as 32 bit uint since max hash is low enough")
            kmers = kmers.astype(np.uint32)
        return cls(position_to_index, kmers, gather(ref_positions, perm), gather(flat_kmers._nodes, perm))

    def to_file(self, file_name):
        """:123-138: only the arrays the index has."""
        arrays = {"kmers": self.kmers}
        if self.ref_position_to_index is not None:
            arrays["ref_position_to_index"] = self.ref_position_to_index
            if self.ref_positions is not None or self.nodes is not None:
                arrays.update(ref_positions=self.ref_positions, nodes=self.nodes)
        np.savez(file_name, **arrays)

    @classmethod
    def from_file(cls, file_name):
        """:140-160: `file_name + ".npz"`, else `file_name`; absent arrays become None."""
        try:
            archive = np.load(file_name + ".npz")
        except FileNotFoundError:
            archive = np.load(file_name)
        optional = {name: (archive[name] if name in archive else None) for name in ("ref_position_to_index", "ref_positions", "nodes")}
        return cls(optional["ref_position_to_index"], archive["kmers"], optional["ref_positions"], optional["nodes"])
