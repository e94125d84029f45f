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
    """reference_kmer_index.py:24-160."""
    properties = {"ref_position_to_index", "kmers", "ref_positions", "nodes"}

    def __init__(self, ref_position_to_index=None, kmers=None, ref_positions=None, nodes=None):
        self.ref_position_to_index = ref_position_to_index
        self.kmers = kmers
        self.ref_positions = ref_positions
        self.nodes = nodes

    def get_between(self, ref_start, ref_end):
        return self.kmers[self.ref_position_to_index[ref_start]:
                          self.ref_position_to_index[min(len(self.ref_position_to_index) - 1, ref_end)]]

    def get_between_except(self, ref_start, ref_end, except_position):
        assert self.ref_positions is None
        indexes = [i for i in np.arange(ref_start, ref_end) if i != except_position]
        return self.kmers[indexes]

    def get_all_between(self, ref_start, ref_end):
        if self.ref_positions is None:
            raise Exception("This index is missing reference positions and cannot be used to get all between. "
                            "Is it made from a linear reference? If so, use get_between() instead")
        start = self.ref_position_to_index[ref_start]
        end = self.ref_position_to_index[ref_end]
        return self.kmers[start:end], self.ref_positions[start:end], self.nodes[start:end]

    @classmethod
    def from_sequence(cls, genome_sequence, k, only_store_kmers=False):
        """reference_kmer_index.py:50-67."""
        kmers = ReadKmers.get_kmers_from_read_dynamic(genome_sequence, power_array(k))
        ref_position_to_index = None
        if not only_store_kmers:
            ref_position_to_index = np.arange(0, len(genome_sequence), dtype=np.uint32)
        kmers = kmers.astype(np.uint32) if k <= 16 else kmers.astype(np.uint64)
        return cls(ref_position_to_index, kmers)

    @classmethod
    def from_linear_reference(cls, fasta_file_name, reference_name="ref", k=15, only_store_kmers=False):
        """reference_kmer_index.py:69-74 without pyfaidx: the named record of a plain FASTA file."""
        sequence, keep = [], False
        with open(fasta_file_name) as f:
            for line in f:
                if line.startswith(">"):
                    keep = line[1:].split()[0] == reference_name if line[1:].split() else False
                elif keep:
                    sequence.append(line.strip())
        return cls.from_sequence("".join(sequence), k, only_store_kmers)

    @classmethod
    def from_flat_kmers(cls, flat_kmers):
        """reference_kmer_index.py:76-121 (stable sort; see ReverseKmerIndex.from_flat_kmers)."""
        ref_positions = np.asarray(flat_kmers._ref_offsets)
        assert len(ref_positions) < 4294967295, "Too many kmers to store (32 bit limit reached). There are %d kmers" % len(ref_positions)
        last = int(np.max(ref_positions))
        perm, ref_position_to_index, _ = group_by_key(ref_positions, last + 1, reference_index=True)
        kmers = gather(flat_kmers._hashes, perm)
        if np.max(kmers) < 2 ** 32:
            logging.warning("Storing kmers as 32 bit uint since max hash is low enough")
            kmers = kmers.astype(np.uint32)
        return cls(ref_position_to_index, kmers, gather(ref_positions, perm), gather(flat_kmers._nodes, perm))

    def to_file(self, file_name):
        """reference_kmer_index.py:123-138."""
        if self.ref_position_to_index is None:
            np.savez(file_name, kmers=self.kmers)
        elif self.ref_positions is None and self.nodes is None:
            np.savez(file_name, ref_position_to_index=self.ref_position_to_index, kmers=self.kmers)
        else:
            np.savez(file_name, ref_position_to_index=self.ref_position_to_index, kmers=self.kmers,
                     ref_positions=self.ref_positions, nodes=self.nodes)

    @classmethod
    def from_file(cls, file_name):
        """reference_kmer_index.py:140-160."""
        try:
            data = np.load(file_name + ".npz")
        except FileNotFoundError:
            data = np.load(file_name)
        return cls(data["ref_position_to_index"] if "ref_position_to_index" in data else None, data["kmers"],
                   data["ref_positions"] if "ref_positions" in data else None, data["nodes"] if "nodes" in data else None)
