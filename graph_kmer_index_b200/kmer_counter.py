"""Drop-in for graph_kmer_index/kmer_counter.py: KmerCounter.from_kmers (kmer_counter.py:33-43) is `np.unique(kmers,
return_counts=True)` into an npstructures HashTable; here the k-mers are indexed and counted on the device (K2 + the count
table of csrc/count.cu) and `counter[kmers]` is a batched lookup."""
import numpy as np

from .collision_free_kmer_index import DeviceIndex, build_index_arrays


def choose_modulo(n_elements):
    """kmer_counter.py:9-15."""
    if n_elements < 1000000:
        return 2000003
    elif n_elements < 10000000:
        return 19999999
    else:
        return 200000003


class DeviceKmerCounts:
    """What `KmerCounter.counter` exposes of the reference's HashTable: `counter[kmer]` -> one-element array with the number of
    occurrences (empty array for a k-mer that was never seen), `counter[array]` -> counts, 0 where unseen."""

    def __init__(self, device_index):
        self._device = device_index

    def __getitem__(self, keys):
        if np.ndim(keys) == 0:
            c = self._device.query_counts(np.array([int(keys)], dtype=np.uint64))
            return c.astype(np.int64) if c[0] else np.zeros(0, dtype=np.int64)
        return self._device.query_counts(np.asarray(keys).astype(np.uint64)).astype(np.int64)


class KmerCounter:
    """kmer_counter.py:19-87."""

    def __init__(self, counter):
        self.counter = counter

    @classmethod
    def from_flat_kmersv2(cls, flat, modulo, subsample_ratio=1):
        """kmer_counter.py:23-31."""
        return cls.from_kmers(flat._hashes[::subsample_ratio], modulo)

    @classmethod
    def from_kmers(cls, kmers, modulo):
        """kmer_counter.py:33-43: occurrences of every distinct k-mer."""
        kmers = np.ascontiguousarray(np.asarray(kmers).astype(np.uint64))
        if modulo == 0:
            modulo = choose_modulo(len(kmers))      # the reference sizes it by the distinct count; only a table size
        h2i, n_kmers, s_kmers, s_nodes, _, _, _ = build_index_arrays(kmers, np.zeros(len(kmers), dtype=np.uint32), None, None, modulo, True)
        device = DeviceIndex(h2i, n_kmers, s_kmers, s_nodes, modulo)
        device.prepare_counting(0)                  # raw keys: a k-mer and its reverse complement are different keys here
        device.count_kmers(kmers)
        return cls(DeviceKmerCounts(device))

    from_flat_kmers = classmethod(lambda cls, flat, modulo, chunk_size=50000000: cls.from_kmers(flat._hashes, modulo))   # kmer_counter.py:45-70

    def get_frequency(self, kmer):
        """kmer_counter.py:72-74."""
        return self.counter[int(kmer)]

    def score_kmers(self, kmers):
        """kmer_counter.py:76-84."""
        hits = [self.counter[int(k)] for k in kmers]
        hits = [h[0] for h in hits if len(h) > 0]
        if len(hits) == 0:
            return 1
        return -np.max(hits)
