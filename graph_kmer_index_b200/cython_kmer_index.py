"""Drop-in for graph_kmer_index/cython_kmer_index.pyx (the reference's only native probe loop), on the device."""
import numpy as np

from .collision_free_kmer_index import DeviceIndex


class CythonKmerIndex:
    """cython_kmer_index.pyx:21-109.  get(kmers) -> (5, n_hits) uint64 rows
    [node, ref_offset, query index, frequency, uint64(1000 * allele_frequency)], with the .pyx's gates:
    bucket 0 skipped (pyx:59-60), buckets > 10000 entries skipped (pyx:62-63), frequency > 20 skipped (pyx:70-71)."""

    def __init__(self, index):
        self._device = index.device_index() if hasattr(index, "device_index") else DeviceIndex.from_index(index)

    def get(self, kmers):
        return self._device.lookup_hits(np.asarray(kmers, dtype=np.uint64), skip_bucket0=True, max_bucket=10000, max_frequency=20)
