"""Drop-in for graph_kmer_index/multi_value_hashtable.py:5-25: keys (repeats allowed) -> row numbers -> a dict of value columns.

The reference keeps an ``npstructures.HashTable`` from key to row number; here that table is the index of
collision_free_kmer_index.py built on the device (gki_index_build over (key, row): the stable bucket order keeps the rows of a
key in input order) and probed with gki_lookup_entries.  The value columns stay on the host untouched, any dtype."""
import numpy as np


class MultiValueHashTable:
    def __init__(self, hash_table, values):
        self._hash_table = hash_table          # CollisionFreeKmerIndex whose `_nodes` column holds row numbers
        self._values = values

    def get_unique_keys(self):
        return np.unique(self._hash_table._kmers.ravel())

    def get_all_keys(self):
        return self._hash_table._kmers.ravel()

    @classmethod
    def from_keys_and_values(cls, keys, values: dict, mod=None):
        from .collision_free_kmer_index import CollisionFreeKmerIndex, build_index_arrays
        keys = np.asarray(keys)
        assert len(keys) < 2 ** 32
        for column in values.values():
            assert len(column) == len(keys)
        modulo = int(mod) if mod is not None else 2 * len(keys) + 1
        h2i, nk, kmers, rows, _, _, freq = build_index_arrays(keys, np.arange(len(keys), dtype=np.uint32), None, None, modulo, True)
        return cls(CollisionFreeKmerIndex(h2i, nk, rows, np.zeros(1, dtype=np.uint64), kmers, modulo, _frequencies=freq,
                                          _allele_frequencies=np.zeros(1, dtype=np.float32)), values)

    def rows(self, keys):
        """Row numbers of every entry whose key is in `keys` (scalar or array): ordered by query, then by row."""
        entries, _ = self._hash_table.device_index().lookup_entries(keys)
        return self._hash_table._nodes[entries].astype(np.int64)

    def __getitem__(self, keys):
        indexes = self.rows(keys)
        return {name: np.asarray(value)[indexes] for name, value in self._values.items()}
