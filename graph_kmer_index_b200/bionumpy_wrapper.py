"""Drop-in for graph_kmer_index/bionumpy_wrapper.py:4-10: the k-mer hashes of a numeric sequence,
``out[i] = sum_j seq[i+j] * 4**j`` -- the same numbers as ``ReadKmers.get_kmers_from_read_dynamic`` (the reference's
tests/test_structural_variants.py:44-49 pins them to ``sequence_to_kmer_hash``).  A one-read call of K1 (gki_hash_reads)."""
import numpy as np

from .read_kmers import hash_read_matrix

_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def bionumpy_hash(numeric_sequence, k):
    codes = np.asarray(numeric_sequence)
    if len(codes) < k:
        return np.zeros(0, dtype=np.uint64)
    assert codes.min() >= 0 and codes.max() <= 3, "numeric sequence must hold base codes 0..3"
    fwd, _ = hash_read_matrix(_LETTERS[codes.astype(np.int64)][None, :], k, forward=True, reverse=False)
    return fwd[0]
