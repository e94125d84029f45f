"""Drop-in for graph_kmer_index/kmer_hashing.py (+ the scalar helpers of snp_kmer_finder.py:14-26 and the
letter<->numeric encoders of flat_kmers.py:134-154).  All array arithmetic runs in libgki.so (csrc/hash.cu)."""
import numpy as np

from . import _lib


def power_array(k):
    """kmer_hashing.py:4-5 (a k-element constant, not a data path)."""
    return np.power(4, np.arange(k - 1, -1, -1)).astype(np.uint64)


def reverse_power_array(k):
    """kmer_hashing.py:8-9."""
    return np.power(4, np.arange(k)).astype(np.uint64)


def _u64(hashes):
    return np.ascontiguousarray(np.asarray(hashes).astype(np.uint64, copy=False))


def kmer_hash_to_reverse_complement_hash(hash, k):
    """kmer_hashing.py:12-13."""
    return kmer_hashes_to_reverse_complement_hash(np.array([hash]), k)[0]


def kmer_hashes_to_reverse_complement_hash_chunked(hashes, k, chunk_size=1000000):
    """kmer_hashing.py:16-22.  The device path needs no (N,k) temporary, so chunking only bounds staging."""
    hashes = np.asarray(hashes)
    if len(hashes) == 0:
        return np.concatenate([])  # the reference raises ValueError on empty input too
    return np.concatenate([kmer_hashes_to_reverse_complement_hash(hashes[i:i + chunk_size], k)
                           for i in range(0, len(hashes), chunk_size)])


def kmer_hashes_to_reverse_complement_hash(hashes, k):
    """kmer_hashing.py:24-28."""
    assert k <= 31
    h = _u64(hashes)
    out = np.empty(len(h), dtype=np.uint64)
    _lib.call("gki_revcomp_hashes", _lib.ptr(h), len(h), k, _lib.ptr(out), _lib.current_stream())
    return out


def kmer_hashes_to_complement_hashes(hashes, k):
    """kmer_hashing.py:31-36."""
    assert k <= 31
    h = _u64(hashes)
    out = np.empty(len(h), dtype=np.uint64)
    _lib.call("gki_complement_hashes", _lib.ptr(h), len(h), k, _lib.ptr(out), _lib.current_stream())
    return out


def kmer_hashes_to_bases(hashes, k):
    """kmer_hashing.py:53-65: (N, k) uint64, column j = base j of the sequence."""
    h = _u64(hashes)
    out = np.empty((len(h), k), dtype=np.uint64)
    _lib.call("gki_hashes_to_bases", _lib.ptr(h), len(h), k, _lib.ptr(out), _lib.current_stream())
    return out


def kmer_hashes_to_complement_bases(hashes, k):
    """kmer_hashing.py:40-49: bases of the complemented k-mer (0<->3, 1<->2)."""
    return kmer_hashes_to_bases(kmer_hashes_to_complement_hashes(hashes, k), k)


# ---- flat_kmers.py:134-154 ---------------------------------------------------------------------------
def _ascii_bytes(sequence):
    """str / bytes / uint8 array -> uint8 ASCII array.  A numpy array of single characters is NOT lower-cased
    by the reference (flat_kmers.py:135-136 only lowers non-arrays), so upper-case letters in such an
    array encode to 0; that quirk is reproduced by mapping them to a non-ACGT byte."""
    if isinstance(sequence, str):
        return np.frombuffer(sequence.encode("latin-1", "replace"), dtype=np.uint8)
    if isinstance(sequence, (bytes, bytearray)):
        return np.frombuffer(bytes(sequence), dtype=np.uint8)
    a = np.asarray(sequence)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a)
    if a.dtype.kind in "US":
        b = np.frombuffer("".join(str(c)[:1] or "-" for c in a.astype(str)).encode("latin-1", "replace"), dtype=np.uint8).copy()
        b[(b >= ord("A")) & (b <= ord("Z"))] = ord("-")
        return b
    raise TypeError("sequence must be str, bytes or an array of characters / ASCII bytes")


def letter_sequence_to_numeric(sequence):
    """flat_kmers.py:134-145: a/n/m/other -> 0, c -> 1, g -> 2, t -> 3 as uint64."""
    b = _ascii_bytes(sequence)
    out = np.empty(len(b), dtype=np.uint64)
    _lib.call("gki_encode_bases", _lib.ptr(b), len(b), _lib.ptr(out), _lib.current_stream())
    return out


_LETTERS = np.array(["a", "c", "g", "t"], dtype=object)


def numeric_to_letter_sequence(sequence):
    """flat_kmers.py:147-154 (a 4-entry table lookup; values outside 0..3 give 0 like the reference)."""
    seq = np.asarray(sequence)
    out = np.zeros(seq.shape, dtype=object)
    for v in range(4):
        out[seq == v] = _LETTERS[v]
    return out


# ---- snp_kmer_finder.py:14-26 ------------------------------------------------------------------------
def _hash_one(ascii_bytes):
    k = len(ascii_bytes)
    if k == 0:
        return 0
    if k > 31:
        raise ValueError("k-mers longer than 31 bases do not fit the 2-bit uint64 hash (reference asserts k <= 31)")
    out = np.empty(1, dtype=np.uint64)
    b = np.ascontiguousarray(ascii_bytes)
    _lib.call("gki_hash_reads", _lib.ptr(b), 1, k, k, k, _lib.ptr(out), None, _lib.current_stream())
    return int(out[0])


def sequence_to_kmer_hash(sequence):
    """snp_kmer_finder.py:19-20."""
    return _hash_one(_ascii_bytes(sequence))


def kmer_to_hash_fast(kmer, k):
    """snp_kmer_finder.py:23-26: numeric (0..3) uint64 array -> python int."""
    assert kmer.dtype == np.uint64
    return _hash_one(np.frombuffer(b"ACGT", dtype=np.uint8)[np.asarray(kmer[:k]).astype(np.int64) & 3])


def kmer_hash_to_sequence(hash, k):
    """snp_kmer_finder.py:14-16."""
    bases = kmer_hashes_to_bases(np.array([hash], dtype=np.uint64), k)[0]
    return "".join(numeric_to_letter_sequence(bases))
