"""Drop-in for graph_kmer_index/read_kmers.py: read k-mer hashing on the device (csrc/hash.cu).

The reference hashes one read at a time with np.convolve (read_kmers.py:67-70) and gets the reverse strand by
reverse-complementing the read string (read_kmers.py:21-26).  Here whole batches go to the GPU at once:
``hash_read_matrix`` for equal-length reads, ``hash_ragged_reads`` for any lengths; the reference-shaped entry
points are thin views over those."""
import itertools
import logging

import numpy as np

from . import _lib
from .kmer_hashing import _ascii_bytes, kmer_to_hash_fast, letter_sequence_to_numeric, power_array


def hash_read_matrix(reads, k, forward=True, reverse=True):
    """(n_reads, L) uint8 ASCII (numpy, or a torch uint8 tensor on host/device) -> (fwd, rc), each
    (n_reads, L-k+1) uint64 of the same kind as the input; rc[r] are the hashes of the reverse-complemented
    read r in its own left-to-right order (what read_kmers.py:24 yields)."""
    n, L = reads.shape
    nk = max(L - k + 1, 0)
    if isinstance(reads, np.ndarray):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        mk = lambda: np.empty((n, nk), dtype=np.uint64)
    else:
        import torch
        reads = reads.contiguous()
        mk = lambda: torch.empty((n, nk), dtype=torch.uint64, device=reads.device)
    fwd = mk() if forward else None
    rc = mk() if reverse else None
    _lib.call("gki_hash_reads", _lib.ptr(reads), n, L, L, k, _lib.ptr(fwd), _lib.ptr(rc), _lib.current_stream())
    return fwd, rc


def hash_ragged_reads(reads, k, reverse=True):
    """list of str/bytes reads of any length -> (fwd, rc, out_offsets); read r owns
    fwd[out_offsets[r]:out_offsets[r+1]].  Reads shorter than k own nothing."""
    chunks = [_ascii_bytes(r) for r in reads]
    lens = np.array([len(c) for c in chunks], dtype=np.int64)
    offsets = np.zeros(len(chunks) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    out_offsets = np.zeros(len(chunks) + 1, dtype=np.int64)
    np.cumsum(np.maximum(lens - k + 1, 0), out=out_offsets[1:])
    seq = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    fwd = np.empty(int(out_offsets[-1]), dtype=np.uint64)
    rc = np.empty(int(out_offsets[-1]), dtype=np.uint64) if reverse else None
    if len(chunks):
        _lib.call("gki_hash_reads_ragged", _lib.ptr(seq), _lib.ptr(offsets), _lib.ptr(out_offsets), len(chunks), k,
                  _lib.ptr(fwd), _lib.ptr(rc), _lib.current_stream())
    return fwd, rc, out_offsets


def pack_reads(reads, n_threads=0, force_scalar=False):
    """(n_reads, L) uint8 ASCII host array -> (packed, dirty_index): the rows made only of ACGTacgt as 2-bit codes,
    ceil(L/32) uint64 words each (base i at bits 2*(i%32) of word i/32, a0 c1 g2 t3), in input order, and the indices of
    the other rows.  Runs on the host cores (csrc/ingest.cpp: AVX-512 + BMI2 when the CPU has them); the packed rows feed
    ``DeviceIndex.count_packed_reads`` at a quarter of the PCIe bytes of the ASCII form."""
    import ctypes
    reads = np.asarray(reads)
    assert reads.dtype == np.uint8 and reads.ndim == 2 and (reads.shape[0] < 2 or reads.strides[1] == 1)
    n, L = reads.shape
    packed = np.empty((n, (L + 31) // 32), dtype=np.uint64)
    dirty = np.empty(n, dtype=np.int64)
    n_clean, n_dirty = ctypes.c_int64(), ctypes.c_int64()
    _lib.call("gki_pack_reads", reads.ctypes.data, n, L, reads.strides[0] if n > 1 else L, _lib.ptr(packed), _lib.ptr(dirty), n,
              ctypes.byref(n_clean), ctypes.byref(n_dirty), int(n_threads), 1 if force_scalar else 0)
    return packed[:n_clean.value], dirty[:n_dirty.value]


class FastxFile:
    """A FASTA / FASTQ file mapped by libgki with its sequence lines located by the host threads (csrc/ingest.cpp): FASTA -- every
    line not starting with '>' (what read_kmers.py:16-18 treats as a read), FASTQ -- the second line of every record."""

    def __init__(self, path):
        import ctypes
        self.path = str(path)
        self.handle = ctypes.c_void_p()
        n, max_len, fmt = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
        _lib.call("gki_fastx_open", self.path.encode(), ctypes.byref(self.handle), ctypes.byref(n), ctypes.byref(max_len), ctypes.byref(fmt))
        self.n_reads, self.max_len, self.format = n.value, max_len.value, ("fasta", "fastq")[fmt.value]

    def lines(self):
        """(offsets int64, lengths int32) of the sequence lines, blanks stripped."""
        offsets, lengths = np.empty(self.n_reads, dtype=np.int64), np.empty(self.n_reads, dtype=np.int32)
        _lib.call("gki_fastx_lines", self.handle, _lib.ptr(offsets), _lib.ptr(lengths))
        return offsets, lengths

    def close(self):
        if self.handle:
            _lib.load().gki_fastx_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _k_from_power_vector(power_vector):
    k = len(power_vector)
    if not np.array_equal(np.asarray(power_vector, dtype=np.uint64), power_array(k)):
        raise ValueError("power_vector must be power_array(k): the device hash is the 2-bit packing it defines")
    return k


class ReadKmers:
    """read_kmers.py:9-89."""

    def __init__(self, kmers):
        self.kmers = kmers
        self._power_vector = None

    @classmethod
    def from_fasta_file(cls, fasta_file_name, k, small_k=None, smallest_k=8):
        """read_kmers.py:14-49: forward k-mers of every read, then the k-mers of every reverse-complemented
        read (or, with small_k, per-read triples of chained forward+reverse streams for k, small_k, smallest_k)."""
        with open(fasta_file_name) as f:
            lines = [l.strip() for l in f.readlines() if not l.startswith(">")]
        logging.info("Number of lines: %d" % len(lines))
        per_k = {}
        for kk in ([k] if small_k is None else [k, small_k, smallest_k]):
            fwd, rc, off = hash_ragged_reads(lines, kk)
            per_k[kk] = ([fwd[off[i]:off[i + 1]] for i in range(len(lines))], [rc[off[i]:off[i + 1]] for i in range(len(lines))])
        if small_k is None:
            kmers = itertools.chain(iter(per_k[k][0]), iter(per_k[k][1]))
        else:
            kmers = zip(*[(itertools.chain(f, r) for f, r in zip(*per_k[kk])) for kk in (k, small_k, smallest_k)])
        return cls(kmers)

    @classmethod
    def from_list_of_string_kmers(cls, string_kmers):
        """read_kmers.py:51-57."""
        return cls([[kmer_to_hash_fast(letter_sequence_to_numeric(k), len(k)) for k in read_kmers] for read_kmers in string_kmers])

    @staticmethod
    def get_kmers_from_read(read, k):
        """read_kmers.py:59-65 (including its range(len(read) - k): the last k-mer is not produced)."""
        fwd, _, _ = hash_ragged_reads([read], k, reverse=False)
        return [int(h) for h in fwd[:max(len(read) - k, 0)]]

    @staticmethod
    def get_kmers_from_read_dynamic(read, power_vector):
        """read_kmers.py:67-70: all k-mer hashes of one read (uint64).  A read shorter than k gives an empty
        array (the reference's np.convolve swaps its arguments there and returns k-len+1 meaningless values)."""
        k = _k_from_power_vector(power_vector)
        fwd, _, _ = hash_ragged_reads([read], k, reverse=False)
        return fwd

    def __iter__(self):
        return self.kmers.__iter__()

    def __next__(self):
        return self.kmers.__next__()
