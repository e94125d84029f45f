"""ctypes front-end of oracle/gki_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgki_oracle.so")
_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    src = os.path.join(_HERE, "gki_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgki_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_lookup_hits.restype = ctypes.c_int64
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the bench's reference arm asks for the cores it may use instead"""
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def hash_reads(reads_u8, k, rc=True):
    reads_u8 = np.ascontiguousarray(reads_u8, dtype=np.uint8)
    n, L = reads_u8.shape
    nk = max(L - k + 1, 0)
    fwd = np.zeros((n, nk), dtype=np.uint64)
    rcv = np.zeros((n, nk), dtype=np.uint64)
    lib().orc_hash_reads(_p(reads_u8, _u8p), ctypes.c_int64(n), ctypes.c_int32(L), ctypes.c_int64(L),
                         ctypes.c_int32(k), _p(fwd, _u64p), _p(rcv, _u64p) if rc else None)
    return fwd, rcv


def revcomp_hashes(h, k):
    h = np.ascontiguousarray(h, dtype=np.uint64)
    out = np.empty_like(h)
    lib().orc_revcomp_hashes(_p(h, _u64p), ctypes.c_int64(len(h)), ctypes.c_int32(k), _p(out, _u64p))
    return out


def complement_hashes(h, k):
    h = np.ascontiguousarray(h, dtype=np.uint64)
    out = np.empty_like(h)
    lib().orc_complement_hashes(_p(h, _u64p), ctypes.c_int64(len(h)), ctypes.c_int32(k), _p(out, _u64p))
    return out


def build_index(kmers, nodes, ref_offsets, af, modulo, skip_frequencies=False):
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    ref_offsets = np.ascontiguousarray(ref_offsets, dtype=np.uint64)
    af = np.ascontiguousarray(af, dtype=np.float32)
    n = len(kmers)
    out = dict(_hashes_to_index=np.empty(modulo, np.int32), _n_kmers=np.empty(modulo, np.uint32),
               _kmers=np.empty(n, np.uint64), _nodes=np.empty(n, np.uint32), _ref_offsets=np.empty(n, np.uint64),
               _allele_frequencies=np.empty(n, np.float32), _frequencies=np.empty(n, np.uint16), _modulo=int(modulo))
    rc = lib().orc_build(_p(kmers, _u64p), _p(nodes, _u32p), _p(ref_offsets, _u64p), _p(af, _f32p),
                         ctypes.c_int64(n), ctypes.c_uint64(modulo), ctypes.c_int32(int(skip_frequencies)),
                         _p(out["_hashes_to_index"], _i32p), _p(out["_n_kmers"], _u32p), _p(out["_kmers"], _u64p),
                         _p(out["_nodes"], _u32p), _p(out["_ref_offsets"], _u64p),
                         _p(out["_allele_frequencies"], _f32p), _p(out["_frequencies"], _u16p))
    assert rc == 0
    return out


def build_index_kmers_nodes(kmers, nodes, modulo):
    """k-mer / node columns and both tables of build_index (same order), on all host threads (bench CPU arm at 1 B entries)"""
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    n = len(kmers)
    out = dict(_hashes_to_index=np.empty(modulo, np.int32), _n_kmers=np.empty(modulo, np.uint32),
               _kmers=np.empty(n, np.uint64), _nodes=np.empty(n, np.uint32), _modulo=int(modulo))
    rc = lib().orc_build_kmers_nodes(_p(kmers, _u64p), _p(nodes, _u32p), ctypes.c_int64(n), ctypes.c_uint64(modulo),
                                     _p(out["_hashes_to_index"], _i32p), _p(out["_n_kmers"], _u32p), _p(out["_kmers"], _u64p),
                                     _p(out["_nodes"], _u32p))
    assert rc == 0
    return out


def _index_args(index):
    h2i = np.ascontiguousarray(index["_hashes_to_index"], dtype=np.int32)
    nk = np.ascontiguousarray(index["_n_kmers"], dtype=np.uint32)
    km = np.ascontiguousarray(index["_kmers"], dtype=np.uint64)
    return h2i, nk, km


def count_kmers(index, queries, entry_counts=None):
    h2i, nk, km = _index_args(index)
    q = np.ascontiguousarray(queries, dtype=np.uint64)
    if entry_counts is None:
        entry_counts = np.zeros(len(km), dtype=np.uint32)
    lib().orc_count_kmers(_p(h2i, _i32p), _p(nk, _u32p), _p(km, _u64p), ctypes.c_uint64(index["_modulo"]),
                          _p(q, _u64p), ctypes.c_int64(len(q)), _p(entry_counts, _u32p))
    return entry_counts


def count_reads(index, reads_u8, k, both_strands=True, entry_counts=None, prepared=None):
    h2i, nk, km = prepared if prepared is not None else _index_args(index)
    reads_u8 = np.ascontiguousarray(reads_u8, dtype=np.uint8)
    n, L = reads_u8.shape
    if entry_counts is None:
        entry_counts = np.zeros(len(km), dtype=np.uint32)
    lib().orc_count_reads(_p(h2i, _i32p), _p(nk, _u32p), _p(km, _u64p), ctypes.c_uint64(index["_modulo"]),
                          _p(reads_u8, _u8p), ctypes.c_int64(n), ctypes.c_int32(L), ctypes.c_int64(L),
                          ctypes.c_int32(k), ctypes.c_int32(int(both_strands)), _p(entry_counts, _u32p))
    return entry_counts


def node_counts_from_entry_counts(index, entry_counts, min_nodes=0, parallel=False, size=None):
    nodes = np.ascontiguousarray(index["_nodes"], dtype=np.uint32)
    if size is None:
        size = max(int(min_nodes), int(nodes.max()) + 1 if len(nodes) else 0)
    out = np.zeros(size, dtype=np.float64)
    fn = lib().orc_node_counts_parallel if parallel else lib().orc_node_counts
    fn(_p(nodes, _u32p), _p(entry_counts, _u32p), ctypes.c_int64(len(nodes)), _p(out, _f64p))
    return out


def read_node_counts(index, reads_u8, k, min_nodes=0, both_strands=True):
    return node_counts_from_entry_counts(index, count_reads(index, reads_u8, k, both_strands), min_nodes)


def lookup_hits(index, queries, skip_bucket0=True, max_bucket=10000, max_frequency=20):
    h2i, nk, km = _index_args(index)
    nodes = np.ascontiguousarray(index["_nodes"], dtype=np.uint32)
    ref = np.ascontiguousarray(index["_ref_offsets"], dtype=np.uint64)
    freq = np.ascontiguousarray(index["_frequencies"], dtype=np.uint16)
    af = np.ascontiguousarray(index["_allele_frequencies"], dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.uint64)
    args = [_p(h2i, _i32p), _p(nk, _u32p), _p(km, _u64p), _p(nodes, _u32p), _p(ref, _u64p), _p(freq, _u16p),
            _p(af, _f32p), ctypes.c_uint64(index["_modulo"]), _p(q, _u64p), ctypes.c_int64(len(q)),
            ctypes.c_int32(int(skip_bucket0)), ctypes.c_int64(-1 if max_bucket is None else max_bucket),
            ctypes.c_int32(-1 if max_frequency is None else max_frequency)]
    n_hits = lib().orc_lookup_hits(*args, None, ctypes.c_int64(0))
    out = np.zeros((5, n_hits), dtype=np.uint64)
    lib().orc_lookup_hits(*args, _p(out, _u64p), ctypes.c_int64(n_hits))
    return out
