"""Drop-in for graph_kmer_index/collision_free_kmer_index.py.

Same classes, attribute names (host numpy arrays), npz layout and call signatures as the reference; the
arithmetic runs in libgki.so: construction in csrc/build.cu (K2), probing / counting in csrc/index.cu (K3).
A device-resident copy of the index (``DeviceIndex``) is created lazily the first time a lookup is made.
"""
import ctypes
import gc
import logging
import weakref

import numpy as np

from . import _lib
from .flat_kmers import FlatKmers
from .kmer_hashing import (kmer_hash_to_sequence, kmer_hashes_to_complement_hashes,  # noqa: F401
                           kmer_hashes_to_reverse_complement_hash, sequence_to_kmer_hash)

DEFAULT_MODULO = 452930477


def _c(a, dtype=None):
    a = np.asarray(a)
    if dtype is not None and a.dtype != dtype:
        a = a.astype(dtype)
    return np.ascontiguousarray(a)


def _queries(kmers):
    kmers = np.atleast_1d(np.asarray(kmers))
    if kmers.dtype == np.int64 and kmers.flags.c_contiguous:
        return kmers.view(np.uint64)          # same bits as astype(uint64), without the copy
    return _c(kmers, np.uint64)


class DeviceIndex:
    """Owner of a gki_index_t* (include/gki.h): the index laid out in HBM for probing."""

    def __init__(self, hashes_to_index, n_kmers, kmers, nodes, modulo, ref_offsets=None, frequencies=None,
                 allele_frequencies=None, flags=0):
        # host numpy arrays or device torch tensors; dtypes must already be int32/uint32/uint64/uint32/...
        self._keep = (hashes_to_index, n_kmers, kmers, nodes, ref_offsets, frequencies, allele_frequencies)
        self.handle = ctypes.c_void_p()
        n = int(kmers.shape[0])
        _lib.call("gki_index_create", _lib.ptr(hashes_to_index), _lib.ptr(n_kmers), _lib.ptr(kmers), _lib.ptr(nodes),
                  _lib.ptr(ref_offsets), _lib.ptr(frequencies), _lib.ptr(allele_frequencies), n, int(modulo), flags,
                  ctypes.byref(self.handle), _lib.current_stream())
        self._keep = None
        info = self.info()
        self.n, self.modulo, self.max_node = info["n"], info["modulo"], info["max_node"]

    @classmethod
    def from_index(cls, index, flags=0):
        """From a CollisionFreeKmerIndex-like object holding host arrays."""
        freq = index._frequencies
        freq = _c(freq, np.uint16) if isinstance(freq, np.ndarray) and freq.shape == np.shape(index._kmers) else None
        ref = index._ref_offsets
        ref = _view64(_c(ref)) if isinstance(ref, np.ndarray) and ref.shape == np.shape(index._kmers) else None
        af = index._allele_frequencies
        af = _c(af, np.float32) if isinstance(af, np.ndarray) and af.shape == np.shape(index._kmers) else None
        return cls(_c(index._hashes_to_index, np.int32), _c(index._n_kmers, np.uint32), _c(index._kmers, np.uint64),
                   _c(index._nodes, np.uint32), int(index._modulo), ref, freq, af, flags)

    def info(self):
        n, mod, mx, nbytes, bm, ne = ctypes.c_int64(), ctypes.c_uint64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int64()
        _lib.call("gki_index_info", self.handle, ctypes.byref(n), ctypes.byref(mod), ctypes.byref(mx), ctypes.byref(nbytes), ctypes.byref(bm),
                  ctypes.byref(ne))
        return dict(n=n.value, modulo=mod.value, max_node=mx.value, device_bytes=nbytes.value, has_filter=bool(bm.value),
                    nonempty_buckets=ne.value)

    def prepare_counting(self, k=0):
        """Build the counting structure (L2 Bloom filter + table over the distinct k-mers) now instead of inside the
        first counting call.  k = k-mer length of the reads to be counted (0 if unknown); results never depend on it."""
        _lib.call("gki_prepare_counting", self.handle, int(k), _lib.current_stream())

    def close(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                _lib.load().gki_index_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except (TypeError, AttributeError):      # interpreter shutdown: modules already torn down
            pass

    __del__ = close

    # ---- counting (cfki:30-40) ----
    def reset_counts(self):
        _lib.call("gki_reset_counts", self.handle, _lib.current_stream())

    def count_kmers(self, kmers):
        """kmers: uint64 numpy array or device tensor."""
        _lib.call("gki_count_kmers", self.handle, _lib.ptr(kmers), int(kmers.shape[0]), _lib.current_stream())

    def count_reads(self, reads, k, both_strands=True):
        """reads: (n_reads, L) uint8 ASCII numpy array (host) or torch tensor (host / device); rows may be strided."""
        n, L = reads.shape
        if isinstance(reads, np.ndarray):
            assert reads.dtype == np.uint8 and reads.strides[1] == 1
            stride = reads.strides[0] if n > 1 else L
            p = reads.ctypes.data
        else:
            assert reads.stride(1) == 1
            stride = reads.stride(0) if n > 1 else L
            p = reads.data_ptr()
        _lib.call("gki_count_reads", self.handle, p, n, L, stride, k, int(both_strands), _lib.current_stream())

    def count_packed_reads(self, packed, read_len, k, both_strands=True):
        """packed: (n_reads, ceil(read_len/32)) uint64 rows from read_kmers.pack_reads (numpy, or a torch tensor on host / device;
        torch has no uint64 arithmetic, an int64 tensor with the same bits is fine)."""
        assert packed.shape[1] == (read_len + 31) // 32
        _lib.call("gki_count_packed_reads", self.handle, _lib.ptr(packed), int(packed.shape[0]), int(read_len), k, int(both_strands),
                  _lib.current_stream())

    def count_fastx(self, fastx, k, both_strands=True):
        """Count every sequence line of a FASTA / FASTQ file (path or read_kmers.FastxFile); returns the number of k-mers looked up."""
        import ctypes
        from .read_kmers import FastxFile
        own = not isinstance(fastx, FastxFile)
        f = FastxFile(fastx) if own else fastx
        try:
            n_kmers = ctypes.c_int64()
            _lib.call("gki_count_fastx", self.handle, f.handle, int(k), int(both_strands), ctypes.byref(n_kmers), _lib.current_stream())
            return n_kmers.value
        finally:
            if own:
                f.close()

    def node_counts(self, min_nodes=0, out=None, wrap_uint16=False):
        n_out = max(int(min_nodes), self.max_node + 1)
        if out is None:
            out = np.empty(n_out, dtype=np.float64)
        assert out.shape[0] >= n_out
        _lib.call("gki_node_counts", self.handle, _lib.ptr(out), int(out.shape[0]),
                  _lib.GKI_COUNTS_WRAP_UINT16 if wrap_uint16 else 0, _lib.current_stream())
        return out

    def entry_counts(self):
        out = np.empty(self.n, dtype=np.uint32)
        _lib.call("gki_entry_counts", self.handle, _lib.ptr(out), _lib.current_stream())
        return out

    def query_counts(self, kmers):
        q = _queries(kmers)
        out = np.empty(len(q), dtype=np.uint32)
        _lib.call("gki_query_counts", self.handle, _lib.ptr(q), len(q), _lib.ptr(out), _lib.current_stream())
        return out

    # ---- lookups ----
    def map_kmers(self, kmers, n_nodes, skip_bucket0=False, max_frequency=None):
        q = _queries(kmers)
        out = np.zeros(int(n_nodes), dtype=np.uint64)
        _lib.call("gki_map_kmers", self.handle, _lib.ptr(q), len(q), _lib.ptr(out), int(n_nodes),
                  _lib.GKI_PROBE_SKIP_BUCKET0 if skip_bucket0 else 0, -1 if max_frequency is None else int(max_frequency),
                  _lib.current_stream())
        return out

    def has_kmers(self, kmers, skip_bucket0=False):
        q = _queries(kmers)
        out = np.empty(len(q), dtype=np.uint8)
        _lib.call("gki_has_kmers", self.handle, _lib.ptr(q), len(q), _lib.ptr(out),
                  _lib.GKI_PROBE_SKIP_BUCKET0 if skip_bucket0 else 0, _lib.current_stream())
        return out.astype(bool)

    def _gates(self, skip_bucket0, max_bucket, max_frequency):
        return (_lib.GKI_PROBE_SKIP_BUCKET0 if skip_bucket0 else 0, -1 if max_bucket is None else int(max_bucket),
                -1 if max_frequency is None else int(max_frequency))

    def lookup_entries(self, kmers, skip_bucket0=False, max_bucket=None, max_frequency=None):
        """-> (entry positions int64, query indices int64), ordered by query then entry."""
        q = _queries(kmers)
        flags, mb, mf = self._gates(skip_bucket0, max_bucket, max_frequency)
        n_hits = ctypes.c_int64()
        _lib.call("gki_lookup_entries", self.handle, _lib.ptr(q), len(q), flags, mb, mf, None, None, 0, ctypes.byref(n_hits), _lib.current_stream())
        entries = np.empty(n_hits.value, dtype=np.int64)
        qidx = np.empty(n_hits.value, dtype=np.int64)
        if n_hits.value:
            _lib.call("gki_lookup_entries", self.handle, _lib.ptr(q), len(q), flags, mb, mf, _lib.ptr(entries), _lib.ptr(qidx),
                      n_hits.value, ctypes.byref(n_hits), _lib.current_stream())
        return entries, qidx

    def lookup_hits(self, kmers, skip_bucket0=True, max_bucket=10000, max_frequency=20):
        """cython_kmer_index.pyx:47-109 -> (5, n_hits) uint64."""
        q = _queries(kmers)
        flags, mb, mf = self._gates(skip_bucket0, max_bucket, max_frequency)
        n_hits = ctypes.c_int64()
        _lib.call("gki_lookup_hits", self.handle, _lib.ptr(q), len(q), flags, mb, mf, None, 0, ctypes.byref(n_hits), _lib.current_stream())
        out = np.zeros((5, n_hits.value), dtype=np.uint64)
        if n_hits.value:
            _lib.call("gki_lookup_hits", self.handle, _lib.ptr(q), len(q), flags, mb, mf, _lib.ptr(out), n_hits.value,
                      ctypes.byref(n_hits), _lib.current_stream())
        return out


def _view64(a):
    """8-byte view used to move a column bit-exactly / to compare ref offsets for set_frequencies."""
    a = np.ascontiguousarray(a)
    if a.dtype.itemsize == 8:
        return a.view(np.uint64)
    if a.dtype.kind == "f":
        return np.ascontiguousarray(a.astype(np.float64)).view(np.uint64)
    return np.ascontiguousarray(a.astype(np.int64)).view(np.uint64)


def build_index_arrays(kmers, nodes, ref_offsets, allele_frequencies, modulo, skip_frequencies):
    """collision_free_kmer_index.py:433-457 + :267-293 on the device.  Inputs are host arrays of any dtype the
    reference accepts; every payload column keeps its dtype (moved bit-exactly through the stable bucket sort)."""
    kmers = np.asarray(kmers)
    n = len(kmers)
    if n == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # what cfki:455 raises on empty input
    modulo = int(modulo)
    k64 = _c(kmers, np.uint64) if kmers.dtype.itemsize != 8 else np.ascontiguousarray(kmers).view(np.uint64)
    cols = {"nodes": (nodes, 4), "ref": (ref_offsets, 8), "af": (allele_frequencies, 4)}
    direct, generic = {}, {}
    for name, (col, size) in cols.items():
        if col is None:
            continue
        col = np.ascontiguousarray(np.asarray(col))
        assert len(col) == n
        (direct if col.dtype.itemsize == size else generic)[name] = col
    h2i = np.empty(modulo, dtype=np.int32)
    nk = np.empty(modulo, dtype=np.uint32)
    kmers_out = np.empty(n, dtype=np.uint64)
    outs = {name: np.empty_like(col) for name, col in direct.items()}
    freq = np.empty(n, dtype=np.uint16)
    perm = np.empty(n, dtype=np.uint32) if generic else None
    # set_frequencies compares ref offsets: a ref column of another width is widened for that purpose only
    ref_in = direct.get("ref")
    ref_for_freq = ref_in if ref_in is not None else (_view64(generic["ref"]) if ("ref" in generic and not skip_frequencies) else None)
    _lib.call("gki_index_build", _lib.ptr(k64), _lib.ptr(direct.get("nodes")), _lib.ptr(ref_for_freq), _lib.ptr(direct.get("af")),
              n, modulo, _lib.GKI_BUILD_SKIP_FREQUENCIES if skip_frequencies else 0, _lib.ptr(h2i), _lib.ptr(nk),
              _lib.ptr(kmers_out), _lib.ptr(outs.get("nodes")), _lib.ptr(outs.get("ref")), _lib.ptr(outs.get("af")),
              _lib.ptr(freq), _lib.ptr(perm), _lib.current_stream())
    for name, col in generic.items():
        out = np.empty_like(col)
        _lib.call("gki_gather", _lib.ptr(col), col.dtype.itemsize, _lib.ptr(perm), n, _lib.ptr(out), _lib.current_stream())
        outs[name] = out
    return h2i, nk, kmers_out.view(kmers.dtype) if kmers.dtype.itemsize == 8 else kmers_out, outs.get("nodes"), outs.get("ref"), outs.get("af"), freq


class DeviceCounter:
    """Stands where npstructures.Counter stands in the reference (cfki:27): one counter per distinct index k-mer,
    held in HBM next to the index.  ``count`` ignores absent keys; ``counter[keys]`` reads counters."""

    def __init__(self, device_index):
        self._device = device_index

    def count(self, keys):
        self._device.count_kmers(keys if not isinstance(keys, np.ndarray) else _queries(keys))

    def __getitem__(self, keys):
        return self._device.query_counts(keys)

    def fill(self, value):
        assert value == 0
        self._device.reset_counts()


class CounterKmerIndex:
    """collision_free_kmer_index.py:14-40."""

    def __init__(self, kmers, nodes, counter):
        self.kmers = kmers
        self.nodes = nodes
        self.counter = counter

    @classmethod
    def from_kmer_index(cls, kmer_index, k=None):
        """cfki:20-28.  The counters live on the device in a table over the distinct index k-mers (csrc/count.cu).
        `k` (extension, optional): the k-mer length, which lets both strands of a read position share one probe."""
        kmers = kmer_index._kmers      # cfki:22 takes an int64 copy; 64-bit integers keep their bits, so a view serves
        kmers = kmers.view(np.int64) if isinstance(kmers, np.ndarray) and kmers.dtype == np.uint64 else np.asarray(kmers).astype(np.int64)
        nodes = kmer_index._nodes
        # cfki:27 gives every CounterKmerIndex a fresh zeroed Counter.  The counters live inside the device index's count table, so
        # the first counter made from an index uses the index's cached device copy and every further one that is alive at the same
        # time gets a device copy (and count table) of its own.
        device = kmer_index.device_index()
        owner = getattr(kmer_index, "_counter_owner", None)
        if owner is not None and owner() is not None and owner().counter._device is device:
            device = DeviceIndex.from_index(kmer_index)
            shared = False
        else:
            shared = True
        if k is not None:
            device.prepare_counting(k)
        device.reset_counts()
        obj = cls(kmers, nodes, DeviceCounter(device))
        if shared:
            try:
                kmer_index._counter_owner = weakref.ref(obj)
            except AttributeError:        # an index-like object with __slots__: nothing to share, nothing to guard
                pass
        return obj

    def reset(self):
        """cfki:30-31 (the reference rebinds the Counter to np.zeros_like(counter); the intent -- zero every count -- is kept)."""
        self.counter.fill(0)

    def count_kmers(self, kmers, update_counter=True):
        """cfki:33-37."""
        if not update_counter:
            self.reset()
        if isinstance(kmers, np.ndarray):
            # `kmers.astype(np.int64)` of the reference keeps the bits of a 64-bit integer array: no copy is needed for those
            kmers = kmers.view(np.uint64) if kmers.dtype in (np.uint64, np.int64) and kmers.flags.c_contiguous else kmers.astype(np.int64).view(np.uint64)
        self.counter.count(kmers)

    def count_reads(self, reads, k, both_strands=True, update_counter=True):
        """Fused read hashing + counting (read_kmers.py:14-26 feeding cfki:33-37) without materialising hashes."""
        if not update_counter:
            self.reset()
        self.counter._device.count_reads(reads, k, both_strands)

    def count_fasta(self, fasta_file_name, k, both_strands=True):
        """the reads of ReadKmers.from_fasta_file(fasta_file_name, k) (read_kmers.py:14-27: forward, then reverse complement)
        counted without materialising their hashes; FASTQ files work too"""
        return self.counter._device.count_fastx(fasta_file_name, k, both_strands)

    def get_node_counts(self, min_nodes=0):
        """cfki:39-40: np.bincount(nodes, counter[kmers], minlength=min_nodes) -> float64."""
        return self.counter._device.node_counts(min_nodes)


class MinimalKmerIndex:
    """collision_free_kmer_index.py:44-106."""

    def __init__(self, hashes_to_index, n_kmers, nodes, kmers, modulo):
        self._hashes_to_index = hashes_to_index.astype(np.int64)
        self._n_kmers = n_kmers.astype(np.uint32)
        self._nodes = nodes.astype(np.uint32)
        self._kmers = kmers
        self._modulo = np.int64(modulo)

    def max_node_id(self):
        return np.max(self._nodes)

    def to_file(self, file_name):
        np.savez(file_name, hashes_to_index=self._hashes_to_index, n_kmers=self._n_kmers, nodes=self._nodes,
                 kmers=self._kmers, modulo=self._modulo)

    @classmethod
    def from_file(cls, file_name):
        try:
            data = np.load(file_name + ".npz")
        except FileNotFoundError:
            data = np.load(file_name)
        return cls(data["hashes_to_index"], data["n_kmers"], data["nodes"], data["kmers"], data["modulo"])

    @classmethod
    def from_flat_kmers(cls, flat_kmers, modulo=DEFAULT_MODULO):
        """cfki:75-106 (the reference's np.int at :98 no longer exists in numpy; the int64 table it meant is produced)."""
        h2i, nk, kmers, nodes, _, _, _ = build_index_arrays(flat_kmers._hashes, flat_kmers._nodes, None, None, modulo, True)
        return cls(h2i, nk, nodes, kmers, modulo)


class _KeyedCounts:
    """What `HashTable(unique_kmers, counts)` is to KmerIndex2 (cfki:148-158): counts[kmer] for a scalar or an array of known keys."""

    def __init__(self, keys, counts):
        self._keys, self._values = keys, counts

    def __getitem__(self, keys):
        keys = np.asarray(keys).astype(self._keys.dtype)
        at = np.searchsorted(self._keys, keys)
        if np.any(at >= len(self._keys)) or np.any(self._keys[np.minimum(at, len(self._keys) - 1)] != keys):
            raise IndexError("key not in the table")
        return self._values[at]


class KmerIndex2:
    """cfki:110-158: like CollisionFreeKmerIndex, with start node / start offset instead of ref_offset -- the index of
    DenseKmerFinder.get_flat_kmers(v="2").  Built on the device through MultiValueHashTable."""

    def __init__(self, data, frequencies=None):
        self._data = data
        self._frequencies = frequencies

    def get_start_nodes(self, kmer):
        return self._data[kmer]["start_nodes"]

    def get_start_offsets(self, kmer):
        return self._data[kmer]["start_offsets"]

    def get_nodes(self, kmer):
        return self._data[kmer]["nodes"]

    def get_all_kmers(self):
        return self._data.get_all_keys()

    def get_kmer_frequency(self, kmer):
        assert self._frequencies is not None, "Frequencies not set"
        return self._frequencies[kmer]

    @classmethod
    def from_flat_kmers(cls, flat_kmers, modulo=None, skip_frequencies=False):
        from .multi_value_hashtable import MultiValueHashTable
        hash_table = MultiValueHashTable.from_keys_and_values(
            flat_kmers._hashes, {"nodes": flat_kmers._nodes, "start_nodes": flat_kmers._start_nodes,
                                 "start_offsets": flat_kmers._start_offsets, "allele_frequencies": flat_kmers._allele_frequencies},
            mod=modulo)
        index = cls(hash_table)
        if not skip_frequencies:
            index.count_unique_kmer_occurences()
        return index

    def count_unique_kmer_occurences(self):
        """cfki:148-158: per distinct k-mer the number of distinct (start node, start offset) pairs.  That is set_frequencies
        (cfki:267-293) with the pair as the ref offset, so it runs in the same device pass (uint16 there: k-mers of a bucket with
        65536 or more entries are recounted here)."""
        table = self._data._hash_table
        rows = table._nodes
        start = (np.asarray(self._data._values["start_nodes"]).astype(np.int64) << 32) | \
            (np.asarray(self._data._values["start_offsets"]).astype(np.int64) & 0xffffffff)      # nodes < 2^31, offsets of any integer width
        start = start[rows]                                              # in the table's entry order
        # the entries are already in bucket order and the sort is stable: frequencies come back aligned with them
        frequencies = build_index_arrays(table._kmers, None, np.ascontiguousarray(start).view(np.uint64), None, table._modulo, False)[6]
        unique_kmers, first = np.unique(table._kmers, return_index=True)
        counts = np.zeros_like(unique_kmers)
        counts[:] = frequencies[first]
        for bucket in np.flatnonzero(table._n_kmers >= 65536):
            lo = int(table._hashes_to_index[bucket])
            kmers, pairs = table._kmers[lo:lo + int(table._n_kmers[bucket])], start[lo:lo + int(table._n_kmers[bucket])]
            for kmer in np.unique(kmers):
                counts[np.searchsorted(unique_kmers, kmer)] = len(np.unique(pairs[kmers == kmer]))
        self._frequencies = _KeyedCounts(unique_kmers, counts)


class CollisionFreeKmerIndex:
    """collision_free_kmer_index.py:163-490."""

    properties = {"_hashes_to_index", "_n_kmers", "_nodes", "_ref_offsets", "_kmers", "_modulo", "_frequencies",
                  "_allele_frequencies"}

    def __init__(self, _hashes_to_index=None, _n_kmers=None, _nodes=None, _ref_offsets=None, _kmers=None,
                 _modulo=DEFAULT_MODULO, _frequencies=None, _allele_frequencies=None):
        self._hashes_to_index = _hashes_to_index
        self._n_kmers = _n_kmers
        self._nodes = _nodes
        self._ref_offsets = _ref_offsets
        self._kmers = _kmers
        self._modulo = int(_modulo)
        self._frequencies = 0 if _frequencies is None else _frequencies
        self._allele_frequencies = _allele_frequencies
        self._device = None
        self._device_key = None
        self._device_version = 0
        self._counter_owner = None

    # ---- device residency ----
    _DEVICE_COLUMNS = ("_hashes_to_index", "_n_kmers", "_nodes", "_kmers", "_frequencies", "_ref_offsets", "_allele_frequencies")

    def invalidate_device(self):
        """Call after editing any of the host arrays in place (``index._frequencies[...] = x``): the next lookup uploads the index
        again.  Every method of this class that changes an array does it itself.  A device copy that a CounterKmerIndex or a
        CythonKmerIndex still holds stays alive (and keeps its counts) until they let go of it."""
        self._device_version += 1

    def device_index(self):
        key = (self._device_version, int(self._modulo)) + tuple(id(getattr(self, a)) for a in self._DEVICE_COLUMNS)
        if self._device is None or self._device_key != key:
            self._device = DeviceIndex.from_index(self)      # the previous copy is released when its last holder drops it
            self._device_key = key
        return self._device

    def clear(self):
        self._hashes_to_index = None
        self._n_kmers = None
        self._nodes = None
        self._kmers = None
        self._modulo = None
        self._device = None              # closed by DeviceIndex.__del__ once no counter holds it any more
        self._device_key = None
        gc.collect()

    def copy(self):
        return CollisionFreeKmerIndex(self._hashes_to_index.copy(), self._n_kmers.copy(), self._nodes.copy(),
                                      self._ref_offsets.copy(), self._kmers.copy(), self._modulo,
                                      self._frequencies.copy(), self._allele_frequencies.copy())

    # ---- batched lookups (cfki:210-232; kmer_mapper.mapper in the reference) ----
    def map_kmers(self, kmers, n_nodes):
        return self.device_index().map_kmers(kmers, n_nodes)

    def has_kmers(self, kmers):
        return self.device_index().has_kmers(kmers)

    def get_kmers(self):
        return self._kmers

    def has_kmers_parallel(self, kmers, n_threads):
        """cfki:222-232: the reference slices the queries over n_threads processes; one GPU launch covers them all."""
        return self.has_kmers(kmers)

    def set_allele_frequencies(self, frequencies):
        pass

    def max_node_id(self):
        return np.max(self._nodes)

    def convert_to_int32(self):
        """cfki:240-244."""
        self._hashes_to_index = self._hashes_to_index.astype(np.int32)
        self._nodes = self._nodes.astype(np.int32)
        self._n_kmers = self._n_kmers.astype(np.int32)
        self._modulo = np.uint64(self._modulo)
        self.invalidate_device()

    def remove_ref_offsets(self):
        self._ref_offsets = np.array([0])
        self.invalidate_device()

    def remove_frequencies(self):
        self._frequencies = np.array([0])
        self.invalidate_device()

    def set_frequencies_using_other_index(self, other, multiplier=1, min_frequency=1):
        """cfki:252-265, batched: one device lookup for the positions of every distinct k-mer."""
        unique = np.unique(self._kmers)
        entries, qidx = self.device_index().lookup_entries(unique)
        if hasattr(other, "get_frequencies"):
            values = np.maximum(min_frequency, other.get_frequencies(unique) * multiplier)
        else:
            values = np.array([max(min_frequency, other.get_frequency(int(kmer)) * multiplier) for kmer in unique])
        self._frequencies[entries] = values[qidx]
        self.invalidate_device()

    def set_frequencies(self, skip=False):
        """cfki:267-293 on the device (gki_index_build computes the same quantity during construction)."""
        self._frequencies = np.zeros(len(self._kmers), dtype=np.uint16)
        if skip:
            return
        # rebuild frequencies only: the entries are already in bucket order, a stable sort keeps them in place
        _, _, _, _, _, _, freq = build_index_arrays(self._kmers, None, self._ref_offsets, None, self._modulo, False)
        self._frequencies = freq
        self.invalidate_device()

    def __contains__(self, item):
        return self.get(int(item), 100000000000)[0] is not None

    def get_nodes(self, kmer, max_hits=10):
        return self.get(kmer, max_hits)[0]

    def get(self, kmer, max_hits=10):
        """cfki:303-315."""
        hit, _ = self.device_index().lookup_entries(np.array([int(kmer)], dtype=np.uint64))
        frequencies = self._frequencies[hit]
        allele_frequencies = self._allele_frequencies[hit]
        if len(hit) == 0 or frequencies[0] > max_hits:
            return None, None, None, None
        return self._nodes[hit], self._ref_offsets[hit], frequencies, allele_frequencies

    def get_grouped_nodes(self, kmer, max_hits=10):
        """cfki:317-334."""
        hits = self.get(kmer, max_hits)
        if hits[0] is None:
            return None
        ref_offsets, nodes = hits[1], hits[0]
        sorting = np.argsort(ref_offsets)
        ref_offsets, nodes = ref_offsets[sorting], nodes[sorting]
        _, hit_indexes = np.unique(ref_offsets, return_index=True)
        hit_indexes = list(hit_indexes) + [len(ref_offsets)]
        return [nodes[s:e] for s, e in zip(hit_indexes[0:-1], hit_indexes[1:])]

    def get_frequency(self, kmer, include_reverse_complement=True, k=31):
        """cfki:336-352."""
        nodes, _, frequencies, _ = self.get(kmer, max_hits=1000000000000000)
        f = 0 if nodes is None else int(frequencies[0])
        if include_reverse_complement:
            rev_kmer = int(kmer_hashes_to_reverse_complement_hash(np.array([kmer], dtype=np.uint64), k)[0])
            nodes, _, frequencies, _ = self.get(rev_kmer, max_hits=1000000000000000)
            if nodes is not None:
                f += int(frequencies[0])
        return f

    def get_frequencies(self, kmers, include_reverse_complement=True, k=31):
        """`get_frequency` (cfki:336-352) of every k-mer of an array in two device lookups -> int64 array."""
        def first_hit_frequency(queries):
            entries, qidx = self.device_index().lookup_entries(queries)
            found = np.zeros(len(queries), dtype=np.int64)
            if len(entries):
                first_of_query = np.ones(len(qidx), dtype=bool)
                first_of_query[1:] = qidx[1:] != qidx[:-1]
                found[qidx[first_of_query]] = self._frequencies[entries[first_of_query]]     # frequencies[0] of the hits
            return found
        queries = _queries(kmers)
        frequencies = first_hit_frequency(queries)
        if include_reverse_complement:
            frequencies += first_hit_frequency(kmer_hashes_to_reverse_complement_hash(queries, k))
        return frequencies

    def _multi_get(self, kmers, max_hits):
        """Batched form of the per-k-mer loops at cfki:354-391 incl. the frequency gate of `get` (cfki:312)."""
        entries, qidx = self.device_index().lookup_entries(kmers)
        if len(entries) == 0:
            return entries, qidx
        first_of_query = np.ones(len(qidx), dtype=bool)
        first_of_query[1:] = qidx[1:] != qidx[:-1]
        gate = self._frequencies[entries[first_of_query]] > max_hits      # frequencies[0] > max_hits
        blocked = np.zeros(int(qidx.max()) + 1, dtype=bool)
        blocked[qidx[first_of_query]] = gate
        keep = ~blocked[qidx]
        return entries[keep], qidx[keep]

    def get_nodes_and_ref_offsets_from_multiple_kmers(self, kmers, max_hits=10):
        """cfki:354-378."""
        entries, qidx = self._multi_get(kmers, max_hits)
        if len(entries) == 0:
            return np.array([]), np.array([]), np.array([]), np.array([])
        return self._nodes[entries], self._ref_offsets[entries], qidx.astype(np.float64), self._frequencies[entries]

    def get_nodes_from_multiple_kmers(self, kmers, max_hits=10):
        """cfki:380-391."""
        entries, _ = self._multi_get(kmers, max_hits)
        if len(entries) == 0:
            return np.array([])
        return self._nodes[entries]

    # ---- npz (cfki:393-420) ----
    def to_file(self, file_name):
        np.savez(file_name, hashes_to_index=self._hashes_to_index, n_kmers=self._n_kmers, nodes=self._nodes,
                 ref_offsets=self._ref_offsets, kmers=self._kmers, modulo=self._modulo, frequencies=self._frequencies,
                 allele_frequencies=self._allele_frequencies)

    @classmethod
    def from_file(cls, file_name):
        try:
            data = np.load(file_name + ".npz")
        except FileNotFoundError:
            data = np.load(file_name)
        if "allele_frequencies" in data:
            allele_frequencies = data["allele_frequencies"]
        else:
            allele_frequencies = np.zeros(len(data["ref_offsets"]))
        return cls(data["hashes_to_index"], data["n_kmers"], data["nodes"], data["ref_offsets"], data["kmers"],
                   data["modulo"], data["frequencies"], allele_frequencies)

    # ---- construction (cfki:422-467) ----
    @classmethod
    def from_flat_kmers(cls, flat_kmers, modulo=DEFAULT_MODULO, skip_frequencies=False, skip_singletons=False):
        if skip_singletons:
            flat_kmers = flat_kmers.get_new_without_singletons()
        h2i, nk, kmers, nodes, ref_offsets, af, freq = build_index_arrays(
            flat_kmers._hashes, flat_kmers._nodes, flat_kmers._ref_offsets, flat_kmers._allele_frequencies, modulo,
            skip_frequencies)
        obj = cls(h2i, nk, nodes, ref_offsets, kmers, modulo, _frequencies=freq, _allele_frequencies=af)
        if skip_singletons:
            obj._frequencies += 1      # cfki:463-465
        return obj

    def convert_kmers_to_complement(self, k=31, skip_frequencies=True):
        """cfki:470-490."""
        new_kmers = kmer_hashes_to_complement_hashes(self._kmers, k)
        return CollisionFreeKmerIndex.from_flat_kmers(
            FlatKmers(new_kmers, self._nodes, self._ref_offsets, self._allele_frequencies),
            modulo=self._modulo, skip_frequencies=skip_frequencies)


KmerIndex = CollisionFreeKmerIndex
