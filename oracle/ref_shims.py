"""Shims that let the UNMODIFIED reference import in this container (test infrastructure).

Only usable where /root/reference exists (the build container).  Never imported by the
GPU-side tests, smoke() or bench.py: those use the committed fixtures in tests/golden/.

What is stubbed and why (reference file:line of the import that needs it):
  * Bio.Seq.Seq            read_kmers.py:5, collision_free_kmer_index.py:9  (string reverse complement)
  * pyfaidx.Fasta          reference_kmer_index.py (package __init__ pulls it in)
  * npstructures           collision_free_kmer_index.py:8 (Counter, HashTable) -- absent third party
  * obgraph / sortedcontainers / SharedArray ... pulled in by graph_kmer_index/__init__.py:1-12
  * np.ediff1d             numpy>=2 rejects to_begin=1 for uint64 input
                           (collision_free_kmer_index.py:444, :455); numpy 1.x cast by value.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graph_kmer_index"))


class Seq:
    """Stand-in for Bio.Seq.Seq: only reverse_complement() and str()."""

    _table = str.maketrans("ACGTacgtMRWSYKVHDBNmrwsykvhdbn", "TGCAtgcaKYWSRMBDHVNkywsrmbdhvn")

    def __init__(self, x):
        self.x = str(x)

    def reverse_complement(self):
        return Seq(self.x[::-1].translate(self._table))

    def __str__(self):
        return self.x


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _DictHashTable:
    """dict-backed stand-in for npstructures.HashTable (get semantics only)."""

    def __init__(self, keys, values, mod=None, **kw):
        self._d = {int(k): v for k, v in zip(keys, values)}

    def __getitem__(self, keys):
        if np.ndim(keys) == 0:
            return self._d[int(keys)]
        return np.array([self._d[int(k)] for k in keys])

    def __contains__(self, k):
        return int(k) in self._d


class _DictCounter(_DictHashTable):
    """Stand-in for npstructures.Counter(keys, 0, mod=..., value_dtype=...).

    Semantics restated from the npstructures documentation: count(keys) adds one to the
    stored value of every key occurrence that is present, absent keys are ignored.
    """

    def __init__(self, keys, values=0, mod=None, value_dtype=np.uint64, **kw):
        self._keys = np.asarray(keys)
        self._values = np.zeros(len(self._keys), dtype=np.int64)
        self._order = np.argsort(self._keys, kind="stable")
        self._sorted = self._keys[self._order]

    def _find(self, q):
        q = np.asarray(q)
        pos = np.searchsorted(self._sorted, q)
        pos[pos >= len(self._sorted)] = 0
        found = (self._sorted[pos] == q) if len(self._sorted) else np.zeros(len(q), bool)
        return self._order[pos], found

    def count(self, keys):
        idx, found = self._find(keys)
        np.add.at(self._values, idx[found], 1)

    def __getitem__(self, keys):
        idx, found = self._find(keys)
        assert np.all(found)
        return self._values[idx]


_installed = False


def install():
    """Install stubs + numpy shim, put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _stub("Bio")
    _stub("Bio.Seq", Seq=Seq)
    _stub("pyfaidx", Fasta=object)
    _stub("npstructures", Counter=_DictCounter, HashTable=_DictHashTable)
    _stub("npstructures.hashtable", HashTable=_DictHashTable)
    for name in ("obgraph", "obgraph.position_id", "obgraph.variant_to_nodes", "obgraph.haplotype_nodes",
                 "obgraph.graph", "obgraph.variants", "obgraph.genotype_matrix", "obgraph.haplotype_matrix",
                 "shared_memory_wrapper", "shared_memory_wrapper.shared_memory", "SharedArray", "pathos",
                 "pathos.multiprocessing", "sortedcontainers", "kivs", "bionumpy"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    # attributes the reference imports by name
    from oracle import obgraph_standin
    sys.modules["obgraph"].__dict__["Graph"] = obgraph_standin.Graph
    sys.modules["obgraph.position_id"].__dict__["PositionId"] = obgraph_standin.PositionId
    sys.modules["obgraph"].__dict__.setdefault("Graph", object)
    sys.modules["obgraph"].__dict__.setdefault("VariantNotFoundException", Exception)
    sys.modules["obgraph.position_id"].__dict__.setdefault("PositionId", object)
    sys.modules["obgraph.variant_to_nodes"].__dict__.setdefault("VariantToNodes", object)
    sys.modules["obgraph.haplotype_nodes"].__dict__.setdefault("HaplotypeToNodes", object)
    sys.modules["obgraph.graph"].__dict__.setdefault("VariantNotFoundException", Exception)
    sys.modules["obgraph.graph"].__dict__.setdefault("Graph", object)
    sys.modules["obgraph.variants"].__dict__.setdefault("VcfVariants", object)
    sys.modules["obgraph.variants"].__dict__.setdefault("VcfVariant", object)
    sys.modules["obgraph.genotype_matrix"].__dict__.setdefault("GenotypeMatrix", object)
    sys.modules["obgraph.haplotype_matrix"].__dict__.setdefault("HaplotypeMatrix", object)
    sys.modules["shared_memory_wrapper"].__dict__.setdefault("from_shared_memory", None)
    sys.modules["shared_memory_wrapper"].__dict__.setdefault("to_shared_memory", None)
    sys.modules["shared_memory_wrapper"].__dict__.setdefault("SingleSharedArray", object)
    sys.modules["shared_memory_wrapper"].__dict__.setdefault("object_to_shared_memory", None)
    sys.modules["shared_memory_wrapper"].__dict__.setdefault("object_from_shared_memory", None)
    sys.modules["pathos.multiprocessing"].__dict__.setdefault("Pool", object)
    sys.modules["sortedcontainers"].__dict__.setdefault("SortedList", list)

    _orig = np.ediff1d
    if not getattr(_orig, "_gki_shim", False):
        def ediff1d(ary, to_end=None, to_begin=None):
            ary = np.asanyarray(ary)
            cast = lambda v: None if v is None else np.asarray(v).astype(ary.dtype)
            return _orig(ary, to_end=cast(to_end), to_begin=cast(to_begin))
        ediff1d._gki_shim = True
        np.ediff1d = ediff1d

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class stable_argsort:
    """Context manager: make np.argsort default to kind='stable' (canonical payload order,
    SURVEY.md section 8c(ii)); no reference file is edited."""

    def __enter__(self):
        self._orig = np.argsort

        def argsort(a, axis=-1, kind=None, order=None, **kw):
            return self._orig(a, axis=axis, kind="stable" if kind is None else kind, order=order, **kw)
        np.argsort = argsort
        return self

    def __exit__(self, *exc):
        np.argsort = self._orig
        return False
