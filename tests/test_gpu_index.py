"""GPU parity of K2 (index construction) and K3 (lookup / counting / node counts) against the oracle, the
reference's known answers and the golden fixtures generated from the unmodified reference."""
import os

import numpy as np
import pytest

from conftest import golden_index, load_golden
from oracle import c_oracle, numpy_oracle as no

pytestmark = pytest.mark.gpu

COLS = ("hashes_to_index", "n_kmers", "kmers", "nodes", "ref_offsets", "frequencies", "allele_frequencies")


@pytest.fixture(scope="module")
def gki():
    import graph_kmer_index_b200 as g
    return g


def product_index(gki, idx):
    return gki.CollisionFreeKmerIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_nodes"], idx["_ref_offsets"], idx["_kmers"],
                                      idx["_modulo"], idx["_frequencies"], idx["_allele_frequencies"])


def assert_index_equal(index, want, dtypes=True):
    for key in COLS:
        got = getattr(index, "_" + key)
        assert np.array_equal(got, want["_" + key]), key
        if dtypes:
            assert got.dtype == want["_" + key].dtype, (key, got.dtype, want["_" + key].dtype)
    assert int(index._modulo) == want["_modulo"]


# ---- the reference's own test (tests/test_collision_free_kmer_index.py) --------------------------------
def test_reference_collision_free_test(gki, tmp_path):
    flat = gki.FlatKmers(np.array([1, 1, 2, 2, 4, 5, 3], dtype=np.uint64), np.array([5, 6, 7, 8, 10, 11, 100]),
                         np.array([1, 1, 2, 3, 10, 11, 100]))
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=4)
    assert list(index.get(1)[0]) == [5, 6]
    assert list(index.get(1)[1]) == [1, 1]
    index.to_file(str(tmp_path / "tmp.index"))
    index = gki.CollisionFreeKmerIndex.from_file(str(tmp_path / "tmp.index"))
    assert list(index.get(5)[0]) == [11]
    assert list(index.get(3)[0]) == [100] and index.get(7)[0] is None
    n, o, r, f = index.get_nodes_and_ref_offsets_from_multiple_kmers(np.array([1, 5]))
    assert list(n) == [5, 6, 11] and list(o) == [1, 1, 11] and list(r) == [0, 0, 1]
    assert 4 in index and 12 not in index
    index.convert_to_int32()
    kmers = np.array([1, 2, 3, 10, 10, 12, 100, 101, 102, 5], dtype=np.uint64)
    result = index.has_kmers_parallel(kmers, n_threads=3)
    assert np.all(result == [True, True, True, False, False, False, False, False, False, True]), result
    g = load_golden("tiny_index")
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=4)
    assert_index_equal(index, golden_index(g))


@pytest.mark.parametrize("name", ["index_small", "index_sparse"])
def test_build_golden(gki, name):
    g = load_golden(name)
    with_freq = bool(g["stable_frequencies"].any())
    flat = gki.FlatKmers(g["in_hashes"], g["in_nodes"], g["in_ref_offsets"], g["in_allele_frequencies"])
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=int(g["stable_modulo"]), skip_frequencies=not with_freq)
    assert_index_equal(index, golden_index(g))                                   # (ii) canonical (stable) order
    for key in ("hashes_to_index", "n_kmers"):                                  # (i) tables vs the as-run reference
        assert np.array_equal(getattr(index, "_" + key), g["asrun_" + key])
    def canon(kmers, nodes, ref, af, fr):                                        # (iii) per-bucket multisets
        b = kmers % np.uint64(int(g["stable_modulo"]))
        return np.sort(np.rec.fromarrays([b, kmers, nodes, ref, af, fr]))
    assert np.array_equal(canon(index._kmers, index._nodes, index._ref_offsets, index._allele_frequencies, index._frequencies),
                          canon(*[g["asrun_" + k] for k in ("kmers", "nodes", "ref_offsets", "allele_frequencies", "frequencies")]))


@pytest.mark.parametrize("n,modulo,skip", [(1, 1, False), (2, 1, False), (1000, 1, True), (5000, 2, False), (4097, 255, False),
                                           (70000, 256, True), (70000, 257, False), (200000, 65536, False),
                                           (200000, 1000003, False), (300001, 19999999, True), (50000, 4294967291, True)])
def test_build_vs_oracle(gki, n, modulo, skip):
    from graph_kmer_index_b200 import synthetic
    if modulo > 2 ** 31:
        pytest.skip("16 GB of host tables; covered by the device-resident build test")
    hashes, nodes, ref, af = synthetic.flat_kmers(n, max(n // 10, 1), 31)
    if not skip:                                       # several ref offsets per k-mer + one heavy k-mer
        ref = ref.copy()
        ref[::5] += np.uint64(77)
        hashes = hashes.copy()
        hashes[::13] = hashes[0]
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=skip)
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(hashes, nodes, ref, af), modulo=modulo, skip_frequencies=skip)
    assert_index_equal(index, want)
    assert np.all(np.diff((index._kmers % np.uint64(modulo)).astype(np.int64)) >= 0)        # sortedness


@pytest.mark.parametrize("n,modulo,skip,dup", [(400000, 19999999, False, 20), (400000, 19999999, True, 1), (250000, 1000003, False, 6),
                                               (100000, 65521, False, 3), (2000000, 452930477 // 8, False, 40), (3000000, 1000003, True, 2),
                                               (1500000, 300007, False, 4)])
def test_build_paths_vs_oracle(gki, monkeypatch, n, modulo, skip, dup):
    """the three build paths (slab: append-scatter + per-slab ordering in shared memory; binned: 16-entry bins; radix) against the
    oracle and against each other: moderate repeats of a k-mer (several nodes / ref offsets), all columns, frequencies, and the
    permutation output"""
    import ctypes
    from graph_kmer_index_b200 import _lib, synthetic
    hashes, nodes, ref, af = synthetic.flat_kmers(n, max(n // 10, 1), 31)
    rng = np.random.default_rng(n + dup)
    if dup > 1:                                        # groups of up to `dup` entries share a k-mer, with 1-3 distinct ref offsets
        hashes = hashes.copy()
        src = (np.arange(n) // dup) * dup
        pick = rng.random(n) < 0.3
        hashes[pick] = hashes[src[pick]]
        ref = (ref // np.uint64(3)).astype(np.uint64)
    af = rng.random(n).astype(np.float32)
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=skip)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    for path in ("slab", "binned", "radix"):
        monkeypatch.setenv("GKI_BUILD_PATH", path)
        index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=skip)
        assert_index_equal(index, want)
    monkeypatch.setenv("GKI_BUILD_PATH", "slab")
    monkeypatch.setenv("GKI_SLAB_RUNS", "1")               # the run-reserving scatter kernel on rows the sample would not give it
    assert_index_equal(gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=skip), want)
    monkeypatch.delenv("GKI_SLAB_RUNS")
    # kmers + nodes only and the permutation
    outs = {}
    for path in ("slab", "binned", "radix"):
        monkeypatch.setenv("GKI_BUILD_PATH", path)
        h2i, nk = np.empty(modulo, np.int32), np.empty(modulo, np.uint32)
        k_o, n_o, perm = np.empty(n, np.uint64), np.empty(n, np.uint32), np.empty(n, np.uint32)
        _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, modulo, 1, _lib.ptr(h2i), _lib.ptr(nk),
                  _lib.ptr(k_o), _lib.ptr(n_o), None, None, None, _lib.ptr(perm), None)
        outs[path] = (h2i, nk, k_o, n_o, perm)
    monkeypatch.delenv("GKI_BUILD_PATH")
    for path in ("binned", "radix"):
        for a, b in zip(outs["slab"], outs[path]):
            assert np.array_equal(a, b)
    h2i, nk, k_o, n_o, perm = outs["slab"]
    assert np.array_equal(h2i, want["_hashes_to_index"]) and np.array_equal(nk, want["_n_kmers"])
    assert np.array_equal(k_o, hashes[perm]) and np.array_equal(n_o, nodes[perm]) and np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))


@pytest.mark.parametrize("n,modulo", [(400_003, 19999999), (1_200_001, 452930477 // 16), (70_001, 65521)])
def test_build_rows_of_a_kmer_adjacent_as_the_finder_emits_them(gki, monkeypatch, n, modulo):
    """FlatKmers in the order DenseKmerFinder produces (kmer_finder.py:223-240): the rows of one k-mer -- one per node of its path --
    follow each other.  The slab scatter reserves the slots of such a run of lanes with one atomic; runs of 1-5 rows, runs of 40 and
    of 100 rows (longer than a warp), a run across the end of the array."""
    from graph_kmer_index_b200 import _lib, synthetic
    rng = np.random.default_rng(n)
    base_h, _, base_r, _ = synthetic.flat_kmers(n, max(n // 10, 1), 31)
    lens = rng.integers(1, 6, size=n)
    lens[::997] = 40
    lens[5::4999] = 100
    owner = np.repeat(np.arange(n), lens)[:n]                     # row -> k-mer number; the last run is cut by the end of the array
    hashes, ref = base_h[owner], (base_r[owner] + (rng.integers(0, 2, size=n)).astype(np.uint64))
    nodes = rng.integers(0, max(n // 10, 1), size=n).astype(np.uint32)
    af = rng.random(n).astype(np.float32)
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    for path, runs in (("slab", None), ("slab", "0"), ("slab", "1"), ("radix", None)):
        monkeypatch.setenv("GKI_BUILD_PATH", path)
        if runs is None:
            monkeypatch.delenv("GKI_SLAB_RUNS", raising=False)      # the sample of neighbouring rows chooses the scatter kernel
        else:
            monkeypatch.setenv("GKI_SLAB_RUNS", runs)
        assert_index_equal(gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo), want)
    monkeypatch.delenv("GKI_SLAB_RUNS", raising=False)
    monkeypatch.setenv("GKI_BUILD_PATH", "slab")
    h2i, nk, perm = np.empty(modulo, np.int32), np.empty(modulo, np.uint32), np.empty(n, np.uint32)
    _lib.call("gki_index_build", _lib.ptr(hashes), None, None, None, n, modulo, 1, _lib.ptr(h2i), _lib.ptr(nk), None, None, None, None, None,
              _lib.ptr(perm), None)
    assert np.array_equal(h2i, want["_hashes_to_index"]) and np.array_equal(nk, want["_n_kmers"])
    assert np.array_equal(hashes[perm], want["_kmers"]) and np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))


def test_build_dtypes_and_defaults(gki):
    """payload columns keep whatever dtype the caller had (cfki:436-440 are plain fancy-indexing)"""
    rng = np.random.default_rng(2)
    n, modulo = 3000, 1009
    hashes = rng.integers(0, 4 ** 31, n, dtype=np.uint64)
    hashes[::3] = hashes[1]
    nodes = rng.integers(0, 50, n)                      # int64 nodes
    flat = gki.FlatKmers(hashes, nodes)                 # default ref_offsets float64 zeros, af float32 ones
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo)
    want = no.build_index(hashes, nodes, np.zeros(n), np.ones(n, dtype=np.float32), modulo)
    assert_index_equal(index, want)
    for rdt, ndt in ((np.int32, np.uint16), (np.float32, np.int8), (np.uint8, np.uint32)):
        ref = rng.integers(0, 100, n).astype(rdt)
        nd = rng.integers(0, 100, n).astype(ndt)
        af = rng.random(n)                              # float64 allele frequencies
        index = gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(hashes, nd, ref, af), modulo=modulo)
        assert_index_equal(index, no.build_index(hashes, nd, ref, af, modulo))
    with pytest.raises(IndexError):
        gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(np.zeros(0, np.uint64), np.zeros(0, np.uint32)), modulo=7)
    m = gki.MinimalKmerIndex.from_flat_kmers(gki.FlatKmers(hashes, nodes.astype(np.uint32)), modulo=modulo)
    assert m._hashes_to_index.dtype == np.int64 and np.array_equal(m._hashes_to_index, want["_hashes_to_index"])
    assert np.array_equal(m._kmers, want["_kmers"]) and np.array_equal(m._nodes, want["_nodes"])


def test_singletons_and_set_frequencies(gki):
    rng = np.random.default_rng(4)
    hashes = rng.integers(0, 500, 4000).astype(np.uint64)
    nodes = np.arange(4000, dtype=np.uint32)
    ref = rng.integers(0, 5, 4000).astype(np.uint64)
    af = np.ones(4000, dtype=np.float32)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    kept = flat.get_new_without_singletons()
    want = no.without_singletons(hashes, nodes, ref, af)
    for a, b in zip((kept._hashes, kept._nodes, kept._ref_offsets, kept._allele_frequencies), want):
        assert np.array_equal(a, b)
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=101, skip_singletons=True)
    w = no.build_index(*want, 101)
    assert np.array_equal(index._frequencies, w["_frequencies"] + 1) and np.array_equal(index._kmers, w["_kmers"])
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=101, skip_frequencies=True)
    assert not index._frequencies.any()
    index.set_frequencies()
    assert np.array_equal(index._frequencies, no.build_index(hashes, nodes, ref, af, 101)["_frequencies"])
    rc = flat.get_reverse_complement_flat_kmers(31)
    assert np.array_equal(rc._hashes, no.revcomp_hashes(hashes, 31))
    both = gki.FlatKmers.from_multiple_flat_kmers([flat, rc])
    assert len(both._hashes) == 8000 and both._nodes.dtype == np.uint32 and both._ref_offsets.dtype == np.uint64


def test_one_kmer_hundreds_of_thousands_of_times(gki):
    """a k-mer present 200 000 times (poly-A, N runs on real graphs) and one present 70 000 times at distinct ref offsets: the build
    falls back to the radix path and its frequency passes, and the first-occurrence marking, must stay linear in the length of such
    a bucket (pairwise comparison would be 10^10 loads here).  Frequencies are checked against cfki:286-290 restated with np.unique
    (the C oracle compares pairwise too), the other arrays against the oracle without frequencies."""
    import time
    from graph_kmer_index_b200 import synthetic
    n, modulo = 600_000, 1_000_003
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 5000, 31)
    rng = np.random.default_rng(11)
    hashes, ref = hashes.copy(), ref.copy()
    a = rng.choice(n, 270_000, replace=False)
    hashes[a[:200_000]] = hashes[a[0]]
    ref[a[:200_000]] = rng.integers(0, 1000, 200_000).astype(np.uint64)          # 1000 distinct ref offsets
    hashes[a[200_000:]] = hashes[a[200_000]]
    ref[a[200_000:]] = np.arange(70_000, dtype=np.uint64) + np.uint64(5)          # 70 000 distinct: the uint16 frequency wraps (cfki:270)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(hashes[:50_000], nodes[:50_000], ref[:50_000], af[:50_000]), modulo=modulo)   # warm
    t0 = time.perf_counter()
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo)
    seconds = time.perf_counter() - t0
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    for key in ("_hashes_to_index", "_n_kmers", "_kmers", "_nodes", "_ref_offsets", "_allele_frequencies"):
        assert np.array_equal(getattr(index, key), want[key]), key
    pairs = np.unique(np.stack([hashes, ref], axis=1), axis=0)
    kmers_u, refs_per_kmer = np.unique(pairs[:, 0], return_counts=True)
    freq = refs_per_kmer[np.searchsorted(kmers_u, index._kmers)].astype(np.uint16)
    assert np.array_equal(index._frequencies, freq)
    assert refs_per_kmer.max() >= 70_000 and 1000 <= int(index._frequencies[index._kmers == hashes[a[0]]][0]) <= 1001      # one wraps uint16, one does not
    assert seconds < 5.0, seconds
    t0 = time.perf_counter()
    kept = flat.get_new_without_singletons()
    seconds = time.perf_counter() - t0
    _, first = np.unique(hashes, return_index=True)
    keep = np.ones(n, dtype=bool)
    keep[first] = False
    assert np.array_equal(kept._hashes, hashes[keep]) and np.array_equal(kept._nodes, nodes[keep]) and np.array_equal(kept._ref_offsets, ref[keep])
    assert seconds < 5.0, seconds


@pytest.mark.parametrize("name", ["index_small", "index_sparse"])
def test_lookup_and_counts_golden(gki, name):
    g = load_golden(name)
    idx = golden_index(g)
    index = product_index(gki, idx)
    q = g["queries"]
    nodes, offs, ns = [], [], []
    for kmer in q[:200]:
        r = index.get(kmer, max_hits=10 ** 9)
        ns.append(0 if r[0] is None else len(r[0]))
        if r[0] is not None:
            nodes.extend(r[0]); offs.extend(r[1])
    assert ns == list(g["get_n"]) and nodes == list(g["get_nodes"]) and offs == list(g["get_ref_offsets"])
    counter = gki.CounterKmerIndex.from_kmer_index(index)
    counter.count_kmers(q)
    counter.count_kmers(q[:100])
    got = counter.get_node_counts()
    assert got.dtype == np.float64 and np.array_equal(got, g["node_counts_min0"])
    assert np.array_equal(counter.get_node_counts(int(g["n_nodes"]) + 17), g["node_counts_min_big"])
    assert np.array_equal(counter.counter[counter.kmers], c_oracle.count_kmers(idx, np.concatenate([q, q[:100]])))
    counter.count_kmers(q[:100], update_counter=False)                       # cfki:34-35: reset first
    assert np.array_equal(counter.get_node_counts(), no.node_counts(idx, q[:100]))
    if "cython_get" in g.files:
        assert np.array_equal(gki.CythonKmerIndex(index).get(q), g["cython_get"])
    assert np.array_equal(index.device_index().lookup_hits(q, False, None, None), no.lookup_hits(idx, q, False, None, None))
    assert np.array_equal(index.has_kmers(q), no.has_kmers(idx, q))
    n_nodes = int(g["n_nodes"])
    assert np.array_equal(index.map_kmers(q, n_nodes), no.map_kmers(idx, q, n_nodes))
    # reads -> node counts, fused
    counter.reset()
    counter.count_reads(g["reads"], int(g["k"]))
    assert np.array_equal(counter.get_node_counts(n_nodes), g["read_node_counts"])


@pytest.mark.parametrize("mode", ["canonical", "raw", "nofilter", "tinyfilter", "k_mismatch", "minimizer", "minimizer_k_mismatch"])
@pytest.mark.parametrize("n,modulo,n_reads,L,k", [(40000, 200003, 3000, 150, 31), (40000, 4099, 1500, 150, 31), (5000, 7, 300, 100, 15),
                                                   (40000, 200003, 777, 64, 31), (3000, 1009, 400, 90, 16)])
def test_count_reads_vs_oracle(gki, monkeypatch, mode, n, modulo, n_reads, L, k):
    """every layout of the counting structure gives the oracle's node counts: canonical keys (both strands share a
    probe), raw keys, no Bloom filter, a saturated filter, and a table prepared for another k"""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 997, k)
    if k % 2 == 0:                                     # even k: palindromic k-mers exist; make sure some are indexed and read
        pal = np.array([no.sequence_to_kmer_hash("ACGT" * (k // 4))], dtype=np.uint64)
        assert no.revcomp_hashes(pal, k)[0] == pal[0]
        hashes = hashes.copy()
        hashes[:3] = pal[0]
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    reads = synthetic.reads(n_reads, L, n, k, p_hit_permille=400, n_permille=10)
    if k % 2 == 0:
        reads[:5, :k] = np.frombuffer(("ACGT" * (k // 4)).encode(), dtype=np.uint8)
    if mode == "raw":
        monkeypatch.setenv("GKI_TABLE_RAW", "1")
    elif mode == "nofilter":
        monkeypatch.setenv("GKI_FILTER_MAX_MB", "0")
    elif mode == "tinyfilter":
        monkeypatch.setenv("GKI_FILTER_K", "1")
    elif mode.startswith("minimizer"):                 # filter words addressed by minimizer (odd k in 27..31, else it stays off)
        monkeypatch.setenv("GKI_FILTER_MZ", "1")
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], modulo)
    dev.prepare_counting((k + 1 if k < 31 else k - 1) if mode == "k_mismatch" else (29 if mode == "minimizer_k_mismatch" else (0 if mode == "raw" else k)))
    assert dev.info()["has_filter"] == (mode != "nofilter")
    want = c_oracle.read_node_counts(idx, reads, k, 1000)
    assert want.sum() > 0
    dev.count_reads(reads, k)                                   # host buffer: chunked H2D inside the call
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    dreads = torch.from_numpy(reads).cuda()
    dev.count_reads(dreads, k)                                  # device-resident
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    dev.count_reads(dreads, k, both_strands=False)
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads, k, 1000, both_strands=False))
    # unfused: hashes -> count_kmers equals the fused path; linearity over two halves
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    fwd, rc = hash_read_matrix(reads, k)
    dev.reset_counts()
    half = n_reads // 2
    dev.count_kmers(np.concatenate([fwd[:half].ravel(), rc[:half].ravel()]))
    first = dev.node_counts(1000)
    dev.count_kmers(torch.from_numpy(np.concatenate([fwd[half:].ravel(), rc[half:].ravel()]).view(np.int64)).cuda())
    assert np.array_equal(dev.node_counts(1000), want)
    assert np.array_equal(want - first, c_oracle.read_node_counts(idx, reads[half:], k, 1000))
    # strided host rows
    padded = np.full((n_reads, L + 10), ord("A"), dtype=np.uint8)
    padded[:, :L] = reads
    dev.reset_counts()
    dev.count_reads(padded[:, :L], k)
    assert np.array_equal(dev.node_counts(1000), want)
    assert np.array_equal(dev.entry_counts(), c_oracle.count_reads(idx, reads, k))
    dev.close()


def test_uint16_wrap_flag_and_hot_kmer(gki):
    """a k-mer hit > 65535 times: exact by default, modulo 2^16 with the reference-Counter compat flag (cfki:27)"""
    kmers = np.array([5, 5, 9], dtype=np.uint64)
    nodes = np.array([1, 2, 3], dtype=np.uint32)
    idx = no.build_index(kmers, nodes, np.zeros(3, np.uint64), np.ones(3, np.float32), 11, skip_frequencies=True)
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], 11)
    dev.count_kmers(np.full(70000, 5, dtype=np.uint64))
    dev.count_kmers(np.array([9, 9, 4, 16], dtype=np.uint64))
    assert list(dev.node_counts()) == [0, 70000, 70000, 2]
    assert list(dev.node_counts(wrap_uint16=True)) == [0, 70000 - 65536, 70000 - 65536, 2]
    assert list(dev.node_counts(6)) == [0, 70000, 70000, 2, 0, 0]


def test_counter_past_2_pow_32_does_not_leak_into_the_other_orientation(gki):
    """Counters are 32 bits per k-mer orientation and wrap (the reference's Counter wraps at 2^16, cfki:27).  Poly-A reads counted on
    both strands push the counter of AAA..A (orientation 0 of the canonical key) and of TTT..T (orientation 1) past 2^32: each must
    come back modulo 2^32 on its own -- a packed 64-bit add of {1, 1} carried the overflow of one into the other."""
    import torch
    k, L = 31, 150
    poly_a, poly_t = 0, 4 ** k - 1
    kmers = np.array([poly_a, poly_t, 12345], dtype=np.uint64)
    nodes = np.array([1, 2, 3], dtype=np.uint32)
    idx = no.build_index(kmers, nodes, np.zeros(3, np.uint64), np.ones(3, np.float32), 101, skip_frequencies=True)
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], 101)
    dev.prepare_counting(k)
    n_reads, calls = 1_200_000, 30
    reads = torch.full((n_reads, L), ord("A"), dtype=torch.uint8, device="cuda")
    for _ in range(calls):
        dev.count_reads(reads, k, True)
    positions = n_reads * (L - k + 1) * calls
    assert positions > 2 ** 32
    dev.count_kmers(np.full(3, poly_a, dtype=np.uint64))          # orientation 0 alone
    dev.count_kmers(np.full(5, poly_t, dtype=np.uint64))          # orientation 1 alone
    got = dev.node_counts()
    assert got[1] == (positions + 3) % 2 ** 32 and got[2] == (positions + 5) % 2 ** 32 and got[3] == 0
    assert list(dev.query_counts(np.array([poly_a, poly_t, 12345, 7], dtype=np.uint64))) == [(positions + 3) % 2 ** 32, (positions + 5) % 2 ** 32, 0, 0]
    dev.close()


def test_device_resident_build_and_count_full_config1(gki):
    """BASELINE config 1 (1M entries, 100k nodes, 100k x 150bp reads) entirely on the device; the oracle checks it."""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    n, n_nodes, modulo, k, n_reads, L = 1_000_000, 100_000, 19_999_999, 31, 100_000, 150
    glen = synthetic.genome_length(n, k)
    dev = torch.device("cuda")
    genome = torch.empty(glen, dtype=torch.uint8, device=dev)
    _lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
    hashes = torch.empty(n, dtype=torch.int64, device=dev)
    nodes = torch.empty(n, dtype=torch.int32, device=dev)
    ref = torch.empty(n, dtype=torch.int64, device=dev)
    af = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), None)
    reads = torch.empty((n_reads, L), dtype=torch.uint8, device=dev)
    _lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, n_reads, L, 100, 0, _lib.ptr(reads), None)
    torch.cuda.synchronize()
    # device generators == host mirror
    h_hashes, h_nodes, h_ref, h_af = synthetic.flat_kmers(n, n_nodes, k)
    assert np.array_equal(hashes.cpu().numpy().view(np.uint64), h_hashes) and np.array_equal(nodes.cpu().numpy().view(np.uint32), h_nodes)
    assert np.array_equal(ref.cpu().numpy().view(np.uint64), h_ref) and np.array_equal(af.cpu().numpy(), h_af)
    h_reads = synthetic.reads(n_reads, L, n, k, 100)
    assert np.array_equal(reads.cpu().numpy(), h_reads)
    # build on the device
    h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
    nk = torch.empty(modulo, dtype=torch.int32, device=dev)
    o_k, o_r = torch.empty_like(hashes), torch.empty_like(ref)
    o_n, o_a = torch.empty_like(nodes), torch.empty_like(af)
    o_f = torch.empty(n, dtype=torch.int16, device=dev)
    _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), n, modulo, 0, _lib.ptr(h2i), _lib.ptr(nk),
              _lib.ptr(o_k), _lib.ptr(o_n), _lib.ptr(o_r), _lib.ptr(o_a), _lib.ptr(o_f), None, None)
    torch.cuda.synchronize()
    want = c_oracle.build_index(h_hashes, h_nodes, h_ref, h_af, modulo)
    assert np.array_equal(h2i.cpu().numpy(), want["_hashes_to_index"]) and np.array_equal(nk.cpu().numpy().view(np.uint32), want["_n_kmers"])
    assert np.array_equal(o_k.cpu().numpy().view(np.uint64), want["_kmers"]) and np.array_equal(o_n.cpu().numpy().view(np.uint32), want["_nodes"])
    assert np.array_equal(o_r.cpu().numpy().view(np.uint64), want["_ref_offsets"]) and np.array_equal(o_a.cpu().numpy(), want["_allele_frequencies"])
    assert np.array_equal(o_f.cpu().numpy().view(np.uint16), want["_frequencies"])
    # count on the device
    index = gki.DeviceIndex(h2i, nk, o_k, o_n, modulo)
    index.count_reads(reads, k)
    got = index.node_counts(n_nodes)
    assert np.array_equal(got, c_oracle.read_node_counts(want, h_reads, k, n_nodes))
    # checksum-of-checksums: total node count == sum over entries of their k-mer's hit count
    assert got.sum() == float(index.entry_counts().astype(np.int64).sum())


def test_large_modulo_device_build(gki):
    """default-sized tables (modulo 452930477) stay on the device: tables checked through their invariants"""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    n, modulo, k = 2_000_000, 452930477, 31
    h_hashes, h_nodes, _, _ = synthetic.flat_kmers(n, 1000, k)
    dev = torch.device("cuda")
    hashes = torch.from_numpy(h_hashes.view(np.int64)).to(dev)
    nodes = torch.from_numpy(h_nodes.view(np.int32)).to(dev)
    h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
    nk = torch.empty(modulo, dtype=torch.int32, device=dev)
    o_k, o_n = torch.empty_like(hashes), torch.empty_like(nodes)
    _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, modulo, 1, _lib.ptr(h2i), _lib.ptr(nk),
              _lib.ptr(o_k), _lib.ptr(o_n), None, None, None, None, None)
    torch.cuda.synchronize()
    assert int(nk.sum()) == n
    b = (h_hashes % np.uint64(modulo)).astype(np.int64)
    order = np.argsort(b, kind="stable")
    assert np.array_equal(o_k.cpu().numpy().view(np.uint64), h_hashes[order]) and np.array_equal(o_n.cpu().numpy().view(np.uint32), h_nodes[order])
    ub, first, cnt = np.unique(b[order], return_index=True, return_counts=True)
    sel = torch.from_numpy(ub).to(dev)
    assert np.array_equal(h2i[sel].cpu().numpy(), first) and np.array_equal(nk[sel].cpu().numpy(), cnt)
    assert int((nk != 0).sum()) == len(ub) and int((h2i != 0).sum()) <= len(ub)


@pytest.mark.parametrize("world,n,heavy", [(1, 50_000, True), (2, 50_000, True), (3, 50_000, True), (8, 50_000, True), (2, 400_000, False),
                                           (3, 600_000, False)])
def test_partitioned_build_emulated_on_one_gpu(gki, world, n, heavy):
    """hash-range partitioned build with the ranks emulated one after another on one device: partition by
    owner -> exchange -> per-range build; the slices concatenated in rank order are the oracle's single index"""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    from graph_kmer_index_b200.distributed import bucket_range, shard_bounds
    modulo, k = (100_003 if heavy else 1_000_003), 31
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 777, k)
    hashes = hashes.copy()
    if heavy:
        hashes[::17] = hashes[3]                 # a heavy k-mer: one bucket far longer than the others
    else:
        m3 = len(hashes[1::3])
        hashes[:3 * m3:3] = hashes[1::3]             # pairs of entries share a k-mer (the larger shards take the slab path, with a position offset)
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=False)
    dev = torch.device("cuda")
    shards = [shard_bounds(n, r, world) for r in range(world)]
    sent = []                                   # sent[src][dst] = dict of column chunks
    for lo, hi in shards:
        cols = dict(kmers=torch.from_numpy(hashes[lo:hi].view(np.int64)).to(dev), nodes=torch.from_numpy(nodes[lo:hi].view(np.int32)).to(dev),
                    ref=torch.from_numpy(ref[lo:hi].view(np.int64)).to(dev), af=torch.from_numpy(af[lo:hi]).to(dev))
        m = hi - lo
        perm = torch.empty(m, dtype=torch.int32, device=dev)
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        _lib.call("gki_partition_by_bucket_range", _lib.ptr(cols["kmers"]), m, modulo, world, _lib.ptr(perm), _lib.ptr(counts), None)
        torch.cuda.synchronize()
        assert int(counts.sum()) == m
        bounds = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
        g = {}
        for name, col in cols.items():
            out = torch.empty_like(col)
            _lib.call("gki_gather", _lib.ptr(col), col.element_size(), _lib.ptr(perm), m, _lib.ptr(out), None)
            g[name] = out
        sent.append([{name: g[name][bounds[d]:bounds[d + 1]] for name in g} for d in range(world)])
    got = {key: [] for key in ("h2i", "nk", "kmers", "nodes", "ref", "af", "freq")}
    offset = 0
    for r in range(world):
        recv = {name: torch.cat([sent[src][r][name] for src in range(world)]).contiguous() for name in ("kmers", "nodes", "ref", "af")}
        m = int(recv["kmers"].shape[0])
        lo, hi = bucket_range(modulo, r, world)
        h2i = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
        nk = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
        o = {name: torch.empty_like(col) for name, col in recv.items()}
        fr = torch.empty(m, dtype=torch.int16, device=dev)
        if m:
            _lib.call("gki_index_build_range", _lib.ptr(recv["kmers"]), _lib.ptr(recv["nodes"]), _lib.ptr(recv["ref"]), _lib.ptr(recv["af"]), m, modulo,
                      lo, hi, offset, 0, _lib.ptr(h2i), _lib.ptr(nk), _lib.ptr(o["kmers"]), _lib.ptr(o["nodes"]), _lib.ptr(o["ref"]), _lib.ptr(o["af"]),
                      _lib.ptr(fr), None)
        torch.cuda.synchronize()
        offset += m
        for key, t in (("h2i", h2i), ("nk", nk), ("kmers", o["kmers"]), ("nodes", o["nodes"]), ("ref", o["ref"]), ("af", o["af"]), ("freq", fr)):
            got[key].append(t.cpu().numpy())
    cat = {key: np.concatenate(v) for key, v in got.items()}
    assert np.array_equal(cat["h2i"], want["_hashes_to_index"]) and np.array_equal(cat["nk"].view(np.uint32), want["_n_kmers"])
    assert np.array_equal(cat["kmers"].view(np.uint64), want["_kmers"]) and np.array_equal(cat["nodes"].view(np.uint32), want["_nodes"])
    assert np.array_equal(cat["ref"].view(np.uint64), want["_ref_offsets"]) and np.array_equal(cat["af"], want["_allele_frequencies"])
    assert np.array_equal(cat["freq"].view(np.uint16), want["_frequencies"])


@pytest.mark.parametrize("world,n,heavy,columns", [(1, 50_000, True, "all"), (2, 50_000, True, "all"), (8, 50_000, False, "all"), (2, 400_000, False, "all"),
                                                   (3, 600_000, False, "all"), (4, 500_000, False, "narrow"), (32, 300_000, True, "all")])
@pytest.mark.parametrize("runs", [None, "1"])
def test_partitioned_build_with_packed_records_emulated_on_one_gpu(gki, monkeypatch, world, n, heavy, columns, runs):
    """the one-exchange form of the hash-range partitioned build (gki_partition_pack -> all-to-all of 32-byte records ->
    gki_index_build_records), ranks emulated one after another: the slices concatenated in rank order are the oracle's index;
    `runs`: the scatter kernel chosen by the sample of neighbouring rows / the run-reserving one forced"""
    import torch
    if runs is not None:
        monkeypatch.setenv("GKI_SLAB_RUNS", runs)
    from graph_kmer_index_b200 import _lib, synthetic
    from graph_kmer_index_b200.distributed import bucket_range, shard_bounds
    modulo, k = (100_003 if heavy else 1_000_003), 31
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 777, k)
    hashes = hashes.copy()
    if heavy:
        hashes[::17] = hashes[3]
    else:
        m3 = len(hashes[1::3])
        hashes[:3 * m3:3] = hashes[1::3]
    narrow = columns == "narrow"
    want = c_oracle.build_index(hashes, nodes, ref if not narrow else np.zeros(n, np.uint64), af if not narrow else np.zeros(n, np.float32), modulo,
                                skip_frequencies=narrow)
    dev = torch.device("cuda")
    sent = []                                   # sent[src] = (records, bounds)
    for lo, hi in [shard_bounds(n, r, world) for r in range(world)]:
        m = hi - lo
        t = lambda a, dt: torch.from_numpy(a[lo:hi].view(dt)).to(dev)
        d_h, d_n, d_r, d_a = t(hashes, np.int64), t(nodes, np.int32), t(ref, np.int64), t(af, np.float32)      # alive until the call has run
        records = torch.zeros((max(m, 1), 4), dtype=torch.int64, device=dev)
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        _lib.call("gki_partition_pack", _lib.ptr(d_h), _lib.ptr(d_n), None if narrow else _lib.ptr(d_r), None if narrow else _lib.ptr(d_a), m, modulo, world,
                  _lib.ptr(records), _lib.ptr(counts), None)
        torch.cuda.synchronize()
        assert int(counts.sum()) == m
        sent.append((records, np.concatenate([[0], np.cumsum(counts.cpu().numpy())])))
    got = {key: [] for key in ("h2i", "nk", "kmers", "nodes", "ref", "af", "freq")}
    offset = 0
    for r in range(world):
        recv = torch.cat([rec[b[r]:b[r + 1]] for rec, b in sent]).contiguous()
        m = int(recv.shape[0])
        lo, hi = bucket_range(modulo, r, world)
        h2i = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
        nk = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
        o = dict(kmers=torch.empty(m, dtype=torch.int64, device=dev), nodes=torch.empty(m, dtype=torch.int32, device=dev),
                 ref=torch.empty(m, dtype=torch.int64, device=dev), af=torch.empty(m, dtype=torch.float32, device=dev))
        fr = torch.empty(m, dtype=torch.int16, device=dev)
        if m:
            _lib.call("gki_index_build_records", _lib.ptr(recv), m, modulo, lo, hi, offset, 1 if narrow else 0, _lib.ptr(h2i), _lib.ptr(nk), _lib.ptr(o["kmers"]),
                      _lib.ptr(o["nodes"]), None if narrow else _lib.ptr(o["ref"]), None if narrow else _lib.ptr(o["af"]), _lib.ptr(fr), None)
        torch.cuda.synchronize()
        offset += m
        for key, tns in (("h2i", h2i), ("nk", nk), ("kmers", o["kmers"]), ("nodes", o["nodes"]), ("ref", o["ref"]), ("af", o["af"]), ("freq", fr)):
            got[key].append(tns.cpu().numpy())
    cat = {key: np.concatenate(v) for key, v in got.items()}
    assert np.array_equal(cat["h2i"], want["_hashes_to_index"]) and np.array_equal(cat["nk"].view(np.uint32), want["_n_kmers"])
    assert np.array_equal(cat["kmers"].view(np.uint64), want["_kmers"]) and np.array_equal(cat["nodes"].view(np.uint32), want["_nodes"])
    if not narrow:
        assert np.array_equal(cat["ref"].view(np.uint64), want["_ref_offsets"]) and np.array_equal(cat["af"], want["_allele_frequencies"])
    assert np.array_equal(cat["freq"].view(np.uint16), want["_frequencies"])


@pytest.mark.parametrize("slice_mb,ranged", [(None, None), ("16", None), ("16", "0"), ("1", "0"), (None, "0"), ("0", None)])
def test_node_counts_with_more_nodes_than_l2(gki, monkeypatch, slice_mb, ranged):
    """node ids spread over 12 M (96 MB of float64 counts, more than L2 keeps): same counts as the oracle, in one pass ("0"), with the
    (node, weight) words grouped by node range tile by tile (the default for such vectors), and in passes over all words per node range
    (GKI_NODE_RANGED=0: one pass per GKI_NODE_SLICE_MB of counts)"""
    from graph_kmer_index_b200 import synthetic
    if slice_mb is not None:
        monkeypatch.setenv("GKI_NODE_SLICE_MB", slice_mb)
    if ranged is not None:
        monkeypatch.setenv("GKI_NODE_RANGED", ranged)
    n, k, modulo, n_nodes = 300000, 31, 1000003, 12_000_000
    hashes, _, ref, af = synthetic.flat_kmers(n, 1000, k)
    rng = np.random.default_rng(8)
    nodes = rng.integers(0, n_nodes, n).astype(np.uint32)
    nodes[0] = n_nodes - 1
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    reads = synthetic.reads(20000, 150, n, k, p_hit_permille=500, n_permille=5)
    want = c_oracle.read_node_counts(idx, reads, k, n_nodes)
    assert want.sum() > 0
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], modulo)
    dev.prepare_counting(k)
    dev.count_reads(reads, k)
    assert np.array_equal(dev.node_counts(n_nodes), want)
    assert np.array_equal(dev.node_counts(n_nodes + 12345), np.concatenate([want, np.zeros(12345)]))
    want16 = c_oracle.node_counts_from_entry_counts(idx, c_oracle.count_reads(idx, reads, k) & np.uint32(0xffff), n_nodes)
    assert np.array_equal(dev.node_counts(n_nodes, wrap_uint16=True), want16)
    # k-mers counted more often than the packed (node, weight) words of the passes can hold (24 node bits leave 8 for the weight)
    hot = np.concatenate([np.full(100_000, idx["_kmers"][5]), np.full(255, idx["_kmers"][77]), np.full(256, idx["_kmers"][1234])]).astype(np.uint64)
    dev.count_kmers(hot)
    ec = c_oracle.count_reads(idx, reads, k)
    c_oracle.count_kmers(idx, hot, ec)
    assert np.array_equal(dev.node_counts(n_nodes), c_oracle.node_counts_from_entry_counts(idx, ec, n_nodes))
    assert np.array_equal(dev.node_counts(n_nodes, wrap_uint16=True), c_oracle.node_counts_from_entry_counts(idx, ec & np.uint32(0xffff), n_nodes))
    dev.close()


def test_build_into_unaligned_table_views(gki):
    """the dense tables handed to gki_index_build may be views that start on any 4-byte boundary (a rank's slice of a larger
    tensor): same result as into fresh arrays"""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    n, modulo = 200000, 1000003
    hashes, nodes, _, _ = synthetic.flat_kmers(n, 1000, 31)
    d_h, d_n = torch.from_numpy(hashes.view(np.int64)).cuda(), torch.from_numpy(nodes.view(np.int32)).cuda()
    outs = []
    for shift in (0, 1, 3):
        big_h = torch.full((modulo + 8,), -7, dtype=torch.int32, device="cuda")
        big_n = torch.full((modulo + 8,), -7, dtype=torch.int32, device="cuda")
        h2i, nk = big_h[shift:shift + modulo], big_n[shift:shift + modulo]
        k_o, n_o = torch.empty_like(d_h), torch.empty_like(d_n)
        _lib.call("gki_index_build", _lib.ptr(d_h), _lib.ptr(d_n), None, None, n, modulo, 1, h2i.data_ptr(), nk.data_ptr(),
                  _lib.ptr(k_o), _lib.ptr(n_o), None, None, None, None, None)
        torch.cuda.synchronize()
        assert int(big_h[:shift].eq(-7).all()) and int(big_h[shift + modulo:].eq(-7).all())      # nothing written outside the view
        outs.append((h2i.cpu().numpy().copy(), nk.cpu().numpy().copy(), k_o.cpu().numpy(), n_o.cpu().numpy()))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)
    want = c_oracle.build_index(hashes, nodes, np.zeros(n, np.uint64), np.ones(n, np.float32), modulo, skip_frequencies=True)
    assert np.array_equal(outs[0][0], want["_hashes_to_index"]) and np.array_equal(outs[0][1].view(np.uint32), want["_n_kmers"])


@pytest.mark.parametrize("threads,chunk", [(3, 4096), (8, 65536), (1, 1 << 20)])
def test_large_pageable_arrays_take_the_parallel_copy(gki, monkeypatch, threads, chunk):
    """runtime.cu parallel_host_copy (pageable host arrays above GKI_HOST_COPY_MIN_BYTES move through pinned double buffers on
    several host threads): forced on for small arrays with odd sizes, results equal to the plain path's, byte for byte."""
    from graph_kmer_index_b200 import synthetic
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    n, modulo = 300_007, 1_000_003
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 5000, 31)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    reads = synthetic.reads(3001, 150, n, 31, 300, n_permille=5)
    monkeypatch.setenv("GKI_HOST_COPY_THREADS", "0")
    plain = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo)
    plain_hashes = hash_read_matrix(reads, 31)
    monkeypatch.setenv("GKI_HOST_COPY_THREADS", str(threads))
    monkeypatch.setenv("GKI_HOST_COPY_MIN_BYTES", "1000")
    monkeypatch.setenv("GKI_HOST_COPY_CHUNK_BYTES", str(chunk))
    for _ in range(2):
        index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo)
        for key in ("_hashes_to_index", "_n_kmers", "_kmers", "_nodes", "_ref_offsets", "_allele_frequencies", "_frequencies"):
            assert np.array_equal(getattr(index, key), getattr(plain, key)), key
        fwd, rc = hash_read_matrix(reads, 31)
        assert np.array_equal(fwd, plain_hashes[0]) and np.array_equal(rc, plain_hashes[1])
    counter = gki.CounterKmerIndex.from_kmer_index(index)
    counter.count_kmers(fwd.ravel())
    counter.count_kmers(rc.ravel())
    want = gki.CounterKmerIndex.from_kmer_index(plain)
    monkeypatch.setenv("GKI_HOST_COPY_THREADS", "0")
    want.count_kmers(plain_hashes[0].ravel())
    want.count_kmers(plain_hashes[1].ravel())
    assert np.array_equal(counter.get_node_counts(5000), want.get_node_counts(5000)) and counter.get_node_counts(5000).sum() > 0
