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


def test_node_counts_with_more_nodes_than_l2(gki):
    """node ids spread over 12 M (96 MB of float64 counts, more than L2 keeps): same counts as the oracle"""
    from graph_kmer_index_b200 import synthetic
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
