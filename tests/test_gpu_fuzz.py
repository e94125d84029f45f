"""Seeded random shapes through the fused counting path (ASCII, packed, host and device buffers, every filter addressing) against
the C oracle: read lengths from k up to a few hundred bases, every k, odd batch sizes, N / lower-case rates from 0 to 30 %."""
import os

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu
EXTRA = int(os.environ.get("GKI_FUZZ_EXTRA", "0"))       # more seeds for a one-off longer run


def random_case(seed):
    rng = np.random.default_rng(seed)
    k = int(rng.choice([1, 2, 5, 8, 15, 16, 20, 27, 28, 29, 30, 31]))
    L = int(rng.choice([k, k + 1, k + 3, 2 * k + 1, 64, 100, 101, 127, 128, 150, 151, 200, 333]))
    L = max(L, k)
    n_reads = int(rng.choice([1, 2, 7, 8, 9, 31, 33, 255, 1000, 4097]))
    n = int(rng.choice([3000, 20000]))
    modulo = int(rng.choice([7, 1009, 200003]))
    return k, L, n_reads, n, modulo, int(rng.choice([0, 10, 300])), bool(rng.integers(0, 2)), int(rng.integers(0, 3))


@pytest.mark.parametrize("seed", range(48 + EXTRA))
def test_random_shapes(monkeypatch, seed):
    import torch
    import graph_kmer_index_b200 as gki
    from graph_kmer_index_b200 import synthetic
    from graph_kmer_index_b200.read_kmers import pack_reads
    k, L, n_reads, n, modulo, n_permille, both, filter_mode = random_case(seed)
    if filter_mode == 1:
        monkeypatch.setenv("GKI_FILTER_MZ", "1")             # minimizer-addressed filter (takes effect for odd k in 27..31)
    elif filter_mode == 2:
        monkeypatch.setenv("GKI_FILTER_MAX_MB", "0")         # no filter
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 997, k)
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    reads = synthetic.reads(n_reads, L, n, k, p_hit_permille=400, n_permille=n_permille)
    want = c_oracle.read_node_counts(idx, reads, k, 1000, both_strands=both)
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], modulo)
    dev.prepare_counting(k)
    case = (k, L, n_reads, n, modulo, n_permille, both, filter_mode)
    dev.count_reads(reads, k, both)                                            # host rows
    assert np.array_equal(dev.node_counts(1000), want), ("host", case)
    dev.reset_counts()
    dev.count_reads(torch.from_numpy(reads).cuda(), k, both)                   # device rows
    assert np.array_equal(dev.node_counts(1000), want), ("device", case)
    packed, dirty = pack_reads(reads, n_threads=2)
    dev.reset_counts()
    if len(packed):
        dev.count_packed_reads(torch.from_numpy(packed.view(np.int64)).cuda(), L, k, both)
    if len(dirty):
        dev.count_reads(np.ascontiguousarray(reads[dirty]), k, both)
    assert np.array_equal(dev.node_counts(1000), want), ("packed + dirty", case)
    dev.close()


@pytest.mark.parametrize("seed", range(20 + EXTRA // 4))
def test_random_builds(seed):
    """gki_index_build on random sizes / table sizes / repeat patterns (binned path, its big-bin variant and the radix fallback are
    all reached) against the C oracle"""
    import graph_kmer_index_b200 as gki
    from graph_kmer_index_b200 import synthetic
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(33000, 300000))
    modulo = int(rng.choice([257, 4099, 65521, 300007, 1000003, 19999999, 50000017]))
    dup = int(rng.choice([1, 2, 8, 40, 300]))
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 1000, 31)
    if dup > 1:
        hashes = hashes.copy()
        pick = rng.random(n) < 0.4
        hashes[pick] = hashes[(np.arange(n) // dup * dup)[pick]]
        ref = (ref // np.uint64(2)).astype(np.uint64)
    skip = bool(modulo < 60000 or rng.integers(0, 2))          # (the oracle's frequency pass is quadratic in the bucket size)
    want = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=skip)
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(hashes, nodes, ref, af), modulo=modulo, skip_frequencies=skip)
    for key in ("_hashes_to_index", "_n_kmers", "_kmers", "_nodes", "_ref_offsets", "_allele_frequencies", "_frequencies"):
        assert np.array_equal(getattr(index, key), want[key]), (key, n, modulo, dup, skip)
