"""GPU parity of ReverseKmerIndex / ReferenceKmerIndex (gki_group_by_key) against the fixtures generated from the
unmodified reference (tests/golden/make_golden_side_indexes.py), the numpy oracle on larger seeded inputs, and the
reference's own test (tests/test_reverse_kmer_index.py)."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gki():
    import graph_kmer_index_b200 as g
    return g


def same(got, want, what):
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    assert np.array_equal(got, want), what


def test_reference_reverse_index_test(gki, tmp_path):
    """tests/test_reverse_kmer_index.py:5-19."""
    flat = gki.FlatKmers(np.array([10, 3, 11, 4]), np.array([5, 3, 5, 8]))
    reverse = gki.ReverseKmerIndex.from_flat_kmers(flat)
    assert 11 in reverse.get_node_kmers(5)
    assert 10 in reverse.get_node_kmers(5)
    assert 3 in reverse.get_node_kmers(3)
    assert 4 in reverse.get_node_kmers(8)
    reverse.to_file(str(tmp_path / "tmp.reverse"))
    new_reverse = gki.ReverseKmerIndex.from_file(str(tmp_path / "tmp.reverse.npz"))
    assert 3 in new_reverse.get_node_kmers(3)
    assert list(new_reverse.get_node_kmers(4)) == []


def test_side_indexes_golden(gki, tmp_path):
    g = load_golden("side_indexes")
    for name in g["names"]:
        flat = gki.FlatKmers(g[name + "/hashes"], g[name + "/nodes"], g[name + "/ref_offsets"])
        rev = gki.ReverseKmerIndex.from_flat_kmers(flat)
        same(rev.nodes_to_index_positions, g[name + "/rev_nodes_to_index_positions"], (name, "nodes_to_index_positions"))
        same(rev.nodes_to_n_hashes, g[name + "/rev_nodes_to_n_hashes"], (name, "nodes_to_n_hashes"))
        same(rev.hashes, g[name + "/rev_hashes"], (name, "hashes"))
        same(rev.ref_positions, g[name + "/rev_ref_positions"], (name, "ref_positions"))
        ref = gki.ReferenceKmerIndex.from_flat_kmers(flat)
        same(ref.ref_position_to_index, g[name + "/ref_ref_position_to_index"], (name, "ref_position_to_index"))
        same(ref.kmers, g[name + "/ref_kmers"], (name, "kmers"))
        same(ref.ref_positions, g[name + "/ref_ref_positions"], (name, "ref_positions"))
        same(ref.nodes, g[name + "/ref_nodes"], (name, "nodes"))
    # npz round trip (reference_kmer_index.py:123-160)
    ref.to_file(str(tmp_path / "refindex"))
    back = gki.ReferenceKmerIndex.from_file(str(tmp_path / "refindex"))
    same(back.ref_position_to_index, ref.ref_position_to_index, "round trip")
    assert list(back.get_between(1, 4)) == list(ref.kmers[ref.ref_position_to_index[1]:ref.ref_position_to_index[4]])


def test_reference_index_from_sequence(gki):
    g = load_golden("side_indexes")
    seq = g["seq"].tobytes().decode()
    for k in (5, 16, 31):
        idx = gki.ReferenceKmerIndex.from_sequence(seq, k)
        same(idx.kmers, g["seq_k%d_kmers" % k], ("kmers", k))
        same(idx.ref_position_to_index, g["seq_k%d_index" % k], ("index", k))
    only = gki.ReferenceKmerIndex.from_sequence(seq, 31, only_store_kmers=True)
    assert only.ref_position_to_index is None
    same(only.kmers, g["seq_k31_kmers"], "only_store_kmers")


@pytest.mark.parametrize("n,n_nodes,max_ref", [(200_000, 30_000, 150_000), (1_000_000, 5, 4_000_000_000), (300_000, 250_000, 40)])
def test_side_indexes_vs_oracle(gki, n, n_nodes, max_ref):
    rng = np.random.default_rng(n + n_nodes)
    hashes = rng.integers(0, 2 ** 62, n, dtype=np.uint64)
    nodes = rng.integers(0, n_nodes, n).astype(np.uint32)
    if max_ref > 2 ** 31:        # sparse positions, long gaps, values above 2^31
        ref = np.sort(rng.choice(np.arange(0, max_ref, 4001, dtype=np.uint64), size=n // 50))[rng.integers(0, n // 50, n)]
        ref = ref // 1000       # keep the table small enough for the oracle's Python loop
    else:
        ref = rng.integers(0, max_ref, n).astype(np.uint64)
    flat = gki.FlatKmers(hashes, nodes, ref)
    rev = gki.ReverseKmerIndex.from_flat_kmers(flat)
    for got, want in zip((rev.nodes_to_index_positions, rev.nodes_to_n_hashes, rev.hashes, rev.ref_positions), no.reverse_index(hashes, nodes, ref)):
        same(got, want, "reverse")
    refi = gki.ReferenceKmerIndex.from_flat_kmers(flat)
    for got, want in zip((refi.ref_position_to_index, refi.kmers, refi.ref_positions, refi.nodes), no.reference_index(hashes, nodes, ref)):
        same(got, want, "reference")


def test_group_by_key_properties(gki):
    """Full-size properties: the permutation is a stable sort, first/count describe the runs."""
    from graph_kmer_index_b200.reverse_kmer_index import group_by_key
    rng = np.random.default_rng(3)
    n, n_keys = 20_000_000, 3_000_000
    keys = rng.integers(0, n_keys, n).astype(np.uint32)
    perm, first, counts = group_by_key(keys, n_keys)
    s = keys[perm]
    assert np.all(s[1:] >= s[:-1])
    same_key = s[1:] == s[:-1]
    assert np.all(perm[1:][same_key] > perm[:-1][same_key])          # stable
    assert np.array_equal(counts, np.bincount(keys, minlength=n_keys).astype(np.uint32))
    present = counts > 0
    assert np.array_equal(first[present], (np.cumsum(counts) - counts)[present].astype(np.uint32))
    assert not first[~present].any()


def test_group_by_key_errors(gki):
    from graph_kmer_index_b200 import _lib
    from graph_kmer_index_b200.reverse_kmer_index import group_by_key
    with pytest.raises(_lib.GkiError):
        group_by_key(np.array([1, 2, 9], dtype=np.uint32), 5)            # key >= n_keys
    with pytest.raises(_lib.GkiError):
        group_by_key(np.zeros(0, dtype=np.uint32), 5)                    # empty (np.max of an empty array raises in the reference)


def test_cli_subcommands(gki, tmp_path):
    """make_from_flat / add_reverse_complements / make_reverse / make_reference_kmer_index (cli:156-193, 656-667) write the
    files the reference's classes would: checked against the golden index and the oracle"""
    from conftest import golden_index
    from graph_kmer_index_b200.command_line_interface import run_argument_parser
    g = load_golden("index_small")
    flat = gki.FlatKmers(g["in_hashes"], g["in_nodes"], g["in_ref_offsets"], g["in_allele_frequencies"])
    flat.to_file(str(tmp_path / "flat"))
    want = golden_index(g)
    run_argument_parser(["make_from_flat", "-f", str(tmp_path / "flat.npz"), "-o", str(tmp_path / "index"), "-m", str(want["_modulo"])])
    index = gki.CollisionFreeKmerIndex.from_file(str(tmp_path / "index"))
    for key in ("_hashes_to_index", "_n_kmers", "_kmers", "_nodes", "_ref_offsets", "_frequencies", "_allele_frequencies"):
        assert np.array_equal(getattr(index, key), want[key]), key
    k = int(g["k"])
    run_argument_parser(["add_reverse_complements", "-f", str(tmp_path / "flat.npz"), "-o", str(tmp_path / "flat_rc"), "-k", str(k)])
    both = gki.FlatKmers.from_file(str(tmp_path / "flat_rc.npz"))
    hashes = np.asarray(flat._hashes).astype(np.uint64)
    assert np.array_equal(both._hashes, np.concatenate([hashes, no.revcomp_hashes(hashes, k)]))
    assert np.array_equal(both._nodes, np.concatenate([flat._nodes, flat._nodes]))
    run_argument_parser(["make_reverse", "-f", str(tmp_path / "flat.npz"), "-o", str(tmp_path / "reverse")])
    rev = gki.ReverseKmerIndex.from_file(str(tmp_path / "reverse.npz"))
    for got, wanted in zip((rev.nodes_to_index_positions, rev.nodes_to_n_hashes, rev.hashes, rev.ref_positions),
                           no.reverse_index(flat._hashes, flat._nodes, flat._ref_offsets)):
        same(got, wanted, "make_reverse")
    run_argument_parser(["make_reference_kmer_index", "-f", str(tmp_path / "flat.npz"), "-o", str(tmp_path / "refidx")])
    refi = gki.ReferenceKmerIndex.from_file(str(tmp_path / "refidx"))
    for got, wanted in zip((refi.ref_position_to_index, refi.kmers, refi.ref_positions, refi.nodes),
                           no.reference_index(flat._hashes, flat._nodes, flat._ref_offsets)):
        same(got, wanted, "make_reference_kmer_index")
    fasta = tmp_path / "ref.fa"
    fasta.write_text(">other\nACGT\n>ref some description\nACGTTGCAAC\nGGTTAACC\n")
    run_argument_parser(["make_reference_kmer_index", "-r", str(fasta), "-n", "ref", "-k", "5", "-o", str(tmp_path / "linear")])
    lin = gki.ReferenceKmerIndex.from_file(str(tmp_path / "linear"))
    same(lin.kmers, no.read_kmer_hashes("ACGTTGCAACGGTTAACC", 5).astype(np.uint32), "linear reference")


def test_kmer_counter_from_kmers(gki):
    """kmer_counter.py:33-43 (np.unique(kmers, return_counts=True) behind a key -> count table)"""
    from graph_kmer_index_b200.kmer_counter import KmerCounter
    rng = np.random.default_rng(12)
    kmers = rng.integers(0, 5000, 200000).astype(np.uint64) * np.uint64(1000003) + np.uint64(7)
    kmers[:70000] = kmers[0]                                    # one k-mer more often than a uint16 holds
    counter = KmerCounter.from_kmers(kmers, 0)
    unique, counts = np.unique(kmers, return_counts=True)
    assert np.array_equal(counter.counter[unique], counts)
    assert counter.get_frequency(int(unique[3]))[0] == counts[3]
    assert len(counter.get_frequency(12345)) == 0 and counter.counter[np.array([12345, int(unique[0])])].tolist() == [0, int(counts[0])]
    assert counter.score_kmers([int(unique[1]), 999]) == -int(counts[1]) and counter.score_kmers([999]) == 1
    flat = gki.FlatKmers(kmers, np.zeros(len(kmers), dtype=np.uint32))
    sub = KmerCounter.from_flat_kmersv2(flat, 19999999, subsample_ratio=3)
    u3, c3 = np.unique(kmers[::3], return_counts=True)
    assert np.array_equal(sub.counter[u3], c3)


def test_multi_value_hashtable_reference_test(gki):
    """The reference's tests/test_multi_value_hashtable.py:5-8, verbatim expectations."""
    from graph_kmer_index_b200.multi_value_hashtable import MultiValueHashTable
    h = MultiValueHashTable.from_keys_and_values([1, 2, 3, 1], {"nodes": np.array([1, 2, 3, 10]), "offsets": np.array([5, 3, 2, 100])}, mod=11)
    assert np.all(h[1]["nodes"] == [1, 10])
    assert np.all(h[2]["offsets"] == [3])
    assert sorted(h.get_all_keys().tolist()) == [1, 1, 2, 3] and h.get_unique_keys().tolist() == [1, 2, 3]
