"""GPU parity of DenseKmerFinder / CriticalGraphPaths: ordered, element-wise equality with the
unmodified reference on its own test graphs (tests/golden/finder_cases.npz, incl. the 38-row ordered golden of
test_case1) and with the oracle port on larger random variant graphs."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import finder_oracle
from oracle.obgraph_standin import Graph
from test_oracle_finder import load_case

pytestmark = pytest.mark.gpu

KEYS = ("kmers", "nodes", "start_nodes", "start_offsets", "allele_frequencies")


def run_product(arrays, opts):
    import graph_kmer_index_b200 as gki
    finder = gki.DenseKmerFinder(arrays, opts["k"], max_variant_nodes=opts["max_variant_nodes"],
                                 only_save_one_node_per_kmer=opts["only_save_one_node_per_kmer"], only_store_nodes=opts["only_store_nodes"],
                                 only_follow_nodes=opts.get("only_follow_nodes"))
    if opts["only_position"] is None:
        finder.find()
    else:
        finder.find_only_kmers_starting_at_position(*opts["only_position"])
    return finder


def test_reference_fixtures():
    import graph_kmer_index_b200 as gki
    g = load_golden("finder_cases")
    for i in range(int(g["n_cases"])):
        arrays, opts, ref, crit = load_case(g, i)
        c = gki.CriticalGraphPaths.from_graph(arrays, opts["k"])
        assert np.array_equal(c.nodes, crit[0]) and np.array_equal(c.offsets, crit[1]), i
        assert c.nodes.dtype == np.uint32 and c.offsets.dtype == np.uint16
        finder = run_product(arrays, opts)
        for key in KEYS:
            got = finder._results[key]
            assert got.dtype == ref[key].dtype or key in ("nodes", "kmers"), (i, key, got.dtype, ref[key].dtype)
            assert np.array_equal(got, ref[key]), (i, key, got[:12], ref[key][:12])


def test_only_follow_nodes_reference_fixtures():
    """kf:385-388: ordered equality with the unmodified reference (tests/golden/finder_follow.npz)."""
    g = load_golden("finder_follow")
    for i in range(int(g["n_cases"])):
        arrays, opts, ref, _ = load_case(g, i)
        finder = run_product(arrays, opts)
        for key in KEYS:
            assert np.array_equal(finder._results[key], ref[key]), (i, key, finder._results[key][:12], ref[key][:12])


def test_reference_test_case1_order():
    """tests/test_kmer_finder.py:412-475 of the reference: the exact ordered (k-mer, node) list"""
    import graph_kmer_index_b200 as gki
    graph = Graph.from_dicts({0: "AGTAGA", 1: "G", 2: "CT", 3: "ACTA", 5: "G", 6: "A", 7: "TCATA"},
                             {0: [1, 2], 1: [3], 2: [3], 3: [5, 6], 5: [7], 6: [7], 7: []}, [0, 1, 3, 5, 7])
    finder = gki.DenseKmerFinder(graph, k=3)
    finder.find()
    kmers, nodes = finder.get_found_kmers_and_nodes()
    correct = [("AGT", 0), ("GTA", 0), ("TAG", 0), ("AGA", 0), ("GAG", 0), ("GAG", 1), ("AGA", 0), ("AGA", 1), ("AGA", 3), ("GAC", 1), ("GAC", 3),
               ("GAC", 0), ("GAC", 2), ("ACT", 0), ("ACT", 2), ("CTA", 2), ("CTA", 3), ("TAC", 2), ("TAC", 3), ("ACT", 3), ("CTA", 3), ("TAG", 3),
               ("TAG", 5), ("AGT", 3), ("AGT", 5), ("AGT", 7), ("GTC", 5), ("GTC", 7), ("TAA", 3), ("TAA", 6), ("AAT", 3), ("AAT", 6), ("AAT", 7),
               ("ATC", 6), ("ATC", 7), ("TCA", 7), ("CAT", 7), ("ATA", 7)]
    assert len(kmers) == len(correct)
    for kmer, node, (seq, want_node) in zip(kmers, nodes, correct):
        assert gki.kmer_hash_to_sequence(kmer, 3).upper() == seq and node == want_node
    flat = finder.get_flat_kmers()
    assert len(flat._hashes) == 38 and flat._start_nodes.dtype == np.int32
    flat0 = finder.get_flat_kmers(v="0")
    assert len(flat0._ref_offsets) == 38


@pytest.mark.parametrize("seed,n_variants,spacing,k,kw", [
    (1, 300, 40, 31, dict(max_variant_nodes=4)),
    (2, 500, 12, 11, dict(max_variant_nodes=12, p_nested=0.2)),
    (3, 200, 300, 31, dict(max_variant_nodes=5, only_save_one_node_per_kmer=True)),
    (4, 400, 6, 5, dict(max_variant_nodes=12, p_nested=0.3)),
    (6, 400, 8, 7, dict(max_variant_nodes=1)),
    (5, 50, 2000, 31, dict(max_variant_nodes=5)),          # long nodes: the _process_whole_node rows
])
def test_random_graphs_vs_oracle(seed, n_variants, spacing, k, kw):
    from graph_kmer_index_b200 import synthetic
    p_nested = kw.pop("p_nested", 0.0)
    seqs, edges, linear, af = synthetic.variant_graph(n_variants, spacing=spacing, seed=seed, p_deletion=0.25, p_nested=p_nested)
    arrays = Graph.from_dicts(seqs, edges, linear, af).to_arrays()
    opts = dict(k=k, max_variant_nodes=kw.get("max_variant_nodes", 4), only_save_one_node_per_kmer=kw.get("only_save_one_node_per_kmer", False),
                only_store_nodes=None, only_position=None)
    try:
        want = finder_oracle.dense_kmer_finder(arrays, **opts)
    except (AssertionError, OverflowError):
        pytest.skip("the reference algorithm itself rejects this graph")
    finder = run_product(arrays, opts)
    assert len(want["kmers"]) > 1000
    for key in KEYS:
        assert np.array_equal(finder._results[key], want[key]), (key, len(finder._results[key]), len(want[key]))


def load_chunk_case(g, i):
    p = "c%d_" % i
    arrays = {key[len(p) + 2:]: g[key] for key in g.files if key.startswith(p + "g_")}
    chunks = [tuple(int(x) for x in row) for row in g[p + "chunks"]]
    results = [{key: g[p + "r%d_%s" % (j, key)] for key in KEYS} for j in range(len(chunks))]
    return arrays, int(g[p + "k"]), int(g[p + "max_variant_nodes"]), (g[p + "crit_nodes"], g[p + "crit_offsets"]), chunks, results


def test_chunked_runs_match_the_reference():
    """start/stop_at_critical_path_number (kf:186-226) per chunk of critical paths, the way `graph_kmer_index index -t T` runs the
    finder (cli:588-608): ordered equality with the unmodified reference for every chunk (tests/golden/finder_chunks.npz), and the
    sharded helper returns the chunks concatenated in order whatever rank computed them"""
    import graph_kmer_index_b200 as gki
    from graph_kmer_index_b200 import distributed
    g = load_golden("finder_chunks")
    assert int(g["n_cases"]) >= 20
    for i in range(int(g["n_cases"])):
        arrays, k, mvn, (cn, co), chunks, results = load_chunk_case(g, i)
        crit = gki.CriticalGraphPaths(cn, co)
        requested = [n for n in (2, 3, 5) if distributed.critical_path_chunks(len(crit), n) == chunks]
        assert requested, (i, chunks)                        # the helper cuts chunks like cli:588-603 (the fixture's own cutter)
        for (s, e), ref in zip(chunks, results):
            finder = gki.DenseKmerFinder(arrays, k, critical_graph_paths=gki.CriticalGraphPaths(cn, co), max_variant_nodes=mvn,
                                         start_at_critical_path_number=s, stop_at_critical_path_number=e)
            finder.find()
            for key in KEYS:
                assert np.array_equal(finder._results[key], ref[key]), (i, (s, e), key)
        wl_finder = gki.DenseKmerFinder(arrays, k, critical_graph_paths=gki.CriticalGraphPaths(cn, co), max_variant_nodes=mvn,
                                        whitelist=set(int(x) for x in g["c%d_whitelist" % i]))
        wl_finder.find()
        for key in KEYS:                                     # whitelist (kf:130-132, 362-365) against the reference
            assert np.array_equal(wl_finder._results[key], g["c%d_wl_%s" % (i, key)]), (i, "whitelist", key)
        # two emulated ranks, chunks dealt round-robin, gathered by hand
        n_chunks = requested[0]
        parts = []
        for rank in range(2):
            parts += distributed.find_kmers_sharded(arrays, k, critical_graph_paths=gki.CriticalGraphPaths(cn, co), n_chunks=n_chunks, rank=rank,
                                                    world_size=2, gather=False, max_variant_nodes=mvn)
        parts.sort(key=lambda item: item[0])
        assert [c for c, _ in parts] == list(range(len(chunks)))
        assert np.array_equal(np.concatenate([np.asarray(f._hashes) for _, f in parts]).astype(np.int64), np.concatenate([r["kmers"] for r in results]))
        assert np.array_equal(np.concatenate([f._nodes for _, f in parts]), np.concatenate([r["nodes"] for r in results]))
        whole = distributed.find_kmers_sharded(arrays, k, critical_graph_paths=gki.CriticalGraphPaths(cn, co), n_chunks=n_chunks, rank=0, world_size=1,
                                               max_variant_nodes=mvn)
        assert len(whole._hashes) == sum(len(f._hashes) for _, f in parts)


def test_kmer_index2_after_the_reference_tests():
    """KmerIndex2 / MultiValueHashTable (cfki:110-158, multi_value_hashtable.py) as the reference's own tests pin them:
    tests/test_indexes2.py:6-22 and the KmerIndex2 assertions of tests/test_kmer_finder.py:10-80."""
    import graph_kmer_index_b200 as gki
    flat = gki.FlatKmers2(np.array([1, 1, 1, 2, 3, 10, 11, 2]), np.array([1, 1, 2, 2, 3, 1, 10, 5]), np.array([0, 0, 1, 2, 3, 4, 5, 6]),
                          np.array([1, 2, 3, 4, 5, 6, 7, 8]), np.array([0.4, 0.1, 0.3, 0.4, 0.1, 0.1, 0.1, 0.1]))
    index = gki.KmerIndex2.from_flat_kmers(flat)
    assert index.get_kmer_frequency(1) == 2
    assert np.all(index.get_start_nodes(1) == [1, 1, 2])
    assert np.all(index.get_nodes(3) == [5])
    assert np.all(index.get_nodes(2) == [4, 8]) and np.all(index.get_start_offsets(2) == [2, 6]) and index.get_kmer_frequency(2) == 2
    assert np.all(index._data[1]["allele_frequencies"] == [0.4, 0.1, 0.3])
    assert len(index.get_nodes(7)) == 0 and sorted(index.get_all_kmers()) == sorted(flat._hashes)

    h = gki.sequence_to_kmer_hash
    graph = Graph.from_dicts({0: "AAA", 1: "C", 2: "T", 3: "AAA"}, {0: [1, 2], 2: [3], 1: [3]}, [0, 1, 3])
    finder = gki.DenseKmerFinder(graph, k=3)
    finder.find()
    index = gki.KmerIndex2.from_flat_kmers(finder.get_flat_kmers(), modulo=15)
    assert np.all(index.get_nodes(h("ATA")) == [0, 2, 3])
    assert np.all(index.get_start_nodes(h("ATA")) == [3, 3, 3])
    assert np.all(index.get_start_offsets(h("ATA")) == [0, 0, 0])
    assert set(index.get_nodes(h("ACA"))) == {0, 1, 3} and set(index.get_nodes(h("AAA"))) == {0, 3}
    assert len(index.get_all_kmers()) == 16

    graph = Graph.from_dicts({0: "ACTGACTG", 1: "A", 2: "T", 3: "AAAAA", 4: "C", 5: "T", 6: "TGGGGG"},
                             {0: [1, 2], 2: [3], 1: [3], 3: [4, 5], 4: [6], 5: [6]}, [0, 1, 3, 4, 6])
    finder = gki.DenseKmerFinder(graph, k=3)
    finder.find()
    index = gki.KmerIndex2.from_flat_kmers(finder.get_flat_kmers())
    assert set(index.get_nodes(h("ACT"))) == {0, 3, 4, 6}
    assert set(index.get_start_nodes(h("AAC"))) == {4} and set(index.get_start_offsets(h("AAC"))) == {0}

    graph = Graph.from_dicts({1: "ATC", 2: "AAAAAAAA", 3: "T", 4: "CTA"}, {1: [2, 3], 2: [4], 3: [4]}, [1, 2, 4])
    finder = gki.DenseKmerFinder(graph, k=3)
    finder.find()
    flat = finder.get_flat_kmers()
    index = gki.KmerIndex2.from_flat_kmers(flat)
    assert len(index.get_nodes(h("AAA"))) == 6 and len(index.get_nodes(h("AAC"))) == 2
    # frequencies against a direct count over the finder's rows
    for kmer in np.unique(flat._hashes):
        rows = flat._hashes == kmer
        assert index.get_kmer_frequency(kmer) == len(set(zip(flat._start_nodes[rows].tolist(), flat._start_offsets[rows].tolist())))
        assert np.array_equal(index.get_nodes(kmer), flat._nodes[rows])            # rows of a key come back in input order


def test_config5_chunk_digests():
    """BASELINE config 5 at full size (100 k variants, 30 Mbp, k=31, max_variant_nodes=5): every critical-path chunk of the device
    finder has the row count and the SHA-256 of the ordered rows that the UNMODIFIED reference produced for that chunk
    (tests/golden/finder_c5.json from make_golden_c5.py, chunked like `graph_kmer_index index -t T`, cli:588-608)."""
    import hashlib
    import json
    import os
    import graph_kmer_index_b200 as gki
    from graph_kmer_index_b200 import synthetic
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "finder_c5.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/finder_c5.json not generated")
    g = json.load(open(path))
    seqs, edges, linear, af = synthetic.variant_graph(g["n_variants"], spacing=g["spacing"], seed=g["seed"], p_deletion=g["p_deletion"])
    arrays = Graph.from_dicts(seqs, edges, linear, af).to_arrays()
    crit = gki.CriticalGraphPaths.from_graph(arrays, g["k"])
    assert len(crit) == g["n_critical_paths"]
    total = 0
    for (s, e), want_rows, want_sha in zip(g["chunks"], g["rows"], g["sha256"]):
        finder = gki.DenseKmerFinder(arrays, g["k"], critical_graph_paths=crit, max_variant_nodes=g["max_variant_nodes"],
                                     start_at_critical_path_number=s, stop_at_critical_path_number=e)
        finder.find()
        r = finder._results
        h = hashlib.sha256()
        for key, dtype in (("kmers", np.int64), ("nodes", np.int32), ("start_nodes", np.int32), ("start_offsets", np.int16), ("allele_frequencies", np.float64)):
            a = np.ascontiguousarray(r[key])
            assert a.dtype == dtype, (key, a.dtype)
            h.update(a.tobytes())
        assert len(r["kmers"]) == want_rows, (s, e, len(r["kmers"]), want_rows)
        assert h.hexdigest() == want_sha, (s, e)
        total += want_rows
    assert total == g["total_rows"]
