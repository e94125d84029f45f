"""The smaller methods of the hot-path classes against outputs of the unmodified reference (tests/golden/api_misc.npz from
make_golden_api.py): get_grouped_nodes, get_frequency, the max_hits gate of get, set_frequencies_using_other_index,
FlatKmers.sum_of_kmer_frequencies / maximum_kmer_frequency, kmer_hashes_to_complement_bases, ReadKmers.from_list_of_string_kmers,
convert_kmers_to_complement (against the oracle: the reference's own chunking fails below 10 M k-mers, cfki:473)."""
import numpy as np
import pytest

from conftest import golden_index, load_golden
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gki():
    import graph_kmer_index_b200 as g
    return g


@pytest.fixture(scope="module")
def index(gki):
    g = load_golden("index_small")
    flat = gki.FlatKmers(g["in_hashes"], g["in_nodes"], g["in_ref_offsets"], g["in_allele_frequencies"])
    return gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=int(g["stable_modulo"])), flat, int(g["k"])


def test_frequency_and_grouped_nodes(gki, index):
    idx, flat, k = index
    a = load_golden("api_misc")
    sizes, nodes = a["grouped_sizes"], a["grouped_nodes"]
    s_at = n_at = 0
    for i, q in enumerate(a["queries"]):
        assert idx.get_frequency(int(q), include_reverse_complement=True, k=k) == a["freq_rc"][i]
        assert idx.get_frequency(int(q), include_reverse_complement=False, k=k) == a["freq_fwd"][i]
        assert (idx.get(int(q), max_hits=1)[0] is None) == bool(a["get_max_hits_1_is_none"][i])
        groups = idx.get_grouped_nodes(int(q), max_hits=10 ** 9)
        assert (groups is None) == bool(a["grouped_none"][i])
        n_groups = int(a["grouped_n_groups"][i])
        if groups is not None:
            assert [len(x) for x in groups] == sizes[s_at:s_at + n_groups].tolist()
            got = np.concatenate([np.sort(np.asarray(x)) for x in groups])
            assert np.array_equal(got, nodes[n_at:n_at + len(got)])
            n_at += len(got)
        s_at += n_groups
    small = gki.FlatKmers(flat._hashes[:200], flat._nodes[:200], flat._ref_offsets[:200], flat._allele_frequencies[:200])
    assert small.sum_of_kmer_frequencies(idx) == int(a["sum_of_kmer_frequencies"])
    assert small.maximum_kmer_frequency(idx) == int(a["maximum_kmer_frequency"])


def test_set_frequencies_using_other_index(gki, index):
    idx, _, _ = index
    a = load_golden("api_misc")
    target = idx.copy()
    target._frequencies = target._frequencies.copy()
    target.set_frequencies_using_other_index(idx.copy(), multiplier=3, min_frequency=2)
    assert target._frequencies.dtype == a["set_from_other"].dtype and np.array_equal(target._frequencies, a["set_from_other"])
    removed = idx.copy()
    removed.remove_ref_offsets()
    removed.remove_frequencies()
    assert removed._ref_offsets.tolist() == [0] and removed._frequencies.tolist() == [0]       # cfki:246-250
    before = removed._allele_frequencies.copy()
    removed.set_allele_frequencies(np.arange(3))                                               # a no-op in the reference (cfki:234-235)
    assert np.array_equal(removed._allele_frequencies, before)


def test_complement_bases_and_string_kmers(gki, index):
    _, _, k = index
    from graph_kmer_index_b200.kmer_hashing import kmer_hashes_to_complement_bases
    a = load_golden("api_misc")
    got = kmer_hashes_to_complement_bases(a["queries"][:25], k)
    assert got.dtype == a["complement_bases"].dtype and np.array_equal(got, a["complement_bases"])
    rk = gki.ReadKmers.from_list_of_string_kmers([["ACGTA", "ttgca", "NACGT"], ["GGGGG"], []])
    assert [len(read) for read in rk.kmers] == a["string_kmers_per_read"].tolist()
    assert [int(h) for read in rk.kmers for h in read] == a["string_kmers"].tolist()


def test_convert_kmers_to_complement(gki, index):
    idx, _, k = index
    comp = idx.convert_kmers_to_complement(k=k)
    want = no.build_index(no.complement_hashes(idx._kmers, k), idx._nodes, idx._ref_offsets, idx._allele_frequencies, int(idx._modulo),
                          skip_frequencies=True)
    for key in ("_hashes_to_index", "_n_kmers", "_kmers", "_nodes", "_ref_offsets", "_allele_frequencies", "_frequencies"):
        assert np.array_equal(getattr(comp, key), want[key]), key


def test_bionumpy_hash_and_structural_variant_kmers(gki):
    """Row a10 (bionumpy_wrapper.py:4-10) and its caller, as the reference's tests/test_structural_variants.py:10-49 pins them: the hashes
    equal sequence_to_kmer_hash of every window, and the sampled k-mers index the nodes the reference expects."""
    from graph_kmer_index_b200.bionumpy_wrapper import bionumpy_hash
    from graph_kmer_index_b200.structural_variants import sample_kmers_from_structural_variants
    seq = "GGGGAAAACCCCAAAA"
    got = bionumpy_hash(gki.letter_sequence_to_numeric(seq), 5)
    assert got.dtype == np.uint64 and got.tolist() == [no.sequence_to_kmer_hash(seq[i:i + 5]) for i in range(len(seq) - 4)]
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 4, 5000)
    want = np.convolve(codes.astype(np.uint64), no.power_array(31), mode="valid").astype(np.uint64)       # read_kmers.py:67-70
    assert np.array_equal(bionumpy_hash(codes, 31), want)
    assert len(bionumpy_hash(codes[:4], 5)) == 0

    class DummyGraph:
        def __init__(self, node_sequences):
            self.node_sequences = node_sequences

        def get_numeric_node_sequence(self, node):
            return gki.letter_sequence_to_numeric(self.node_sequences[node])

        def get_node_size(self, node):
            return len(self.node_sequences[node])

    class DummyKmerIndex:
        def get_frequency(self, kmer):
            return 1

    graph = DummyGraph({1: "AAAAAAAAAAA", 2: "ACTG", 3: "GGGGAAAACCCCAAAA", 4: "AGGGG"})
    kmers = sample_kmers_from_structural_variants(graph, zip(np.array([1, 3]), np.array([2, 4])), DummyKmerIndex(), k=5)
    assert kmers._hashes.dtype == np.uint64 and kmers._nodes.dtype == np.uint32
    assert kmers._nodes.tolist() == [1, 1, 3, 3, 3]              # windows 0, 5 of node 1 and 0, 5, 10 of node 3
    index = gki.KmerIndex.from_flat_kmers(kmers)
    assert np.all(index.get_nodes(gki.sequence_to_kmer_hash("AAAAA")) == [1])
    assert np.all(index.get_nodes(gki.sequence_to_kmer_hash("GGGGA")) == [3])
    assert np.all(index.get_nodes(gki.sequence_to_kmer_hash("AAACC")) == [3])


def test_every_counter_has_its_own_zeroed_counts(gki, index):
    """cfki:27: each CounterKmerIndex.from_kmer_index gets a fresh Counter -- two samples counted against one index do not mix"""
    idx, flat, k = index
    first = gki.CounterKmerIndex.from_kmer_index(idx)
    first.count_kmers(idx._kmers[:50])
    a = first.get_node_counts()
    assert a.sum() > 0
    second = gki.CounterKmerIndex.from_kmer_index(idx)              # while `first` is alive: starts from zero, counts separately
    assert second.get_node_counts().sum() == 0
    second.count_kmers(idx._kmers[50:60])
    b = second.get_node_counts()
    assert np.array_equal(first.get_node_counts(), a)
    third = gki.CounterKmerIndex.from_kmer_index(idx)
    third.count_kmers(idx._kmers[:50])
    third.count_kmers(idx._kmers[50:60])
    assert np.array_equal(third.get_node_counts(), a + b)
    del first, second, third
    again = gki.CounterKmerIndex.from_kmer_index(idx)               # nobody holds the cached device copy any more: reused, but reset
    assert again.get_node_counts().sum() == 0


def test_device_copy_follows_edits_of_the_host_arrays(gki):
    """in-place edits + invalidate_device(), and the methods that rebind or edit arrays themselves"""
    g = load_golden("index_small")
    flat = gki.FlatKmers(g["in_hashes"], g["in_nodes"], g["in_ref_offsets"], g["in_allele_frequencies"])
    idx = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=int(g["stable_modulo"]))
    cy = gki.CythonKmerIndex(idx) if hasattr(gki, "CythonKmerIndex") else None
    q = idx._kmers[:200].copy()
    before = idx.device_index().lookup_hits(q, skip_bucket0=False, max_bucket=None, max_frequency=20)
    assert before.shape[1] > 0
    idx._frequencies[:] = 1000                                      # in-place edit: every hit is gated out after invalidation
    idx.invalidate_device()
    after = idx.device_index().lookup_hits(q, skip_bucket0=False, max_bucket=None, max_frequency=20)
    assert after.shape[1] == 0
    idx.set_frequencies()                                           # rebinds and invalidates itself
    again = idx.device_index().lookup_hits(q, skip_bucket0=False, max_bucket=None, max_frequency=20)
    assert np.array_equal(again, before)
    counter = gki.CounterKmerIndex.from_kmer_index(idx)
    counter.count_kmers(q)
    want = counter.get_node_counts()
    idx.remove_frequencies()                                        # a new device copy for the index; the counter keeps its own, with its counts
    assert idx.device_index() is not counter.counter._device
    assert np.array_equal(counter.get_node_counts(), want)
    del cy


def test_kmer_index2_frequencies_with_wide_offsets(gki):
    """start offsets wider than 16 bits stay distinct in count_unique_kmer_occurences (cfki:148-158)"""
    from graph_kmer_index_b200.collision_free_kmer_index import KmerIndex2

    class Flat2:
        pass
    f = Flat2()
    f._hashes = np.array([5, 5, 5, 9, 9], dtype=np.uint64)
    f._nodes = np.array([1, 2, 3, 4, 5], dtype=np.uint32)
    f._start_nodes = np.array([7, 7, 7, 8, 8], dtype=np.int32)
    f._start_offsets = np.array([3, 3 + 65536, 3, 10, 10], dtype=np.int32)
    f._allele_frequencies = np.ones(5, dtype=np.float32)
    idx = KmerIndex2.from_flat_kmers(f, modulo=101)
    assert idx.get_kmer_frequency(5) == 2 and idx.get_kmer_frequency(9) == 1


def test_counts_allreduce_through_the_c_abi_single_rank(gki):
    """gki_nccl_unique_id / gki_nccl_comm_create / gki_allreduce_counts (include/gki.h) with a one-rank communicator: the sum over one
    rank is the vector itself, for both dtypes; bad arguments are refused.  (Two ranks: tests/test_gpu_multi.py.)"""
    import ctypes
    import torch
    from graph_kmer_index_b200 import _lib
    ident = np.zeros(128, dtype=np.uint8)
    _lib.call("gki_nccl_unique_id", ident.ctypes.data)
    assert ident.any()
    comm = ctypes.c_void_p()
    _lib.call("gki_nccl_comm_create", ident.ctypes.data, 0, 1, ctypes.byref(comm))
    assert comm.value
    stream = torch.cuda.current_stream().cuda_stream
    f = torch.arange(100_000, dtype=torch.float64, device="cuda") * 3
    _lib.call("gki_allreduce_counts", comm, f.data_ptr(), f.numel(), _lib.GKI_COUNTS_FLOAT64, stream)
    u = torch.arange(100_000, dtype=torch.int64, device="cuda") + (1 << 40)
    _lib.call("gki_allreduce_counts", comm, u.data_ptr(), u.numel(), _lib.GKI_COUNTS_UINT64, stream)
    torch.cuda.synchronize()
    assert torch.equal(f.cpu(), torch.arange(100_000, dtype=torch.float64) * 3)
    assert torch.equal(u.cpu(), torch.arange(100_000, dtype=torch.int64) + (1 << 40))
    with pytest.raises(_lib.GkiError):
        _lib.call("gki_allreduce_counts", comm, f.data_ptr(), f.numel(), 7, stream)
    host = np.zeros(8)
    with pytest.raises(_lib.GkiError):
        _lib.call("gki_allreduce_counts", comm, host.ctypes.data, 8, _lib.GKI_COUNTS_FLOAT64, stream)
    _lib.call("gki_nccl_comm_destroy", comm)
    _lib.call("gki_release_scratch")


def test_partitioned_counter_index_single_rank(gki):
    """distributed.PartitionedCounterIndex with one rank (no process group): entries and read k-mers all route to rank 0; node counts
    equal the oracle's.  Two ranks: tests/test_gpu_multi.py."""
    import torch
    from graph_kmer_index_b200 import distributed, synthetic
    from oracle import c_oracle
    n, n_nodes, modulo, k = 60_000, 3_000, 200_003, 31
    hashes, nodes, ref, af = synthetic.flat_kmers(n, n_nodes, k)
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    reads = synthetic.reads(5_003, 150, n, k, p_hit_permille=300, n_permille=4)
    want = c_oracle.read_node_counts(idx, reads, k, n_nodes)
    pc = distributed.PartitionedCounterIndex(torch.from_numpy(hashes.view(np.int64)).cuda(), torch.from_numpy(nodes.view(np.int32)).cuda(), modulo)
    pc.count_reads(reads, k, chunk_reads=1000)
    assert np.array_equal(pc.get_node_counts(n_nodes), want) and want.sum() > 0
    pc.reset_counts()
    pc.count_reads(torch.from_numpy(reads).cuda(), k)
    assert np.array_equal(pc.get_node_counts(n_nodes), want)
