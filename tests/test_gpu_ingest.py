"""GPU parity of the read-ingestion path: 2-bit packed reads through gki_count_packed_reads, and large
host batches through gki_count_reads' packing lanes + copy engine, against the oracle and the plain device path."""
import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gki():
    import graph_kmer_index_b200 as g
    return g


def make_index(gki, n, k, modulo, table_k=None):
    from graph_kmer_index_b200 import synthetic
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 997, k)
    idx = c_oracle.build_index(hashes, nodes, ref, af, modulo, skip_frequencies=True)
    dev = gki.DeviceIndex(idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"], idx["_nodes"], modulo)
    dev.prepare_counting(k if table_k is None else table_k)
    return idx, dev


@pytest.mark.parametrize("n,modulo,n_reads,L,k,table_k", [(40000, 200003, 3001, 150, 31, None), (40000, 4099, 1500, 128, 31, None),
                                                           (5000, 7, 300, 100, 15, None), (40000, 200003, 777, 33, 31, None),
                                                           (3000, 1009, 400, 90, 16, None), (40000, 200003, 900, 150, 31, 30)])
@pytest.mark.parametrize("minimizer_filter", ["0", "1"])
def test_count_packed_reads_vs_oracle(gki, monkeypatch, minimizer_filter, n, modulo, n_reads, L, k, table_k):
    import torch
    monkeypatch.setenv("GKI_FILTER_MZ", minimizer_filter)
    from graph_kmer_index_b200 import synthetic
    from graph_kmer_index_b200.read_kmers import pack_reads
    idx, dev = make_index(gki, n, k, modulo, table_k)
    reads = synthetic.reads(n_reads, L, n, k, p_hit_permille=400, n_permille=10)
    packed, dirty = pack_reads(reads, n_threads=3)
    clean = np.ones(n_reads, dtype=bool)
    clean[dirty] = False
    assert 0 < len(dirty) < n_reads and len(packed) == clean.sum()
    want = c_oracle.read_node_counts(idx, reads[clean], k, 1000)
    assert want.sum() > 0
    dev.count_packed_reads(packed, L, k)                                    # host rows
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    dpacked = torch.from_numpy(packed.view(np.int64)).cuda()
    dev.count_packed_reads(dpacked, L, k)                                   # device rows
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    dev.count_packed_reads(dpacked[1:], L, k)                               # unaligned start: no bulk copies
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads[clean][1:], k, 1000))
    dev.reset_counts()
    dev.count_packed_reads(dpacked, L, k, both_strands=False)
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads[clean], k, 1000, both_strands=False))
    # packed + the dirty rows as ASCII == the whole batch
    dev.reset_counts()
    dev.count_packed_reads(dpacked, L, k)
    dev.count_reads(np.ascontiguousarray(reads[dirty]), k)
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads, k, 1000))
    dev.close()


@pytest.mark.parametrize("threads,mode", [("12", None), ("3", None), ("1", None), ("0", None), ("3", "0"), ("3", "1"), ("3", "2")])
def test_host_pipeline_vs_device_path(gki, monkeypatch, threads, mode):
    """a host batch large enough for the packing lanes: same node counts as the device-resident path and the oracle, with the mode
    chosen adaptively (calls cycle through lanes alone / lanes + copy lane / copy engine alone) or forced by GKI_PIPELINE_DMA"""
    import torch
    from graph_kmer_index_b200 import synthetic
    monkeypatch.setenv("GKI_PACK_THREADS", threads)
    if mode is not None:
        monkeypatch.setenv("GKI_PIPELINE_DMA", mode)
    n, k, L, modulo = 200000, 31, 150, 1000003
    idx, dev = make_index(gki, n, k, modulo)
    n_reads = 5 * 131072 + 777
    reads = synthetic.reads(n_reads, L, n, k, p_hit_permille=300, n_permille=2)
    reads[131072 * 2:131072 * 2 + 30000, 7] = ord("N")          # one chunk with > 1/16 dirty reads: deferred to the ASCII lane
    reads[-1, -1] = ord("n")
    dev.count_reads(torch.from_numpy(reads).cuda(), k)
    want = dev.node_counts(1000)
    sample = slice(0, 20000)
    dev.reset_counts()
    dev.count_reads(reads, k)                                    # host numpy (pageable)
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    pinned = torch.from_numpy(reads).pin_memory()
    dev.count_reads(pinned, k)                                   # pinned host tensor
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    padded = np.full((n_reads, L + 10), ord("A"), dtype=np.uint8)
    padded[:, :L] = reads
    dev.count_reads(padded[:, :L], k, both_strands=False)        # strided rows, forward only
    fwd = dev.node_counts(1000)
    dev.reset_counts()
    dev.count_reads(torch.from_numpy(reads).cuda(), k, both_strands=False)
    assert np.array_equal(fwd, dev.node_counts(1000))
    # the oracle on a slice
    dev.reset_counts()
    dev.count_reads(reads[sample], k)
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads[sample], k, 1000))
    dev.close()


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_count_fastx_vs_oracle(gki, tmp_path, fmt):
    """FASTA / FASTQ file -> node counts: equal to counting the file's sequence lines (as read_kmers.py:16-26 would hash them)
    with the oracle; mixed read lengths, reads shorter than k, N bases, one long group that goes through the packing lanes"""
    from test_host_logic import _write_fastx
    from graph_kmer_index_b200 import synthetic
    n, k, modulo = 100000, 31, 1000003
    idx, dev = make_index(gki, n, k, modulo)
    rng = np.random.default_rng(9)
    big = synthetic.reads(5 * 32768 + 100, 150, n, k, p_hit_permille=300, n_permille=3)
    other = synthetic.reads(3000, 151, n, k, p_hit_permille=300, n_permille=3)
    reads = [row.tobytes().decode() for row in big] + [row.tobytes().decode() for row in other]
    reads += [row[:int(m)].tobytes().decode() for row, m in zip(other[:500], rng.choice([100, 75, 31, 30, 5], size=500))]
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    path = tmp_path / ("reads." + fmt)
    lines = _write_fastx(path, reads, fmt, rng)
    want = np.zeros(1000, dtype=np.float64)
    n_kmers = 0
    by_len = {}
    for line in lines:
        by_len.setdefault(len(line), []).append(line)
    for length, group in by_len.items():
        if length < k:
            continue
        mat = np.frombuffer("".join(group).encode(), dtype=np.uint8).reshape(len(group), length)
        want += c_oracle.read_node_counts(idx, mat, k, 1000)
        n_kmers += 2 * len(group) * (length - k + 1)
    got_kmers = dev.count_fastx(str(path), k)
    assert got_kmers == n_kmers
    assert np.array_equal(dev.node_counts(1000), want)
    # through the reference-shaped class, forward strand only
    counter = gki.CounterKmerIndex(idx["_kmers"].astype(np.int64), idx["_nodes"], gki.collision_free_kmer_index.DeviceCounter(dev))
    counter.reset()
    counter.count_fasta(str(path), k, both_strands=False)
    want_fwd = np.zeros(1000, dtype=np.float64)
    for length, group in by_len.items():
        if length >= k:
            mat = np.frombuffer("".join(group).encode(), dtype=np.uint8).reshape(len(group), length)
            want_fwd += c_oracle.read_node_counts(idx, mat, k, 1000, both_strands=False)
    assert np.array_equal(counter.get_node_counts(1000), want_fwd)
    dev.close()


@pytest.mark.parametrize("L,k,table_k", [(101, 21, None), (250, 31, None), (64, 16, None), (150, 31, 27)])
def test_host_pipeline_other_shapes(gki, monkeypatch, L, k, table_k):
    """the packing lanes with read lengths that are not a multiple of 16 / 32 bases, even k (palindromes), and a table prepared
    for another k (every strand an independent query): equal to the device-resident path"""
    import torch
    from graph_kmer_index_b200 import synthetic
    monkeypatch.setenv("GKI_PACK_THREADS", "3")
    n, modulo = 100000, 1000003
    idx, dev = make_index(gki, n, k, modulo, table_k)
    n_reads = 16 * 32768 + 5
    reads = synthetic.reads(n_reads, L, n, k, p_hit_permille=300, n_permille=2)
    dev.count_reads(torch.from_numpy(reads).cuda(), k)
    want = dev.node_counts(1000)
    assert want.sum() > 0
    dev.reset_counts()
    dev.count_reads(reads, k)
    assert np.array_equal(dev.node_counts(1000), want)
    sample = slice(0, 5000)
    dev.reset_counts()
    dev.count_reads(reads[sample], k)
    assert np.array_equal(dev.node_counts(1000), c_oracle.read_node_counts(idx, reads[sample], k, 1000))
    dev.close()


def test_large_host_packed_batch_is_chunked(gki):
    """a host batch of 2-bit rows above 48 MB crosses the bus in chunks that overlap the count kernels: same counts as the same rows
    resident on the device (which the tests above pin to the oracle)"""
    import torch
    from graph_kmer_index_b200 import _lib, synthetic
    from graph_kmer_index_b200.read_kmers import pack_reads
    n, k, L, modulo = 200000, 31, 150, 1000003
    idx, dev = make_index(gki, n, k, modulo)
    n_reads = 1_400_003
    glen = synthetic.genome_length(n, k)
    genome = torch.empty(glen, dtype=torch.uint8, device="cuda")
    _lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
    d_reads = torch.empty((n_reads, L), dtype=torch.uint8, device="cuda")
    _lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, n_reads, L, 300, 0, _lib.ptr(d_reads), None)
    torch.cuda.synchronize()
    packed, dirty = pack_reads(d_reads.cpu().numpy(), n_threads=4)
    assert len(dirty) == 0 and packed.nbytes > (48 << 20)
    dev.count_packed_reads(torch.from_numpy(packed.view(np.int64)).cuda(), L, k)
    want = dev.node_counts(1000)
    assert want.sum() > 0
    dev.reset_counts()
    dev.count_packed_reads(packed, L, k)                                    # pageable host rows, chunked
    assert np.array_equal(dev.node_counts(1000), want)
    dev.reset_counts()
    pinned = torch.from_numpy(packed.view(np.int64)).pin_memory()
    dev.count_packed_reads(pinned, L, k)                                    # pinned host rows, chunked
    assert np.array_equal(dev.node_counts(1000), want)
    dev.close()
