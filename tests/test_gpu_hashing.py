"""GPU parity of K1 (encoding + k-mer hashing) against the oracle, the reference's known-answer vectors and the
golden fixtures.  Everything goes through the product API -> ctypes -> libgki.so."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle, numpy_oracle as no

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gki():
    import graph_kmer_index_b200 as g
    return g


def random_reads(rng, n, L, dirty=True):
    alphabet = np.frombuffer(b"ACGTacgtNnMx", dtype=np.uint8)
    p = np.array([20, 20, 20, 20, 2, 2, 2, 2, 1, 1, .5, .5]) if dirty else np.array([1, 1, 1, 1] + [0] * 8, dtype=float)
    return alphabet[rng.choice(len(alphabet), size=(n, L), p=p / p.sum())]


# ---- the reference's own tests (tests/test_kmer_hashing.py) against the product ------------------------
def test_reference_known_answers(gki):
    from graph_kmer_index_b200.kmer_hashing import (kmer_hash_to_reverse_complement_hash, kmer_hashes_to_bases,
                                                    kmer_hashes_to_reverse_complement_hash, reverse_power_array)
    from graph_kmer_index_b200 import (sequence_to_kmer_hash, letter_sequence_to_numeric, kmer_hash_to_sequence,
                                       numeric_to_letter_sequence)
    assert sequence_to_kmer_hash("ACTG") == 0 * 1 + 1 * 4 + 3 * 16 + 2 * 64
    for s in ["CAtgAACAtttggtAATCTACAtgAACAttt", "ACAtgAACAtttggtAATCTACAtgAACAtt", "CAtgAACAtttggtAATCTACAtgAACAtta"]:
        assert sequence_to_kmer_hash(s) == np.sum(reverse_power_array(31) * letter_sequence_to_numeric(s))
    assert sequence_to_kmer_hash("T" * 31) == 4611686018427387903
    for s in ["atg", "Acacatacgactacg", "CAtgAACAtttggtAATCTACAtgAACAttt", "G"]:
        assert kmer_hash_to_sequence(sequence_to_kmer_hash(s), len(s)).lower() == s.lower()
    comp = str.maketrans("ACGTacgt", "TGCAtgca")
    for s in ["AcATaCAG", "AGACATTA", "GGGGAAAACCCCTTTTAAAACCCCTTTTGGG", "G" * 31, "ACT"]:
        k = len(s)
        h = sequence_to_kmer_hash(s)
        rc = kmer_hash_to_reverse_complement_hash(h, k)
        assert kmer_hash_to_reverse_complement_hash(rc, k) == h
        assert kmer_hash_to_sequence(rc, k).lower() == s[::-1].translate(comp).lower()
    seqs = ["ACACTTACG", "acgactaca", "AATTGGGGG", "ACACACACT"]
    hashes = np.array([sequence_to_kmer_hash(s) for s in seqs])
    assert np.all(kmer_hashes_to_reverse_complement_hash(kmer_hashes_to_reverse_complement_hash(hashes, 9), 9) == hashes)
    bases = kmer_hashes_to_bases(np.array([sequence_to_kmer_hash(s) for s in ["ACTG", "TGGC"]]), 4)
    assert ["".join(numeric_to_letter_sequence(b)).upper() for b in bases] == ["ACTG", "TGGC"]


def test_golden_hashing(gki):
    from graph_kmer_index_b200.read_kmers import hash_read_matrix, hash_ragged_reads
    from graph_kmer_index_b200.kmer_hashing import power_array
    g = load_golden("hashing")
    reads = g["reads"]
    assert np.array_equal(gki.letter_sequence_to_numeric(g["encode_in"]), g["encode_out"])
    assert gki.letter_sequence_to_numeric("ACGTacgtNnMmxyz").dtype == np.uint64
    for k in (1, 3, 5, 16, 31):
        fwd, rc = hash_read_matrix(reads, k)
        assert np.array_equal(fwd, g["fwd_k%d" % k]) and np.array_equal(rc, g["rc_k%d" % k]), k
        strs = [r.tobytes().decode("latin-1") for r in reads]
        f2, r2, off = hash_ragged_reads(strs, k)
        assert np.array_equal(f2.reshape(fwd.shape), fwd) and np.array_equal(r2.reshape(rc.shape), rc), k
        one = gki.ReadKmers.get_kmers_from_read_dynamic(strs[7], power_array(k))
        assert np.array_equal(one, fwd[7]) and one.dtype == np.uint64


def test_golden_revcomp(gki):
    from graph_kmer_index_b200 import kmer_hashing as kh
    g = load_golden("revcomp")
    for k in (1, 2, 4, 9, 16, 30, 31):
        h = g["in_k%d" % k]
        assert np.array_equal(kh.kmer_hashes_to_reverse_complement_hash(h, k), g["rc_k%d" % k]), k
        assert np.array_equal(kh.kmer_hashes_to_complement_hashes(h, k), g["comp_k%d" % k]), k
        assert np.array_equal(kh.kmer_hashes_to_bases(h, k), g["bases_k%d" % k]), k
        assert np.array_equal(kh.kmer_hashes_to_reverse_complement_hash_chunked(h, k, chunk_size=37), g["rc_k%d" % k])
    for s, h, back in zip(g["seqs"], g["seq_hashes"], g["seq_back"]):
        assert kh.sequence_to_kmer_hash(str(s)) == int(h)
        assert kh.kmer_hash_to_sequence(h, len(str(s))) == str(back)
    with pytest.raises(AssertionError):
        kh.kmer_hashes_to_reverse_complement_hash(np.arange(3, dtype=np.uint64), 32)      # kmer_hashing.py:25


@pytest.mark.parametrize("n,L,k", [(5000, 150, 31), (333, 151, 31), (100, 31, 31), (64, 33, 31), (1000, 150, 1),
                                   (257, 100, 21), (40, 64, 16), (3, 250, 31), (1, 150, 31), (1000, 7, 3)])
def test_hash_matrix_vs_oracle(gki, n, L, k):
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    rng = np.random.default_rng(n * 1000 + L + k)
    reads = random_reads(rng, n, L)
    want_f, want_r = c_oracle.hash_reads(reads, k)
    fwd, rc = hash_read_matrix(reads, k)
    assert np.array_equal(fwd, want_f) and np.array_equal(rc, want_r)
    f_only, none = hash_read_matrix(reads, k, reverse=False)
    assert none is None and np.array_equal(f_only, want_f)


def test_hash_layout_variants(gki):
    """strided rows, a base pointer that is not 16-byte aligned (no TMA bulk path), device-resident tensors"""
    import torch
    from graph_kmer_index_b200 import _lib
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    rng = np.random.default_rng(5)
    n, L, k = 700, 150, 31
    reads = random_reads(rng, n, L)
    want_f, want_r = c_oracle.hash_reads(reads, k)
    # unaligned dense
    buf = np.zeros(n * L + 64, dtype=np.uint8)
    for shift in (1, 8, 16):
        view = buf[shift:shift + n * L].reshape(n, L)
        view[:] = reads
        fwd, rc = hash_read_matrix(view, k)
        assert np.array_equal(fwd, want_f) and np.array_equal(rc, want_r), shift
    # strided rows through the C ABI directly
    stride = 160
    padded = np.full((n, stride), ord("T"), dtype=np.uint8)
    padded[:, :L] = reads
    fwd = np.empty((n, L - k + 1), dtype=np.uint64)
    rc = np.empty_like(fwd)
    _lib.call("gki_hash_reads", _lib.ptr(padded), n, L, stride, k, _lib.ptr(fwd), _lib.ptr(rc), None)
    assert np.array_equal(fwd, want_f) and np.array_equal(rc, want_r)
    # device in, device out
    d = torch.from_numpy(reads).cuda()
    dfwd, drc = hash_read_matrix(d, k)
    torch.cuda.synchronize()
    assert np.array_equal(dfwd.cpu().numpy(), want_f) and np.array_equal(drc.cpu().numpy(), want_r)
    # unaligned device view + strided device rows
    dbuf = torch.zeros(n * L + 16, dtype=torch.uint8, device="cuda")
    dbuf[3:3 + n * L] = d.flatten()
    dfwd, drc = hash_read_matrix(dbuf[3:3 + n * L].view(n, L), k)
    torch.cuda.synchronize()
    assert np.array_equal(dfwd.cpu().numpy(), want_f) and np.array_equal(drc.cpu().numpy(), want_r)


@pytest.mark.parametrize("n,L,k", [(1003, 150, 31), (77, 150, 30), (500, 101, 17), (300, 75, 16), (64, 40, 5), (200, 150, 29)])
def test_single_strand_kernels_and_unaligned_outputs(gki, n, L, k):
    """the three instantiations of the hashing kernel (both strands, forward only, reverse complement only) through the C ABI, with
    output rows that do and do not start on 32-byte boundaries (256-bit stores vs the scalar tail), reads with non-ACGT bytes among them"""
    import torch
    from graph_kmer_index_b200 import _lib
    rng = np.random.default_rng(n + L + k)
    reads = random_reads(rng, n, L, dirty=False).copy()     # most reads take the four-windows-per-lane path, one in seven the masked one
    reads[rng.integers(0, n, max(n // 7, 1)), rng.integers(0, L, max(n // 7, 1))] = ord("N")
    want_f, want_r = c_oracle.hash_reads(reads, k)
    nk = L - k + 1
    d = torch.from_numpy(reads).cuda()
    for shift in (0, 1, 3):                                  # uint64 elements: 0 keeps the rows 32-byte aligned when nk % 4 == 0
        f = torch.zeros(n * nk + 4, dtype=torch.int64, device="cuda")
        r = torch.zeros(n * nk + 4, dtype=torch.int64, device="cuda")
        fv, rv = f[shift:shift + n * nk], r[shift:shift + n * nk]
        for use_f, use_r in ((True, True), (True, False), (False, True)):
            f.zero_(), r.zero_()
            _lib.call("gki_hash_reads", _lib.ptr(d), n, L, L, k, _lib.ptr(fv) if use_f else None, _lib.ptr(rv) if use_r else None, None)
            torch.cuda.synchronize()
            got_f, got_r = fv.cpu().numpy().view(np.uint64).reshape(n, nk), rv.cpu().numpy().view(np.uint64).reshape(n, nk)
            assert np.array_equal(got_f, want_f) if use_f else not got_f.any(), (shift, use_f, use_r)
            assert np.array_equal(got_r, want_r) if use_r else not got_r.any(), (shift, use_f, use_r)
            assert int(f[:shift].abs().sum()) == 0 and int(f[shift + n * nk:].abs().sum()) == 0          # nothing outside the rows
            assert int(r[:shift].abs().sum()) == 0 and int(r[shift + n * nk:].abs().sum()) == 0


def test_long_rows_take_the_stream_path(gki):
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    rng = np.random.default_rng(9)
    reads = random_reads(rng, 3, 30011)
    want_f, want_r = c_oracle.hash_reads(reads, 31)
    fwd, rc = hash_read_matrix(reads, 31)
    assert np.array_equal(fwd, want_f) and np.array_equal(rc, want_r)


def test_ragged_and_edge_cases(gki):
    from graph_kmer_index_b200.read_kmers import hash_ragged_reads, hash_read_matrix
    rng = np.random.default_rng(13)
    lens = [0, 1, 30, 31, 32, 150, 5000, 2, 31, 0, 77, 4097, 4096, 4095]
    reads = [random_reads(rng, 1, L)[0].tobytes() if L else b"" for L in lens]
    for k in (31, 5, 1):
        fwd, rc, off = hash_ragged_reads(reads, k)
        for i, r in enumerate(reads):
            a = np.frombuffer(r, dtype=np.uint8)
            if len(a) >= k:
                wf, wr = c_oracle.hash_reads(a[None, :], k)
                wf, wr = wf[0], wr[0]
            else:
                wf = wr = np.zeros(0, dtype=np.uint64)
            assert np.array_equal(fwd[off[i]:off[i + 1]], wf), (k, i)
            assert np.array_equal(rc[off[i]:off[i + 1]], wr), (k, i)
    # empty batch, read shorter than k, all-N reads
    f, r = hash_read_matrix(np.zeros((0, 150), dtype=np.uint8), 31)
    assert f.shape == (0, 120) and r.shape == (0, 120)
    f, r = hash_read_matrix(random_reads(rng, 4, 10), 31)
    assert f.shape == (4, 0)
    f, r = hash_read_matrix(np.full((3, 50), ord("N"), dtype=np.uint8), 31)
    assert not f.any() and not r.any()
    assert gki.ReadKmers.get_kmers_from_read("ACGTACGT", 3) == [int(x) for x in no.read_kmer_hashes("ACGTACGT", 3)[:5]]
    with pytest.raises(Exception):
        hash_read_matrix(random_reads(rng, 2, 40), 32)


def test_fasta_reader(gki, tmp_path):
    """read_kmers.py:14-49: forward k-mers of every read, then those of every reverse-complemented read."""
    rng = np.random.default_rng(3)
    reads = [random_reads(rng, 1, L)[0].tobytes().decode() for L in (60, 45, 80)]
    p = tmp_path / "r.fa"
    p.write_text("".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    got = list(gki.ReadKmers.from_fasta_file(str(p), 31))
    want = [no.read_kmer_hashes(r, 31) for r in reads] + [no.read_kmer_hashes(no.reverse_complement_ascii(r.encode()), 31) for r in reads]
    assert len(got) == 6 and all(np.array_equal(a, b) for a, b in zip(got, want))
    triples = list(gki.ReadKmers.from_fasta_file(str(p), 31, small_k=16, smallest_k=8))
    assert len(triples) == 3
    for r, tr in zip(reads, triples):
        for kk, stream in zip((31, 16, 8), tr):
            want = np.concatenate([no.read_kmer_hashes(r, kk), no.read_kmer_hashes(no.reverse_complement_ascii(r.encode()), kk)])
            assert np.array_equal(np.array(list(stream), dtype=np.uint64), want)


def test_full_size_properties(gki):
    """BASELINE config 1 size (100k x 150 bp): size-independent properties instead of an element-wise oracle."""
    import torch
    from graph_kmer_index_b200 import synthetic
    from graph_kmer_index_b200.kmer_hashing import kmer_hashes_to_reverse_complement_hash
    from graph_kmer_index_b200.read_kmers import hash_read_matrix
    n, L, k = 100_000, 150, 31
    reads = synthetic.reads(n, L, 1_000_000, k, p_hit_permille=100)
    fwd, rc = hash_read_matrix(reads, k)
    # reverse strand == reverse-complement of the forward hashes, reversed (kmer_hashing.py:24-28 vs read_kmers.py:24)
    assert np.array_equal(kmer_hashes_to_reverse_complement_hash(fwd.ravel(), k).reshape(fwd.shape)[:, ::-1], rc)
    # rolling relation: h[i+1] = h[i] >> 2 | base << 60
    code = no.letter_sequence_to_numeric(reads.ravel()).reshape(n, L)
    assert np.array_equal(fwd[:, 1:], (fwd[:, :-1] >> np.uint64(2)) | (code[:, k:] << np.uint64(2 * (k - 1))))
    # sample rows against the oracle
    rows = np.arange(0, n, 997)
    wf, wr = c_oracle.hash_reads(reads[rows], k)
    assert np.array_equal(fwd[rows], wf) and np.array_equal(rc[rows], wr)
