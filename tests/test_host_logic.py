"""CPU tests of the host-side logic: synthetic generators, sharding, the world_size-2 all-reduce path (gloo)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from graph_kmer_index_b200 import distributed, synthetic
from oracle import c_oracle, numpy_oracle as no


def test_synthetic_is_deterministic_and_shardable():
    h1 = synthetic.flat_kmers(1001, 50, 31)
    h2 = synthetic.flat_kmers(1001, 50, 31)
    for a, b in zip(h1, h2):
        assert np.array_equal(a, b)
    assert h1[0].dtype == np.uint64 and h1[1].dtype == np.uint32 and h1[2].dtype == np.uint64 and h1[3].dtype == np.float32
    # every unique k-mer appears twice (once for the odd tail), with the same ref offset
    u, c = np.unique(h1[2], return_counts=True)
    assert len(u) == 501 and set(c) == {1, 2}
    r = synthetic.reads(64, 150, 1001, 31, p_hit_permille=500)
    parts = np.concatenate([synthetic.reads(20, 150, 1001, 31, 500, first_read=0),
                            synthetic.reads(44, 150, 1001, 31, 500, first_read=20)])
    assert np.array_equal(r, parts)
    assert set(np.unique(r)) <= set(b"ACGT")
    assert ord("N") in synthetic.reads(64, 150, 1001, 31, 500, n_permille=100)


def test_synthetic_reads_hit_the_index():
    n = 4000
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 100, 31)
    idx = no.build_index(hashes, nodes, ref, af, 10007, skip_frequencies=True)
    r = synthetic.reads(200, 150, n, 31, p_hit_permille=1000)
    fwd, rc = no.hash_reads(r, 31)
    hit = no.has_kmers(idx, np.concatenate([fwd.ravel(), rc.ravel()]))
    assert 0.45 < hit.mean() < 0.55          # one strand of every read comes from the indexed sequence


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 12345):
        for w in (1, 2, 3, 8):
            b = [distributed.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from graph_kmer_index_b200 import distributed, synthetic
    from oracle import c_oracle, numpy_oracle as no
    rank, world = distributed.init_process_group("gloo")
    n = 3000
    hashes, nodes, ref, af = synthetic.flat_kmers(n, 200, 31)
    idx = no.build_index(hashes, nodes, ref, af, 5003, skip_frequencies=True)
    reads = synthetic.reads(400, 100, n, 31, p_hit_permille=600)
    lo, hi = distributed.shard_bounds(len(reads), rank, world)
    local = c_oracle.read_node_counts(idx, reads[lo:hi], 31, 200)       # checker stands in for the GPU count
    t = torch.from_numpy(local.copy())
    distributed.allreduce_node_counts(t)
    whole = c_oracle.read_node_counts(idx, reads, 31, 200)
    assert np.array_equal(t.numpy(), whole), rank
    assert whole.sum() > 0
    dist.barrier()
    if rank == 0:
        print("OK", world, int(whole.sum()))
""")


def test_world_size_2_allreduce_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "OK 2" in out.stdout


def test_pack_reads_host_matches_numpy():
    """csrc/ingest.cpp (no device needed): the AVX-512 and the table-driven packers against a numpy restatement of
    flat_kmers.py:134-145 at 2 bits per base; dirty rows are listed, clean rows keep their order"""
    from graph_kmer_index_b200.read_kmers import pack_reads
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for L in (150, 64, 31, 200, 1):
        n = 20000
        reads = alphabet[rng.integers(0, 8, (n, L))].copy()
        bad = rng.choice(n, 300, replace=False)
        reads[bad, rng.integers(0, L, 300)] = np.frombuffer(b"NnX-", dtype=np.uint8)[rng.integers(0, 4, 300)]
        words = (L + 31) // 32
        lut = np.zeros(256, dtype=np.uint64)
        for i, c in enumerate(b"acgt"):
            lut[c] = lut[c - 32] = i
        valid = np.isin(reads, alphabet).all(axis=1)
        codes = lut[reads[valid]]
        want = np.zeros((int(valid.sum()), words), dtype=np.uint64)
        for i in range(L):
            want[:, i // 32] |= codes[:, i] << np.uint64(2 * (i % 32))
        for force_scalar in (False, True):
            for threads in (1, 3):
                packed, dirty = pack_reads(reads, n_threads=threads, force_scalar=force_scalar)
                assert np.array_equal(packed, want), (L, force_scalar, threads)
                assert np.array_equal(dirty, np.flatnonzero(~valid))
        padded = np.full((n, L + 7), ord("N"), dtype=np.uint8)
        padded[:, :L] = reads
        packed, dirty = pack_reads(padded[:, :L])
        assert np.array_equal(packed, want) and np.array_equal(dirty, np.flatnonzero(~valid))
    packed, dirty = pack_reads(np.zeros((0, 150), dtype=np.uint8))
    assert packed.shape == (0, 5) and len(dirty) == 0


def _write_fastx(path, reads, fmt, rng):
    """reads: list of str.  FASTA: header + one line per read, with a few multi-line records, blank-padded and CRLF lines;
    FASTQ: four lines per record.  Returns the sequence lines the reference's line filter would see (stripped)."""
    lines_seen = []
    with open(path, "w", newline="") as f:
        for i, r in enumerate(reads):
            if fmt == "fastq":
                f.write("@read%d some text\n%s\n+\n%s\n" % (i, r, "@" * len(r)))       # '@' in the quality line on purpose
                lines_seen.append(r)
            else:
                f.write(">read%d\n" % i)
                if i % 7 == 3 and len(r) > 20:                                            # multi-line record: two reads for the reference
                    f.write(r[:13] + "\n" + r[13:] + "\n")
                    lines_seen += [r[:13], r[13:]]
                elif i % 7 == 5:
                    f.write("  " + r + " \r\n")                                           # blanks + CRLF are stripped
                    lines_seen.append(r)
                else:
                    f.write(r + "\n")
                    lines_seen.append(r)
        if fmt == "fasta":
            f.write(">last\nACGTACGTAC")                                                  # no newline at the end of the file
            lines_seen.append("ACGTACGTAC")
    return lines_seen


def test_fastx_line_index(tmp_path, monkeypatch):
    """csrc/ingest.cpp fastx_open (no device needed): the sequence lines are the ones read_kmers.py:16-21 would hash; files of
    several MB so that the byte ranges of the parsing threads cut lines at arbitrary places"""
    from graph_kmer_index_b200.read_kmers import FastxFile
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    reads = [alphabet[rng.integers(0, 9, int(n))].tobytes().decode() for n in rng.choice([150, 150, 150, 151, 75, 31, 5, 1], size=40000)]
    for fmt in ("fasta", "fastq"):
        path = tmp_path / ("reads." + fmt)
        want = _write_fastx(path, reads, fmt, rng)
        data = open(path, "rb").read()
        ref_lines = [l.strip() for l in open(path).readlines() if not l.startswith(">")] if fmt == "fasta" else want
        assert ref_lines == want
        for threads in ("0", "2", "6", "13"):
            monkeypatch.setenv("GKI_PACK_THREADS", threads)
            with FastxFile(path) as f:
                assert f.format == fmt and f.n_reads == len(want)
                offsets, lengths = f.lines()
            assert [data[o:o + n].decode() for o, n in zip(offsets, lengths)] == want, (fmt, threads)
    monkeypatch.delenv("GKI_PACK_THREADS")
    empty = tmp_path / "empty.fa"
    empty.write_text("")
    with FastxFile(empty) as f:
        assert f.n_reads == 0
    import pytest
    from graph_kmer_index_b200 import _lib
    with pytest.raises(_lib.GkiError):
        FastxFile(tmp_path / "missing.fa")


def test_critical_path_chunks_like_the_reference_cli():
    """command_line_interface.py:588-603: n_paths // n_chunks paths per chunk, the remainder in further chunks"""
    def reference(n_paths, n_chunks):
        if n_chunks >= n_paths:
            n_chunks = n_paths
        per = n_paths // n_chunks
        starts = list(range(0, n_paths, per))
        return list(zip(starts, starts[1:] + [n_paths]))
    for n_paths in (1, 2, 7, 9, 40, 1001):
        for n_chunks in (1, 2, 3, 5, 20, 160, 5000):
            chunks = distributed.critical_path_chunks(n_paths, n_chunks)
            assert chunks == reference(n_paths, n_chunks)
            assert chunks[0][0] == 0 and chunks[-1][1] == n_paths and all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (no GPU needed): one JSON line with the contract's keys, the oracle port on the host cores"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "kmers/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    native = d["reference_native"]        # the reference's compiled probe (oracle/_ref), when it was built
    assert native is None or (native["kind"] == "reference" and native["cores"] == 1 and 0 < native["value"] < d["value"] * 50)


def test_follow_only_rewrites_the_edge_lists():
    """kmer_finder.follow_only (kf:385-388 as a graph rewrite): a node with successors in the set keeps exactly the set's
    intersection with them, in the set's iteration order, and is flagged; every other node is untouched."""
    from graph_kmer_index_b200.kmer_finder import follow_only
    rng = np.random.default_rng(3)
    n = 200
    degrees = rng.integers(0, 4, n)
    edge_offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(degrees, out=edge_offsets[1:])
    edges = rng.integers(0, n, int(edge_offsets[-1])).astype(np.int32)      # duplicates edges included
    for follow in (set(), {7, 8}, set(int(x) for x in rng.choice(n, 40, replace=False))):
        new_offsets, new_edges, force = follow_only(edge_offsets, edges, follow)
        assert new_offsets.dtype == np.int64 and new_edges.dtype == np.int32 and force.dtype == np.uint8 and len(force) == n
        for u in range(n):
            succ = [int(x) for x in edges[edge_offsets[u]:edge_offsets[u + 1]]]
            got = [int(x) for x in new_edges[new_offsets[u]:new_offsets[u + 1]]]
            want = list(follow.intersection(succ))
            assert got == (want if want else succ) and bool(force[u]) == bool(want), u
    # {7, 8}.intersection([7, 8]) iterates 8 first: the order of the reference's loop, not the edge order
    o, e, f = follow_only(np.array([0, 2, 2, 2, 2, 2, 2, 2, 2, 2]), np.array([7, 8], dtype=np.int32), {7, 8})
    assert list(e) == list({7, 8}.intersection([7, 8])) and list(f) == [1] + [0] * 8


def test_finder_starting_points_in_order():
    """kmer_finder.starting_points_in_order against the reference's list handling restated literally (kf:192-226: reverse, slice,
    append, pop from the end, stop at the chunk boundary)."""
    from graph_kmer_index_b200.kmer_finder import starting_points_in_order
    rng = np.random.default_rng(4)
    for trial in range(300):
        n = int(rng.integers(0, 12))
        crit_nodes = np.sort(rng.choice(40, n, replace=False)).astype(np.uint32)
        if n and trial % 3 == 0:
            crit_nodes[rng.integers(0, n)] = crit_nodes[0]                 # several critical positions on one node
        crit_offsets = rng.integers(1, 50, n).astype(np.uint16)
        start_at = [None, 0, int(rng.integers(0, 14))][trial % 3]
        stop_at = [None, int(rng.integers(0, 14))][trial % 2]
        first = 0 if (start_at is None or start_at == 0) and trial % 5 else None
        starting_points = [(int(a), int(b)) for a, b in zip(crit_nodes, crit_offsets)][::-1]
        stop_at_node = None
        if stop_at is not None and stop_at < len(starting_points):
            stop_at_node = starting_points[-stop_at - 1][0]
        if start_at is not None and start_at > 0:
            starting_points = starting_points[:-start_at]
        if first is not None:
            starting_points.append((first, 0))
        want = []
        while len(starting_points) > 0:
            node, offset = starting_points.pop()
            if stop_at_node is not None and stop_at_node == node:
                break
            want.append((node, offset))
        nodes, offsets = starting_points_in_order(crit_nodes, crit_offsets, first, start_at, stop_at)
        assert list(zip(nodes.tolist(), offsets.tolist())) == want, (trial, crit_nodes, start_at, stop_at, first)


def test_fastx_line_index_edge_files(tmp_path):
    """fastx_open on small odd files, with the AVX-512 newline scan (in this process, if the CPU has it) and with the memchr
    path (GKI_PACK_SCALAR=1 in a child process): empty lines, CRLF, no final newline, only newlines, one line, a file of
    several MB without any newline, lines around the 64-byte steps of the scan."""
    from graph_kmer_index_b200.read_kmers import FastxFile
    cases = {
        "one.fa": b"ACGT", "one_nl.fa": b"ACGT\n", "only_newlines.fa": b"\n\n\n", "header_only.fa": b">x\n", "crlf.fa": b">a\r\nAC GT \r\n\r\n>b\r\nTT",
        "blank_lines.fa": b">a\n\nACGT\n\n\n>b\nGG\n\n", "q1.fq": b"@r\nACGT\n+\n!!!!", "q2.fq": b"@r\nACGT\n+\n!!!!\n@s\n\n+\n\n@t\nGGA\n+\n@@@\n",
        "steps.fa": b"".join(b">" + b"h" * n + b"\n" + b"A" * (127 - n) + b"\n" for n in range(0, 127)),
        "long_no_newline.fa": b"ACGT" * (1 << 20),
        "long_lines.fa": b">h\n" + b"\n".join(b"C" * n for n in [1 << 20, 63, 64, 65, 1 << 21, 0, 1]),
    }
    expected = {}
    for name, content in cases.items():
        (tmp_path / name).write_bytes(content)
        lines = content.decode().split("\n")
        if content.endswith(b"\n"):
            lines = lines[:-1]                       # nothing follows the last newline
        if name.endswith(".fq"):
            expected[name] = [l.strip() for l in lines[1::4]]
        else:
            expected[name] = [l.strip() for l in lines if not l.startswith(">")]
        # the reference's own filter (read_kmers.py:16-18) on the same file
        if not name.endswith(".fq"):
            assert expected[name] == [l.strip() for l in open(tmp_path / name, newline="\n").readlines() if not l.startswith(">")], name

    def check():
        for name, content in cases.items():
            with FastxFile(tmp_path / name) as f:
                offsets, lengths = f.lines()
                got = [content[o:o + n].decode() for o, n in zip(offsets, lengths)]
            assert got == expected[name], (name, got[:5], expected[name][:5])
    check()
    script = tmp_path / "scalar.py"
    script.write_text(textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from graph_kmer_index_b200.read_kmers import FastxFile
        import ast
        cases = ast.literal_eval(open(%r).read())
        for path, want in cases.items():
            content = open(path, "rb").read()
            with FastxFile(path) as f:
                offsets, lengths = f.lines()
            got = [content[o:o + n].decode() for o, n in zip(offsets, lengths)]
            assert got == want, (path, got[:5], want[:5])
        print("ok")
    """ % (ROOT, str(tmp_path / "expected.txt"))))
    (tmp_path / "expected.txt").write_text(repr({str(tmp_path / name): want for name, want in expected.items()}))
    out = subprocess.run([sys.executable, str(script)], env=dict(os.environ, GKI_PACK_SCALAR="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]


def test_kmer_index2_host_logic_with_the_device_calls_emulated(monkeypatch):
    """KmerIndex2 / MultiValueHashTable host logic (row numbers as the node column, frequencies from the set_frequencies pass, input
    order per key) with the two device calls underneath replaced by the numpy oracle -- the GPU tests run the real ones."""
    import graph_kmer_index_b200 as gki
    import graph_kmer_index_b200.collision_free_kmer_index as cfki
    from graph_kmer_index_b200.multi_value_hashtable import MultiValueHashTable

    def build(kmers, nodes, ref, af, modulo, skip):
        kmers, n = np.asarray(kmers), len(kmers)
        idx = no.build_index(kmers.astype(np.uint64), nodes if nodes is not None else np.zeros(n, np.uint32),
                             ref if ref is not None else np.zeros(n, np.uint64), af if af is not None else np.zeros(n, np.float32), modulo, skip)
        return (idx["_hashes_to_index"], idx["_n_kmers"], idx["_kmers"].view(kmers.dtype), idx["_nodes"], idx["_ref_offsets"],
                idx["_allele_frequencies"], idx["_frequencies"])

    class Device:
        def __init__(self, index):
            self.index = index

        def lookup_entries(self, kmers):
            entries, qidx = [], []
            for j, kmer in enumerate(cfki._queries(kmers)):
                bucket = int(kmer) % self.index._modulo
                lo = int(self.index._hashes_to_index[bucket])
                for p in range(lo, lo + int(self.index._n_kmers[bucket])):
                    if int(self.index._kmers.view(np.uint64)[p]) == int(kmer):
                        entries.append(p)
                        qidx.append(j)
            return np.array(entries, dtype=np.int64), np.array(qidx, dtype=np.int64)

    monkeypatch.setattr(cfki, "build_index_arrays", build)
    monkeypatch.setattr(cfki.CollisionFreeKmerIndex, "device_index", lambda self: Device(self))
    h = MultiValueHashTable.from_keys_and_values([1, 2, 3, 1], {"nodes": np.array([1, 2, 3, 10]), "offsets": np.array([5, 3, 2, 100])}, mod=11)
    assert np.all(h[1]["nodes"] == [1, 10]) and np.all(h[2]["offsets"] == [3])                 # the reference's tests/test_multi_value_hashtable.py
    flat = gki.FlatKmers2(np.array([1, 1, 1, 2, 3, 10, 11, 2]), np.array([1, 1, 2, 2, 3, 1, 10, 5]), np.array([0, 0, 1, 2, 3, 4, 5, 6]),
                          np.array([1, 2, 3, 4, 5, 6, 7, 8]), np.array([0.4, 0.1, 0.3, 0.4, 0.1, 0.1, 0.1, 0.1]))
    index = gki.KmerIndex2.from_flat_kmers(flat)                                                # the reference's tests/test_indexes2.py
    assert index.get_kmer_frequency(1) == 2 and np.all(index.get_start_nodes(1) == [1, 1, 2]) and np.all(index.get_nodes(3) == [5])
    rng = np.random.default_rng(8)
    n = 3000
    flat = gki.FlatKmers2(rng.integers(0, 200, n), rng.integers(0, 30, n).astype(np.int32), rng.integers(0, 4, n).astype(np.int16),
                          rng.integers(0, 1000, n).astype(np.int32), rng.random(n))
    index = gki.KmerIndex2.from_flat_kmers(flat, modulo=97)
    for kmer in np.unique(flat._hashes)[::7]:
        rows = flat._hashes == kmer
        assert np.array_equal(index.get_nodes(kmer), flat._nodes[rows]) and np.array_equal(index._data[kmer]["allele_frequencies"], flat._allele_frequencies[rows])
        assert index.get_kmer_frequency(kmer) == len(set(zip(flat._start_nodes[rows].tolist(), flat._start_offsets[rows].tolist())))


def test_pack_reads_every_byte_value():
    """The AVX-512 packer (nibble-table shuffle) and the table-driven scalar packer classify every byte value alike: a row is clean
    iff all its bytes are in ACGTacgt, at every position of a 64-byte step and in the masked tail."""
    from graph_kmer_index_b200.read_kmers import pack_reads
    L = 150
    base = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[np.random.default_rng(2).integers(0, 8, L)]
    rows = []
    for value in range(256):
        for position in (0, 1, 63, 64, 127, 128, 149):
            row = base.copy()
            row[position] = value
            rows.append(row)
    reads = np.array(rows, dtype=np.uint8)
    wide_packed, wide_dirty = pack_reads(reads, n_threads=1)
    scalar_packed, scalar_dirty = pack_reads(reads, n_threads=1, force_scalar=True)
    assert np.array_equal(wide_dirty, scalar_dirty) and np.array_equal(wide_packed, scalar_packed)
    want_dirty = np.flatnonzero(~np.isin(reads, np.frombuffer(b"ACGTacgt", dtype=np.uint8)).all(axis=1))
    assert np.array_equal(np.sort(wide_dirty), want_dirty) and len(want_dirty) == (256 - 8) * 7


def test_reference_module_paths_and_nplist():
    """graph_kmer_index/__init__.py:1-12 and the modules KAGE imports from: snp_kmer_finder, critical_graph_paths, nplist"""
    import importlib
    for module, names in (("snp_kmer_finder", ("kmer_to_hash_fast", "sequence_to_kmer_hash", "kmer_hash_to_sequence")),
                          ("critical_graph_paths", ("CriticalGraphPaths",)), ("nplist", ("NpList",)),
                          ("flat_kmers", ("FlatKmers", "letter_sequence_to_numeric", "numeric_to_letter_sequence"))):
        mod = importlib.import_module("graph_kmer_index_b200." + module)
        for name in names:
            assert hasattr(mod, name), (module, name)
    from graph_kmer_index_b200.nplist import NpList
    a = NpList(dtype=np.int64)
    for i in range(250):
        a.append(i)
    a.extend(np.array([7, 8, 9]))
    assert len(a) == 253 and a[-1] == 9 and a[100] == 100 and a.get_nparray().dtype == np.int64
    b = a.copy()
    assert b == a and len(b) == 253
    a.set_n_elements(10)
    assert len(a) == 10 and a.get_nparray().tolist() == list(range(10))
    c = NpList()
    c.append(5)
    c.append(7)
    assert c.get_nparray().tolist() == [5, 7]
