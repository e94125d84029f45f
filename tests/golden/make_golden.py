"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shims.py) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are committed; the GPU box never needs /root/reference.
"""
import glob
import importlib.util
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_shims  # noqa: E402

ref_shims.install()
logging.disable(logging.CRITICAL)

from graph_kmer_index import (CollisionFreeKmerIndex, CounterKmerIndex, FlatKmers, ReadKmers,  # noqa: E402
                              kmer_hash_to_sequence, letter_sequence_to_numeric, sequence_to_kmer_hash)
from graph_kmer_index.kmer_hashing import (kmer_hashes_to_bases, kmer_hashes_to_complement_hashes,  # noqa: E402
                                           kmer_hashes_to_reverse_complement_hash, power_array)
from graph_kmer_index_b200 import synthetic  # noqa: E402  (numpy generators only; no CUDA touched)


def hashing_fixture():
    rng = np.random.default_rng(7)
    L, n = 64, 48
    alphabet = np.frombuffer(b"ACGTacgtNnMmRYKXU-", dtype=np.uint8)
    probs = np.array([10, 10, 10, 10, 3, 3, 3, 3, 1, 1, 1, 1, .3, .3, .3, .3, .3, .3])
    reads = alphabet[rng.choice(len(alphabet), size=(n, L), p=probs / probs.sum())]
    reads[0] = ord("A")
    reads[1] = ord("T")
    reads[2] = ord("N")
    out = {"reads": reads}
    for k in (1, 3, 5, 16, 31):
        pv = power_array(k)
        fwd, rc = [], []
        for row in reads:
            s = row.tobytes().decode("latin-1")
            fwd.append(ReadKmers.get_kmers_from_read_dynamic(s, pv))
            rc.append(ReadKmers.get_kmers_from_read_dynamic(str(ref_shims.Seq(s).reverse_complement()), pv))
        out["fwd_k%d" % k] = np.array(fwd, dtype=np.uint64)
        out["rc_k%d" % k] = np.array(rc, dtype=np.uint64)
    # numeric encoding of one mixed string (flat_kmers.py:134-145)
    s = "ACGTacgtNnMmxyz"
    out["encode_in"] = np.frombuffer(s.encode(), dtype=np.uint8)
    out["encode_out"] = letter_sequence_to_numeric(s)
    np.savez_compressed(os.path.join(HERE, "hashing.npz"), **out)


def revcomp_fixture():
    rng = np.random.default_rng(11)
    out = {}
    for k in (1, 2, 4, 9, 16, 30, 31):
        h = rng.integers(0, 4 ** k, size=200, dtype=np.uint64)
        h[0] = 0
        h[1] = 4 ** k - 1
        out["in_k%d" % k] = h
        out["rc_k%d" % k] = kmer_hashes_to_reverse_complement_hash(h.copy(), k)
        out["comp_k%d" % k] = kmer_hashes_to_complement_hashes(h.copy(), k)
        out["bases_k%d" % k] = kmer_hashes_to_bases(h.copy(), k)
    seqs = ["atg", "Acacatacgactacg", "CAtgAACAtttggtAATCTACAtgAACAttt", "G", "T" * 31]
    out["seq_hashes"] = np.array([sequence_to_kmer_hash(s) for s in seqs], dtype=np.uint64)
    out["seq_lens"] = np.array([len(s) for s in seqs])
    out["seq_back"] = np.array([kmer_hash_to_sequence(sequence_to_kmer_hash(s), len(s)) for s in seqs])
    out["seqs"] = np.array(seqs)
    np.savez_compressed(os.path.join(HERE, "revcomp.npz"), **out)


def _index_dict(idx):
    return dict(hashes_to_index=idx._hashes_to_index, n_kmers=idx._n_kmers, nodes=idx._nodes,
                ref_offsets=idx._ref_offsets, kmers=idx._kmers, modulo=np.int64(idx._modulo),
                frequencies=idx._frequencies, allele_frequencies=idx._allele_frequencies)


def _load_ref_cython():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cython_kmer_index*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("cython_kmer_index", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def index_fixture(name, n_entries, n_nodes, modulo, k, with_freq, n_reads, read_len):
    hashes, nodes, ref, af = synthetic.flat_kmers(n_entries, n_nodes, k)
    if with_freq:  # make some k-mers repeat at several ref offsets so that set_frequencies is non-trivial
        ref = ref.copy()
        ref[::7] += np.uint64(1000003)
        hashes = hashes.copy()
        hashes[::11] = hashes[5]
    flat = FlatKmers(hashes.copy(), nodes.copy(), ref.copy(), af.copy())
    with ref_shims.stable_argsort():
        stable = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=not with_freq)
    asrun = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=not with_freq)
    out = {"in_hashes": hashes, "in_nodes": nodes, "in_ref_offsets": ref, "in_allele_frequencies": af,
           "k": np.int64(k), "n_nodes": np.int64(n_nodes)}
    out.update({"stable_" + key: v for key, v in _index_dict(stable).items()})
    out.update({"asrun_" + key: v for key, v in _index_dict(asrun).items()})

    # queries: index k-mers, near misses in occupied buckets, random misses
    rng = np.random.default_rng(3)
    q = np.concatenate([hashes[rng.integers(0, n_entries, 400)],
                        hashes[rng.integers(0, n_entries, 200)] + np.uint64(modulo),
                        rng.integers(0, 4 ** k, 400, dtype=np.uint64),
                        np.array([0, modulo, 2 * modulo], dtype=np.uint64)])
    rng.shuffle(q)
    out["queries"] = q
    # single-k-mer get (cfki:303-315) on the stable index, max_hits large
    got_nodes, got_off, got_n = [], [], []
    for kmer in q[:200]:
        n_, o_, f_, a_ = stable.get(kmer, max_hits=10 ** 9)
        got_n.append(0 if n_ is None else len(n_))
        if n_ is not None:
            got_nodes.extend(n_)
            got_off.extend(o_)
    out["get_n"] = np.array(got_n)
    out["get_nodes"] = np.array(got_nodes, dtype=np.uint32)
    out["get_ref_offsets"] = np.array(got_off, dtype=np.uint64)
    # CounterKmerIndex (cfki:14-40) with the dict stand-in for npstructures.Counter
    counter = CounterKmerIndex.from_kmer_index(stable)
    counter.count_kmers(q)
    counter.count_kmers(q[:100])
    out["node_counts_min0"] = counter.get_node_counts()
    out["node_counts_min_big"] = counter.get_node_counts(n_nodes + 17)
    # the reference's compiled Cython probe (pyx:47-109) -- as run (gates on)
    cy = _load_ref_cython()
    if cy is not None:
        class _I:  # pyx:33-41 wants long[:] hashes_to_index
            pass
        view = _I()
        for key in ("_n_kmers", "_nodes", "_ref_offsets", "_kmers", "_frequencies", "_allele_frequencies", "_modulo"):
            setattr(view, key, getattr(stable, key))
        view._hashes_to_index = stable._hashes_to_index.astype(np.int64)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            out["cython_get"] = cy.CythonKmerIndex(view).get(q)
    # reads -> node counts through the reference pieces: ReadKmers hashing + CounterKmerIndex
    reads = synthetic.reads(n_reads, read_len, n_entries, k, p_hit_permille=500, n_permille=20)
    pv = power_array(k)
    counter = CounterKmerIndex.from_kmer_index(stable)
    for row in reads:
        s = row.tobytes().decode()
        counter.count_kmers(ReadKmers.get_kmers_from_read_dynamic(s, pv))
        counter.count_kmers(ReadKmers.get_kmers_from_read_dynamic(str(ref_shims.Seq(s).reverse_complement()), pv))
    out["reads"] = reads
    out["read_node_counts"] = counter.get_node_counts(n_nodes)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def tiny_fixture():
    """tests/test_collision_free_kmer_index.py:8-14 of the reference."""
    flat = FlatKmers(np.array([1, 1, 2, 2, 4, 5, 3], dtype=np.uint64), np.array([5, 6, 7, 8, 10, 11, 100]),
                     np.array([1, 1, 2, 3, 10, 11, 100]))
    with ref_shims.stable_argsort():
        idx = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=4)
    out = {"stable_" + key: v for key, v in _index_dict(idx).items()}
    idx2 = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=4)
    out.update({"asrun_" + key: v for key, v in _index_dict(idx2).items()})
    np.savez_compressed(os.path.join(HERE, "tiny_index.npz"), **out)


if __name__ == "__main__":
    hashing_fixture()
    revcomp_fixture()
    tiny_fixture()
    index_fixture("index_small", n_entries=6000, n_nodes=500, modulo=2003, k=31, with_freq=True,
                  n_reads=300, read_len=80)
    index_fixture("index_sparse", n_entries=3001, n_nodes=2000, modulo=20011, k=15, with_freq=False,
                  n_reads=200, read_len=50)
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        print(os.path.basename(f), os.path.getsize(f))
