"""Generate tests/golden/side_indexes.npz by running the UNMODIFIED reference ReverseKmerIndex.from_flat_kmers
(reverse_kmer_index.py:59-84) and ReferenceKmerIndex.from_flat_kmers / from_sequence (reference_kmer_index.py:50-121)
on small seeded inputs, np.argsort made stable (the canonical order).  Build container only:

    python tests/golden/make_golden_side_indexes.py
"""
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

from graph_kmer_index import FlatKmers, ReferenceKmerIndex, ReverseKmerIndex  # noqa: E402


def cases():
    rng = np.random.default_rng(23)
    out = {}
    # name -> (hashes, nodes, ref_offsets)
    n = 5000
    out["dense"] = (rng.integers(0, 2 ** 62, n, dtype=np.uint64), rng.integers(0, 700, n).astype(np.uint32),
                    rng.integers(0, 3000, n).astype(np.uint64))
    # gaps longer than 32 between reference positions, node ids with holes, small hashes (-> uint32 k-mers, :87-89)
    pos = np.sort(rng.choice(200000, size=900, replace=False)).astype(np.uint64)
    out["gaps"] = (rng.integers(0, 2 ** 31, 2700, dtype=np.uint64), (rng.integers(0, 50, 2700) * 37 + 5).astype(np.uint32),
                   rng.permutation(np.repeat(pos, 3)))
    # first reference position repeated (the unmarked first run of :92), position 0 present
    out["first_run"] = (np.arange(1, 41, dtype=np.uint64) * 1000003, np.arange(40, dtype=np.uint32) % 7,
                        np.array([0] * 5 + [3] * 4 + [4] * 11 + [90] * 20, dtype=np.uint64)[rng.permutation(40)])
    # one distinct position only; one entry only
    out["single_position"] = (np.arange(10, dtype=np.uint64) + 2 ** 40, np.arange(10, dtype=np.uint32), np.full(10, 17, dtype=np.uint64))
    out["one_entry"] = (np.array([123456789], dtype=np.uint64), np.array([4], dtype=np.uint32), np.array([9], dtype=np.uint64))
    # a node with more than 65535 entries (uint16 n_kmers wraps, :70,82)
    m = 70000
    out["wrap"] = (np.arange(m + 10, dtype=np.uint64) * 7 + 2 ** 33, np.concatenate([np.full(m, 3), np.arange(10)]).astype(np.uint32),
                   (np.arange(m + 10, dtype=np.uint64) * 13) % 100)
    # the reference tests' own inputs (tests/test_reverse_kmer_index.py:7, tests/test_reference_kmer_index.py:7-10)
    out["ref_test_reverse"] = (np.array([10, 3, 11, 4], dtype=np.uint64), np.array([5, 3, 5, 8], dtype=np.uint32), np.zeros(4, dtype=np.uint64))
    out["ref_test_reference"] = (np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.uint64), np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.uint32),
                                 np.array([4, 4, 5, 5, 1, 2, 3], dtype=np.uint64))
    return out


def main():
    out = {}
    names = []
    for name, (hashes, nodes, ref) in cases().items():
        names.append(name)
        flat = FlatKmers(hashes, nodes, ref)
        with ref_shims.stable_argsort():
            rev = ReverseKmerIndex.from_flat_kmers(flat)
            refi = ReferenceKmerIndex.from_flat_kmers(flat)
        out[name + "/hashes"], out[name + "/nodes"], out[name + "/ref_offsets"] = hashes, nodes, ref
        out[name + "/rev_nodes_to_index_positions"] = rev.nodes_to_index_positions
        out[name + "/rev_nodes_to_n_hashes"] = rev.nodes_to_n_hashes
        out[name + "/rev_hashes"] = rev.hashes
        out[name + "/rev_ref_positions"] = rev.ref_positions
        out[name + "/ref_ref_position_to_index"] = refi.ref_position_to_index
        out[name + "/ref_kmers"] = refi.kmers
        out[name + "/ref_ref_positions"] = refi.ref_positions
        out[name + "/ref_nodes"] = refi.nodes
    out["names"] = np.array(names)
    # from_sequence (reference_kmer_index.py:50-67) on a letter sequence, k <= 16 (uint32) and k > 16 (uint64)
    rng = np.random.default_rng(5)
    seq = "".join(rng.choice(list("ACGTacgtNn"), 500, p=[.12] * 8 + [.02] * 2))
    out["seq"] = np.frombuffer(seq.encode(), dtype=np.uint8)
    for k in (5, 16, 31):
        idx = ReferenceKmerIndex.from_sequence(seq, k)
        out["seq_k%d_kmers" % k] = idx.kmers
        out["seq_k%d_index" % k] = idx.ref_position_to_index
    np.savez_compressed(os.path.join(HERE, "side_indexes.npz"), **out)
    print("wrote side_indexes.npz:", {k: v.shape for k, v in out.items() if k.startswith("first_run")})


if __name__ == "__main__":
    main()
