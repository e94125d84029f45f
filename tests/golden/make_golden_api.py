"""Generate tests/golden/api_misc.npz: outputs of the UNMODIFIED reference for the smaller CollisionFreeKmerIndex / FlatKmers /
ReadKmers / kmer_hashing methods on the hot-path classes (get_grouped_nodes, get_frequency, set_frequencies_using_other_index,
kmer_hashes_to_complement_bases, FlatKmers.sum_of_kmer_frequencies / maximum_kmer_frequency, ReadKmers.from_list_of_string_kmers).
Build container only: python tests/golden/make_golden_api.py"""
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
from graph_kmer_index import CollisionFreeKmerIndex, FlatKmers, ReadKmers  # noqa: E402
from graph_kmer_index.kmer_hashing import kmer_hashes_to_complement_bases  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "index_small.npz"))
    k = int(g["k"])
    flat = FlatKmers(g["in_hashes"], g["in_nodes"], g["in_ref_offsets"], g["in_allele_frequencies"])
    modulo = int(g["stable_modulo"])
    with ref_shims.stable_argsort():
        index = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo)
    rng = np.random.default_rng(3)
    present = rng.choice(np.unique(index._kmers), size=60, replace=False)
    absent = rng.integers(0, 4 ** k, 20, dtype=np.uint64)
    queries = np.concatenate([present, absent]).astype(np.uint64)
    out = {"queries": queries, "k": np.int64(k)}
    freq_rc, freq_fwd, grouped_flat, grouped_sizes, grouped_none = [], [], [], [], []
    for q in queries:
        freq_rc.append(index.get_frequency(int(q), include_reverse_complement=True, k=k))
        freq_fwd.append(index.get_frequency(int(q), include_reverse_complement=False, k=k))
        groups = index.get_grouped_nodes(int(q), max_hits=10 ** 9)
        grouped_none.append(groups is None)
        if groups is not None:
            grouped_sizes.append(np.array([len(x) for x in groups], dtype=np.int64))
            grouped_flat.append(np.concatenate([np.sort(np.asarray(x)) for x in groups]))   # order inside a group: argsort ties
        else:
            grouped_sizes.append(np.zeros(0, dtype=np.int64))
            grouped_flat.append(np.zeros(0, dtype=index._nodes.dtype))
    out["freq_rc"], out["freq_fwd"] = np.array(freq_rc, dtype=np.int64), np.array(freq_fwd, dtype=np.int64)
    out["grouped_none"] = np.array(grouped_none)
    out["grouped_n_groups"] = np.array([len(x) for x in grouped_sizes], dtype=np.int64)
    out["grouped_sizes"] = np.concatenate(grouped_sizes)
    out["grouped_nodes"] = np.concatenate(grouped_flat)
    # gate of `get` on max_hits (cfki:312)
    out["get_max_hits_1_is_none"] = np.array([index.get(int(q), max_hits=1)[0] is None for q in queries])
    # set_frequencies_using_other_index (cfki:252-265): frequencies of a copy set from the index itself, multiplier 3, min 2
    other = index.copy()
    target = index.copy()
    target._frequencies = target._frequencies.copy()
    target.set_frequencies_using_other_index(other, multiplier=3, min_frequency=2)
    out["set_from_other"] = target._frequencies
    # FlatKmers.sum_of_kmer_frequencies / maximum_kmer_frequency (flat_kmers.py:92-96) on a small slice
    small = FlatKmers(flat._hashes[:200], flat._nodes[:200], flat._ref_offsets[:200], flat._allele_frequencies[:200])
    out["sum_of_kmer_frequencies"] = np.int64(small.sum_of_kmer_frequencies(index))
    out["maximum_kmer_frequency"] = np.int64(small.maximum_kmer_frequency(index))
    # kmer_hashing.kmer_hashes_to_complement_bases (kmer_hashing.py:40-49)
    out["complement_bases"] = kmer_hashes_to_complement_bases(queries[:25], k)
    # ReadKmers.from_list_of_string_kmers (read_kmers.py:51-57)
    strings = [["ACGTA", "ttgca", "NACGT"], ["GGGGG"], []]
    rk = ReadKmers.from_list_of_string_kmers(strings)
    out["string_kmers"] = np.array([h for read in rk.kmers for h in read], dtype=np.uint64)
    out["string_kmers_per_read"] = np.array([len(read) for read in rk.kmers], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "api_misc.npz"), **out)
    print({key: (v.shape, v.dtype) for key, v in out.items()})


if __name__ == "__main__":
    main()
