"""CounterKmerIndex.count_kmers on a host numpy array of read k-mer hashes (the reference's own call, cfki:33-37): plain
cudaMemcpyAsync chunks vs the threaded pinned-buffer copy (GKI_HOST_COPY_THREADS).  python profiles/count_kmers_host.py [queries]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_kmer_index_b200 as gki  # noqa: E402
from graph_kmer_index_b200 import synthetic  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000_000
    n = 10_000_000
    hashes, nodes, ref, af = synthetic.flat_kmers(n, n // 10, 31)
    index = gki.CollisionFreeKmerIndex.from_flat_kmers(gki.FlatKmers(hashes, nodes, ref, af), skip_frequencies=True)
    counter = gki.CounterKmerIndex.from_kmer_index(index)
    rng = np.random.default_rng(0)
    queries = rng.integers(0, 1 << 62, nq, dtype=np.int64).view(np.uint64)
    queries[::10] = hashes[rng.integers(0, n, len(queries[::10]))]
    want = None
    for threads in ("0", "8", "14"):
        os.environ["GKI_HOST_COPY_THREADS"] = threads
        secs = []
        for _ in range(3):
            counter.reset()
            t0 = time.perf_counter()
            counter.count_kmers(queries)
            secs.append(time.perf_counter() - t0)
        got = counter.get_node_counts()
        want = got if want is None else want
        assert np.array_equal(got, want) and got.sum() >= nq // 10
        print(json.dumps({"queries": nq, "host_copy_threads": int(threads), "seconds": secs, "kmers_per_s": nq / min(secs),
                          "gb_per_s": 8 * nq / min(secs) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
