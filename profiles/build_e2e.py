"""CollisionFreeKmerIndex.from_flat_kmers end to end on host numpy arrays (the reference-facing call, cfki:422-467): plain
cudaMemcpyAsync on pageable memory vs the threaded pinned-buffer copy of runtime.cu (GKI_HOST_COPY_THREADS).
python profiles/build_e2e.py [entries] -> JSON lines"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graph_kmer_index_b200 as gki  # noqa: E402
from graph_kmer_index_b200 import synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
    hashes, nodes, ref, af = synthetic.flat_kmers(n, n // 10, 31)
    flat = gki.FlatKmers(hashes, nodes, ref, af)
    modulo = 452930477
    first = None
    for threads, huge in (("0", "1"), ("8", "0"), ("8", "1"), ("14", "0"), ("14", "1"), ("8", "0"), ("8", "1")):
        os.environ["GKI_HOST_COPY_THREADS"] = threads
        os.environ["GKI_HOST_COPY_HUGEPAGES"] = huge      # MADV_HUGEPAGE on the fresh destination arrays
        secs = []
        for _ in range(3):
            t0 = time.perf_counter()
            index = gki.CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=True)
            secs.append(time.perf_counter() - t0)
        if first is None:
            first = index
        else:
            assert np.array_equal(index._hashes_to_index, first._hashes_to_index) and np.array_equal(index._nodes, first._nodes)
        print(json.dumps({"entries": n, "modulo": modulo, "host_copy_threads": int(threads), "madvise_hugepage": int(huge), "seconds": secs, "entries_per_s": n / min(secs),
                          "host_bytes": 24 * n + 26 * n + 8 * modulo, "cpus": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
