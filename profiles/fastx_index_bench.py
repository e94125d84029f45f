"""Line indexing of a FASTQ file (csrc/ingest.cpp fastx_open, host only): writes a synthetic file of `n` 150 bp records to /tmp and
times FastxFile() for several thread counts; GKI_FASTX_DEBUG=1 prints the phases.  python profiles/fastx_index_bench.py [n_reads]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_fastq(path, n, read_len=150):
    rng = np.random.default_rng(0)
    rec = np.empty((n, 1 + 10 + 1 + read_len + 3 + read_len + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    idx = np.arange(n)
    for d in range(10):
        rec[:, 10 - d] = ord("0") + (idx // 10 ** d) % 10
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + read_len] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, read_len))]
    rec[:, 12 + read_len:15 + read_len] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 15 + read_len:15 + 2 * read_len] = ord("I")
    rec[:, -1] = ord("\n")
    rec.tofile(path)
    return rec[:, 12:12 + read_len]


def main():
    if len(sys.argv) > 2:                                     # child: time one setting
        from graph_kmer_index_b200.read_kmers import FastxFile
        path = sys.argv[2]
        best = None
        for _ in range(4):
            t0 = time.perf_counter()
            f = FastxFile(path)
            dt = time.perf_counter() - t0
            n = f.n_reads
            offsets, lengths = f.lines()
            assert n == int(sys.argv[1]) and lengths.min() == 150 and lengths.max() == 150 and offsets[1] - offsets[0] == 316
            f.close()
            best = dt if best is None else min(best, dt)
        print(json.dumps({"pack_threads": os.environ.get("GKI_PACK_THREADS"), "file_gb": os.path.getsize(path) / 1e9, "index_ms": best * 1e3,
                          "gb_per_s": os.path.getsize(path) / best / 1e9, "cpus": os.cpu_count()}), flush=True)
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    path = "/tmp/gki_fastx_bench.fq"
    write_fastq(path, n)
    try:
        for threads in ("1", "3", "7", "14"):
            subprocess.run([sys.executable, os.path.abspath(__file__), str(n), path], env=dict(os.environ, GKI_PACK_THREADS=threads, GKI_FASTX_DEBUG="1"), check=True)
    finally:
        os.remove(path)


if __name__ == "__main__":
    main()
