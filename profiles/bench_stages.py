"""Per-stage device-resident timings (CUDA events, warm): K1 hash_reads, K2 index build, K3 pieces.
Prints one JSON line per stage with the algorithmic bytes of SURVEY.md section 8(d) and the fraction of the measured
HBM peak.  Usage: python profiles/bench_stages.py [entries] [reads] [modulo]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_kmer_index_b200 import DeviceIndex, _lib, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
modulo = int(sys.argv[3]) if len(sys.argv) > 3 else 452_930_477
k, L, n_nodes = 31, 150, max(n // 10, 1)
peak = 6552.3
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
dev = torch.device("cuda")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def report(stage, ms, units, unit_name, alg_bytes, **extra):
    gbs = alg_bytes / (ms / 1e3) / 1e9
    print(json.dumps(dict(stage=stage, ms=ms, rate=units / (ms / 1e3), unit=unit_name + "/s", algorithmic_bytes=alg_bytes,
                          achieved_gbs=gbs, frac_of_measured_hbm=gbs / peak, **extra)), flush=True)


glen = synthetic.genome_length(n, k)
genome = torch.empty(glen, dtype=torch.uint8, device=dev)
_lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
hashes = torch.empty(n, dtype=torch.int64, device=dev)
nodes = torch.empty(n, dtype=torch.int32, device=dev)
ref = torch.empty(n, dtype=torch.int64, device=dev)
af = torch.empty(n, dtype=torch.float32, device=dev)
_lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), None)
reads = torch.empty((R, L), dtype=torch.uint8, device=dev)
_lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, R, L, 100, 0, _lib.ptr(reads), None)
torch.cuda.synchronize()

# ---- K1: hash a chunk of reads, forward + reverse complement (unfused API) ----
Rh = min(R, 2_000_000)
nk = L - k + 1
fwd = torch.empty((Rh, nk), dtype=torch.int64, device=dev)
rc = torch.empty((Rh, nk), dtype=torch.int64, device=dev)
ms = timed(lambda: _lib.call("gki_hash_reads", _lib.ptr(reads), Rh, L, L, k, _lib.ptr(fwd), _lib.ptr(rc), None))
report("K1 hash_reads fwd+rc", ms, Rh * nk * 2, "hashes", Rh * (L + 2 * nk * 8), reads=Rh)
ms = timed(lambda: _lib.call("gki_hash_reads", _lib.ptr(reads), Rh, L, L, k, _lib.ptr(fwd), None, None))
report("K1 hash_reads fwd only", ms, Rh * nk, "hashes", Rh * (L + nk * 8), reads=Rh)
flat = fwd.view(-1)
ms = timed(lambda: _lib.call("gki_revcomp_hashes", _lib.ptr(flat), flat.numel(), k, _lib.ptr(rc), None))
report("K1 revcomp_hashes", ms, flat.numel(), "hashes", flat.numel() * 16)
del fwd, rc, flat

# ---- K2: index build (payload = kmers, nodes, ref_offsets, allele_frequencies) ----
h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
o_k, o_r = torch.empty_like(hashes), torch.empty_like(ref)
o_n, o_a = torch.empty_like(nodes), torch.empty_like(af)
o_f = torch.empty(n, dtype=torch.int16, device=dev)


def build(flags):
    _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), n, modulo, flags, _lib.ptr(h2i), _lib.ptr(nkm),
              _lib.ptr(o_k), _lib.ptr(o_n), _lib.ptr(o_r), _lib.ptr(o_a), _lib.ptr(o_f), None, None)


ms = timed(lambda: build(1), reps=3, warm=1)
report("K2 index_build skip_frequencies", ms, n, "entries", 50 * n + 8 * modulo, entries=n, modulo=modulo)
ms = timed(lambda: build(0), reps=3, warm=1)
report("K2 index_build with frequencies", ms, n, "entries", 50 * n + 8 * modulo, entries=n, modulo=modulo)

# ---- K3 pieces ----
index = DeviceIndex(h2i, nkm, o_k, o_n, modulo)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
index.prepare_counting(k)
b.record()
torch.cuda.synchronize()
print(json.dumps(dict(stage="K3 prepare_counting (filter + table build, once per index)", ms=a.elapsed_time(b))), flush=True)
ms = timed(lambda: index.count_reads(reads, k, True))
print(json.dumps(dict(stage="K3 count_reads fused", ms=ms, rate=R * nk * 2 / (ms / 1e3), unit="kmers/s")), flush=True)
counts = torch.zeros(n_nodes, dtype=torch.float64, device=dev)
ms = timed(lambda: index.node_counts(n_nodes, out=counts))
print(json.dumps(dict(stage="K3 node_counts", ms=ms, rate=n / (ms / 1e3), unit="entries/s")), flush=True)
ms = timed(lambda: index.reset_counts())
print(json.dumps(dict(stage="K3 reset_counts", ms=ms)), flush=True)
q = torch.empty((min(R, 1_000_000), nk), dtype=torch.int64, device=dev)
_lib.call("gki_hash_reads", _lib.ptr(reads), q.shape[0], L, L, k, _lib.ptr(q), None, None)
qq = q.view(-1)
ms = timed(lambda: index.count_kmers(qq))
print(json.dumps(dict(stage="K3 count_kmers unfused (forward hashes)", ms=ms, rate=qq.numel() / (ms / 1e3), unit="kmers/s")), flush=True)
