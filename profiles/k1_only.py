"""K1 alone (gki_hash_reads on 2 M x 150 bp device-resident reads, k=31): fwd+rc and fwd-only, CUDA-event times and the fraction of the
measured HBM peak under SURVEY 8(d)'s byte model -- the ncu target for the hashing kernel.  Usage: python profiles/k1_only.py [reads]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib, synthetic  # noqa: E402

R, L, k = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 150, 31
peak = 6552.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
dev = torch.device("cuda")
glen = synthetic.genome_length(1_000_000, k)
genome = torch.empty(glen, dtype=torch.uint8, device=dev)
_lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
reads = torch.empty((R, L), dtype=torch.uint8, device=dev)
_lib.call("gki_synth_reads", _lib.ptr(genome), glen, 0, R, L, 100, 0, _lib.ptr(reads), None)
fwd = torch.empty((R, L - k + 1), dtype=torch.uint64, device=dev)
rc = torch.empty((R, L - k + 1), dtype=torch.uint64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for name, out_rc, nbytes in (("fwd+rc", rc, R * (L + 2 * (L - k + 1) * 8)), ("fwd only", None, R * (L + (L - k + 1) * 8))):
    best = None
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call("gki_hash_reads", _lib.ptr(reads), R, L, L, k, _lib.ptr(fwd), _lib.ptr(out_rc), stream)
        b.record()
        torch.cuda.synchronize()
        if i:
            best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
    print(json.dumps(dict(stage="K1 hash_reads " + name, reads=R, ms=best, algorithmic_bytes=nbytes, achieved_gbs=nbytes / best / 1e6,
                          frac_of_measured_hbm=nbytes / best / 1e6 / peak)), flush=True)
