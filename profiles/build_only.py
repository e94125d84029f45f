"""gki_index_build twice (c2 sizes) -- target of the ncu launch list for K2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib, synthetic
n, modulo, k = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000, 452_930_477, 31
dev = torch.device("cuda")
glen = synthetic.genome_length(n, k)
genome = torch.empty(glen, dtype=torch.uint8, device=dev)
_lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
hashes = torch.empty(n, dtype=torch.int64, device=dev); nodes = torch.empty(n, dtype=torch.int32, device=dev)
ref = torch.empty(n, dtype=torch.int64, device=dev); af = torch.empty(n, dtype=torch.float32, device=dev)
_lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n // 10, k, _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), None)
h2i = torch.empty(modulo, dtype=torch.int32, device=dev); nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
o_k, o_r, o_n, o_a = torch.empty_like(hashes), torch.empty_like(ref), torch.empty_like(nodes), torch.empty_like(af)
o_f = torch.empty(n, dtype=torch.int16, device=dev)
for flags in (1, 1, 0):
    _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), n, modulo, flags, _lib.ptr(h2i), _lib.ptr(nkm),
              _lib.ptr(o_k), _lib.ptr(o_n), _lib.ptr(o_r), _lib.ptr(o_a), _lib.ptr(o_f), None, None)
torch.cuda.synchronize()
print("ok")
