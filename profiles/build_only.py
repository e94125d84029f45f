"""gki_index_build at c2 sizes (target of the ncu launch list for K2), timed for both build paths and both record widths.
Usage: python profiles/build_only.py [entries]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib, synthetic  # noqa: E402

n, modulo, k = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000, 452_930_477, 31
dev = torch.device("cuda")
glen = synthetic.genome_length(n, k)
genome = torch.empty(glen, dtype=torch.uint8, device=dev)
_lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
hashes = torch.empty(n, dtype=torch.int64, device=dev)
nodes = torch.empty(n, dtype=torch.int32, device=dev)
ref = torch.empty(n, dtype=torch.int64, device=dev)
af = torch.empty(n, dtype=torch.float32, device=dev)
_lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n // 10, k, _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), None)
if os.environ.get("GKI_SYNTH_ORDER") == "finder":      # rows of a k-mer adjacent, as DenseKmerFinder emits them (the generator shuffles them apart)
    j = torch.arange(n, dtype=torch.int64, device=dev)
    pos = (j * synthetic.PERM_MULT + synthetic.PERM_ADD) % n
    order = torch.empty_like(j)
    order[pos] = j
    del j, pos
    hashes, nodes, ref, af = (c[order] for c in (hashes, nodes, ref, af))
    del order
    torch.cuda.empty_cache()
if os.environ.get("GKI_SYNTH_ORDER") == "finder4":     # four adjacent rows per k-mer (a k-mer whose path covers four nodes): every second k-mer of the finder order, twice
    j = torch.arange(n, dtype=torch.int64, device=dev)
    pos = (j * synthetic.PERM_MULT + synthetic.PERM_ADD) % n
    order = torch.empty_like(j)
    order[pos] = j
    del j, pos
    pick = order.view(-1, 4)[:, :2].repeat_interleave(2, dim=0).reshape(-1)[:n] if n % 4 == 0 else order
    hashes, ref, af = hashes[pick], ref[pick], af[pick]
    nodes = nodes[order]
    del order, pick
    torch.cuda.empty_cache()
h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
o_k, o_r, o_n, o_a = torch.empty_like(hashes), torch.empty_like(ref), torch.empty_like(nodes), torch.empty_like(af)
o_f = torch.empty(n, dtype=torch.int16, device=dev)


def build(columns, flags):
    full = columns == "all"
    _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref) if full else None, _lib.ptr(af) if full else None, n, modulo,
              flags, _lib.ptr(h2i), _lib.ptr(nkm), _lib.ptr(o_k), _lib.ptr(o_n), _lib.ptr(o_r) if full else None, _lib.ptr(o_a) if full else None,
              _lib.ptr(o_f) if full else None, None, None)


def compulsory_bytes(columns, flags):
    """bytes the call must move: every requested column read once and written once in bucket order, the frequency column
    written, both dense tables written (50 N + 8 modulo with all columns)"""
    per_entry = 24 if columns == "kmers+nodes" else 48 + 2
    return per_entry * n + 8 * modulo


paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["slab", "binned", "radix"]
only = sys.argv[3] if len(sys.argv) > 3 else None       # "all1": all columns, skip_frequencies (the ncu target)
sums = {}
for path in paths:
    os.environ["GKI_BUILD_PATH"] = path
    for columns, flags in (("kmers+nodes", 1), ("all", 1), ("all", 0)):
        if only and only != "%s%d" % (columns, flags):
            continue
        build(columns, flags)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            build(columns, flags)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        key = (columns, flags)
        check = (int(h2i.sum().item()), int(o_k.sum().item()), int(o_n.sum().item()), int(nkm.sum().item()))
        assert sums.setdefault(key, check) == check, (path, key)
        by = compulsory_bytes(columns, flags)
        print(json.dumps(dict(path=path, columns=columns, skip_frequencies=bool(flags), entries=n, ms=best, g_entries_per_s=n / best / 1e6,
                              compulsory_bytes=by, compulsory_gbs=by / best / 1e6, frac_of_measured_hbm=by / best / 1e6 / 6552.3,
                              slab_mean=os.environ.get("GKI_SLAB_MEAN"), order=os.environ.get("GKI_SYNTH_ORDER", "shuffled"))), flush=True)
