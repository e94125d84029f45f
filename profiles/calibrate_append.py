"""Append-scatter rate vs number of bins: key load -> returning atomic on the bin cursor -> one 256-bit store at the cursor
(the slab scatter of the index build).  With few bins the open line of every cursor stays in L2 and leaves it whole."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib  # noqa: E402

for n in (60_000_000, 500_000_000):
    for bins in (2_000, 8_000, 20_000, 40_000, 100_000, 250_000, 500_000, 1_000_000, 4_000_000):
        ms = ctypes.c_float()
        _lib.call("gki_calibrate_scatter", n, 7, bins, ctypes.byref(ms))
        print(json.dumps(dict(n=n, op="append: key load + returning atomic + 256-bit store at the cursor", bins=bins,
                              frontier_mb=bins * 128 / 2 ** 20, ms=ms.value, g_per_s=n / ms.value / 1e6,
                              gbs=n * 40 / ms.value / 1e6)), flush=True)
