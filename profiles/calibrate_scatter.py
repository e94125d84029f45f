"""Random store / atomic rates (the measurements the index-build design rests on)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib  # noqa: E402

NAMES = {0: "random 32B stores", 1: "random 16B stores", 2: "random 8B stores", 3: "returning atomics", 4: "RED histogram",
         5: "returning atomic + 32B store into the bin", 6: "random 32B stores, one 256-bit store each"}
for n in (60_000_000, 500_000_000):
    for mode in (0, 6, 1, 2):
        ms = ctypes.c_float()
        _lib.call("gki_calibrate_scatter", n, mode, 0, ctypes.byref(ms))
        print(json.dumps(dict(n=n, op=NAMES[mode], ms=ms.value, g_per_s=n / ms.value / 1e6)), flush=True)
    for bins in (1 << 20, 1 << 21, 1 << 23, 1 << 25):
        for mode in (3, 4, 5):
            ms = ctypes.c_float()
            _lib.call("gki_calibrate_scatter", n, mode, bins, ctypes.byref(ms))
            print(json.dumps(dict(n=n, op=NAMES[mode], bins_mb=bins * 4 / 2 ** 20, ms=ms.value, g_per_s=n / ms.value / 1e6)), flush=True)
