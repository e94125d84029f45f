"""Random 8-byte gather rate vs table size (the ceiling K3's probes are compared with)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib  # noqa: E402

n = 1 << 30
for mb in (8, 16, 24, 32, 48, 64, 96, 128, 256, 512, 1024, 4096, 16384):
    for mode in (0, 1, 2):
        ms = ctypes.c_float()
        _lib.call("gki_calibrate_random_gather", mb << 20, n, mode, ctypes.byref(ms))
        acc = n * (2 if mode & 1 else 1)
        print(json.dumps(dict(table_mb=mb, dependent=mode & 1, fill64=mode >> 1, ms=ms.value, g_gathers_per_s=acc / ms.value / 1e6)), flush=True)
ms = ctypes.c_float()
_lib.call("gki_calibrate_copy", 4 << 30, ctypes.byref(ms))
print(json.dumps(dict(copy_gb=4, ms=ms.value, gbs=2 * 4 * 1.073741824 / ms.value * 1e3)))
