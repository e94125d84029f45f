"""What bounds scattered 32-byte stores (~42 G/s, profiles/r1/calibrate_scatter.jsonl)?  Requests, sectors or DRAM:
lanes per run of consecutive records (1 = every record its own request, 4 = one whole line per request), destination size
(L2-resident vs HBM) and number of SMs issuing."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_kmer_index_b200 import _lib  # noqa: E402

n = 120_000_000
for slots, where in ((1 << 20, "32 MB (L2)"), (120_000_000, "3.8 GB (HBM)")):
    for ctas in (148, 74, 37):
        for group in (1, 2, 4, 8, 32):
            ms = ctypes.c_float()
            _lib.call("gki_calibrate_store_groups", n, group, slots, ctas, ctypes.byref(ms))
            print(json.dumps(dict(op="scattered 256-bit stores", records=n, lanes_per_run=group, destination=where, sms=ctas, ms=ms.value,
                                  g_records_per_s=n / ms.value / 1e6, g_requests_per_s=n / min(group, 4) / ms.value / 1e6,
                                  gbs=n * 32 / ms.value / 1e6)), flush=True)
