"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui, mi, idi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('Metric Name'), h.index('ID')
launches = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    d = launches.setdefault(r[idi], {'name': r[ki][:60]})
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if r[mi].startswith('gpu__time'):
        d['ms'] = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else v)
    else:
        d[r[mi]] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
sel = list(launches.values())
if len(sys.argv) > 2:      # print from the N-th occurrence of a kernel name fragment to the next
    frag, nth = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0
    idx = [i for i, d in enumerate(sel) if frag in d['name']]
    sel = sel[idx[nth]:idx[nth + 1] if nth + 1 < len(idx) else None]
tot = 0
for d in sel:
    rd, wr = d.get('dram__bytes_read.sum', 0) / 1e9, d.get('dram__bytes_write.sum', 0) / 1e9
    tot += d['ms']
    print(f"{d['ms']:8.3f} ms  rd {rd:6.2f} GB  wr {wr:6.2f} GB  {(rd + wr) / d['ms'] * 1e3 if d['ms'] else 0:7.0f} GB/s  {d['name']}")
print(f"{tot:8.3f} ms total")
