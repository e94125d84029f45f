#!/bin/bash
python profiles/build_only.py 60000000 window all1 > gpurun_out/plain10.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:window_ -c 6 --csv --log-file gpurun_out/launches_window.csv python profiles/build_only.py 60000000 window all1 > gpurun_out/ncu10.log 2>&1
tail -2 gpurun_out/ncu10.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_window.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
