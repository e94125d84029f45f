#!/bin/bash
for mb in default 750 1000 1500; do
  if [ "$mb" = "default" ]; then unset GKI_FILTER_MAX_MB; else export GKI_FILTER_MAX_MB=$mb; fi
  python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('filter_mb', '$mb', d['value'], d['ms_per_step'], d['stages_ms']['count_reads_kernel'], d['index']['device_bytes'])"
done
