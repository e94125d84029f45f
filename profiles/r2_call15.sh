#!/bin/bash
for p in 1 3; do GKI_COUNT_PIPE=$p python -m pytest tests/test_gpu_index.py tests/test_gpu_ingest.py tests/test_gpu_fuzz.py -x -q -k "count_reads or packed or full_config1 or fuzz or host_pipeline" 2>&1 | tail -2; done
for p in 0 1 3; do GKI_COUNT_PIPE=$p python bench.py --no-c3 --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 pipe', $p, d['ms_per_step'], d['stages_ms']['count_reads_kernel'])"; done
for p in 0 1 3; do GKI_COUNT_PIPE=$p python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 pipe', $p, d['ms_per_step'], d['stages_ms']['count_reads_kernel'])"; done
