#!/bin/bash
python -m pytest tests/test_gpu_index.py tests/test_gpu_ingest.py tests/test_gpu_fuzz.py tests/test_gpu_api_misc.py -x -q -k "count or packed or full_config1 or fuzz or host_pipeline or counter or node_counts" 2>&1 | tail -3
GKI_BUILD_DEBUG=1 python bench.py --config c3 --no-e2e --steps 5 --warmup 2 2> gpurun_out/c3_mzhome.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['stages_ms'], d['parity_checked']['equal'], d['index']['device_bytes'])"
grep "prepare_counting" gpurun_out/c3_mzhome.err
python bench.py --no-c3 --no-e2e --no-cpu-baseline --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], d['stages_ms'])"
