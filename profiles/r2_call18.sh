#!/bin/bash
python -m pytest tests/test_gpu_index.py -x -q -k "node_counts_with_more or uint16" 2>&1 | tail -2
for mb in 64 48; do GKI_BUILD_DEBUG=1 GKI_NODE_SLICE_MB=$mb python bench.py --config c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/prep_c3_$mb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slice_mb', $mb, d['value'], d['ms_per_step'], d['stages_ms'])"; done
grep "prepare_counting" gpurun_out/prep_c3_64.err
