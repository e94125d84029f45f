#!/bin/bash
python -m pytest tests/test_gpu_index.py -x -q -k "node_counts_with_more or uint16 or full_config1 or counter_past" 2>&1 | tail -n 2
for mb in 48 64 96; do GKI_NODE_SLICE_MB=$mb python bench.py --config c3 --steps 5 --warmup 2 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slice_mb', $mb, d['value'], d['ms_per_step'], d['stages_ms'], d['parity_checked']['equal'])"; done
