#!/bin/bash
python -m pytest tests/test_gpu_multi.py tests/test_gpu_api_misc.py -x -q -k "two_gpu or partitioned_counter" > gpurun_out/tests14.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests14.log
GKI_PIPELINE_DEBUG=1 python bench.py --no-c3 --steps 10 --warmup 3 > gpurun_out/bench_c2_r2c.json 2> gpurun_out/bench_c2_r2c.err; echo "bench rc=$?"; grep "gki pipeline" gpurun_out/bench_c2_r2c.err | head -24 | cut -c1-140
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_r2c.json').read().strip().splitlines()[-1])
e=d['e2e']; print(d['value'], 'e2e', e['value'], 'host_frac', e['host_frac'], 'host_read', e['host_read_gbs'], 'pcie', e['pcie_h2d_gbs'], e.get('pageable_numpy'), e.get('packed_2bit'))
print(d['index_build']['ms'], d['stages'])
PY
