#!/bin/bash
python -m pytest tests/test_gpu_index.py tests/test_gpu_ingest.py tests/test_gpu_api_misc.py -x -q -k "partitioned or node_counts_with_more or chunked or c_abi or packed" > gpurun_out/tests6.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests6.log
for h in 0 101 102 202; do echo "scatter variant $h"; GKI_SLAB_HINT=$h python profiles/build_only.py 60000000 slab all1 2>&1 | cut -c1-140; done > gpurun_out/slab_variants.log 2>&1
cat gpurun_out/slab_variants.log
python bench.py --config c3 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_r2b.json 2> gpurun_out/bench_c3_r2b.err; echo "bench c3 rc=$?"; tail -3 gpurun_out/bench_c3_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_r2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'], d['e2e'].get('packed_2bit'))
PY
for mb in 32 48 96; do GKI_NODE_SLICE_MB=$mb python bench.py --config c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slice_mb', $mb, d['ms_per_step'], d['stages_ms'])"; done
