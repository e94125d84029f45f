#!/bin/bash
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/tests8.log 2>&1; echo "multi tests rc=$?"; tail -5 gpurun_out/tests8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_c3_n2_r2.json 2> gpurun_out/bench_c3_n2_r2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_c3_n2_r2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_n2_r2.json').read().strip().splitlines()[-1])
print(d['config']['workload'][:60]); print(d['value'], d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['value'], d['e2e']['host_frac'], d['e2e'].get('pageable_numpy'), d['index_build_partitioned'])
PY
