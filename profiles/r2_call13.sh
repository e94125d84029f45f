#!/bin/bash
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
GKI_PIPELINE_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c3_n8_r2.json 2> gpurun_out/bench_c3_n8_r2.err; echo "bench n8 rc=$?"; grep -v "gki pipeline\|Warning\|warn" gpurun_out/bench_c3_n8_r2.err | tail -5; grep "gki pipeline" gpurun_out/bench_c3_n8_r2.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_n8_r2.json').read().strip().splitlines()[-1])
print(d['config']['workload'][:60]); print(d['value'], d['ms_per_step'], d['stages_ms'])
e=d['e2e']; print('e2e', e['value'], 'host_frac', e['host_frac'], 'host_read', e['host_read_gbs'], 'pcie', e['pcie_h2d_gbs'], 'lanes', e['pack_lanes'], e.get('pageable_numpy'), e.get('packed_2bit'))
print(d['index_build_partitioned'])
PY
