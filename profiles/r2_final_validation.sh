#!/bin/bash
# last validation of the round: full GPU test suite, smoke, the bench line (c2 + c3 record), the reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_n1_r2_final.json 2> gpurun_out/bench_c2_n1_r2_final.err; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_c2_n1_r2_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm_c2_r2.json 2> gpurun_out/bench_reference_arm_c2_r2.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_n1_r2_final.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'stages', {k: round(v, 3) for k, v in d['stages'].items() if 'frac' in k}, 'parity', d['parity_checked']['equal'])
print('build', d['index_build']['ms'], d['index_build']['rows_in_finder_order']['ms'], 'e2e build', d['index_build_e2e']['ms'], d['index_build_e2e']['first_call_ms'])
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['stages_ms'], c3['e2e']['value'], c3['parity_checked']['equal'], c3['index_build']['ms'], c3['index_build']['rows_in_finder_order']['ms'])
PY
