#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_n1_r2_final.json 2> gpurun_out/bench_c2_n1_r2_final.err; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_c2_n1_r2_final.err
CMD3="python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD3 > gpurun_out/plain25c.log 2>&1 && ncu --set full --clock-control none -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c3_r2 $CMD3 > gpurun_out/ncu25c.log 2>&1
tail -n 1 gpurun_out/ncu25c.log
