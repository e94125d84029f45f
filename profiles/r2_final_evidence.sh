#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_n1_r2_final.json 2> gpurun_out/bench_c2_n1_r2_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_c2_n1_r2_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm_c2_r2.json 2> gpurun_out/bench_reference_arm_c2_r2.err; echo "ref rc=$?"
CMD1="python bench.py --no-c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD1 > gpurun_out/plain19a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_r2.csv $CMD1 > gpurun_out/ncu19a.log 2>&1
$CMD1 > gpurun_out/plain19b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c2_r2 $CMD1 > gpurun_out/ncu19b.log 2>&1
CMD3="python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD3 > gpurun_out/plain19c.log 2>&1 && ncu --set full --clock-control none -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c3_r2 $CMD3 > gpurun_out/ncu19c.log 2>&1
tail -1 gpurun_out/ncu19a.log gpurun_out/ncu19b.log gpurun_out/ncu19c.log
