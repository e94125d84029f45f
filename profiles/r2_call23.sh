#!/bin/bash
CMD3="python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD3 > gpurun_out/plain23.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c3_regions $CMD3 > gpurun_out/ncu23.log 2>&1
tail -n 2 gpurun_out/ncu23.log
