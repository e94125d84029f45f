#!/bin/bash
python -m pytest tests/test_gpu_index.py tests/test_gpu_finder.py -x -q -k "partitioned or config5" > gpurun_out/tests7.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests7.log
python profiles/k1_only.py > gpurun_out/k1_only.jsonl 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hash_reads -c 4 -o gpurun_out/prof_hash_r2 python profiles/k1_only.py > gpurun_out/ncu7.log 2>&1
cat gpurun_out/k1_only.jsonl; tail -2 gpurun_out/ncu7.log
