#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r2a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2a.err; cut -c1-1500 gpurun_out/bench_r2a.json
