#!/bin/bash
free -g | head -2; nproc
( time python bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err ) 2>&1 | tail -3; echo "rc=$?"; tail -3 gpurun_out/bench_ref_c3.err; cut -c1-900 gpurun_out/bench_ref_c3.json
OMP_NUM_THREADS=1 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 --no-cpu-baseline | cut -c1-300
python -m pytest tests/test_gpu_index.py tests/test_gpu_ingest.py -x -q -k "build_paths or host_pipeline or partitioned" 2>&1 | tail -3
