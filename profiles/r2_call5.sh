#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r2a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2a.err; cut -c1-1500 gpurun_out/bench_r2a.json
for h in 0 101 102 202; do echo "scatter variant $h"; GKI_SLAB_HINT=$h python profiles/build_only.py 60000000 slab all1 2>&1 | cut -c1-140; done > gpurun_out/slab_variants.log 2>&1
cat gpurun_out/slab_variants.log
