#!/bin/bash
# one GPU call for the slab build: parity tests of the build paths, then timings (60 M and 1 B entries) and the slab-mean sweep
mkdir -p gpurun_out
python -m pytest tests/test_gpu_index.py -x -q -k "build or partitioned" > gpurun_out/build_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/build_tests.log
tail -5 gpurun_out/build_tests.log
python profiles/build_only.py 60000000 > gpurun_out/build_60m.jsonl 2> gpurun_out/build_60m.err; tail -2 gpurun_out/build_60m.err
for m in 1024 1280 1792; do GKI_SLAB_MEAN=$m python profiles/build_only.py 60000000 slab >> gpurun_out/build_60m_sweep.jsonl 2>> gpurun_out/build_60m.err; done
python profiles/build_only.py 1000000000 slab,binned > gpurun_out/build_1b.jsonl 2> gpurun_out/build_1b.err; tail -2 gpurun_out/build_1b.err
for m in 1024 1792; do GKI_SLAB_MEAN=$m python profiles/build_only.py 1000000000 slab >> gpurun_out/build_1b_sweep.jsonl 2>> gpurun_out/build_1b.err; done
cat gpurun_out/build_60m.jsonl gpurun_out/build_60m_sweep.jsonl gpurun_out/build_1b.jsonl gpurun_out/build_1b_sweep.jsonl | cut -c1-230
