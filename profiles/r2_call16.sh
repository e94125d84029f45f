#!/bin/bash
python -m pytest tests/test_gpu_index.py tests/test_gpu_api_misc.py tests/test_gpu_side_indexes.py -x -q -k "build or partitioned or frequenc or singleton or kmer_index2 or golden" > gpurun_out/tests16.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests16.log
python profiles/build_only.py 60000000 slab > gpurun_out/build_60m_freq.jsonl 2>&1; cut -c1-210 gpurun_out/build_60m_freq.jsonl
python profiles/build_only.py 1000000000 slab > gpurun_out/build_1b_freq.jsonl 2>&1; cut -c1-210 gpurun_out/build_1b_freq.jsonl
