#!/bin/bash
GKI_BUILD_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_index.py -x -q -k "build_paths or partitioned" > gpurun_out/tests9.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests9.log
GKI_BUILD_DEBUG=1 timeout 300 python profiles/build_only.py 60000000 window,slab > gpurun_out/build_60m_w.jsonl 2> gpurun_out/build_60m_w.err; echo "rc=$?"; tail -3 gpurun_out/build_60m_w.err; cut -c1-200 gpurun_out/build_60m_w.jsonl
GKI_BUILD_DEBUG=1 timeout 300 python profiles/build_only.py 1000000000 window,slab all1 > gpurun_out/build_1b_w.jsonl 2> gpurun_out/build_1b_w.err; echo "rc=$?"; tail -3 gpurun_out/build_1b_w.err; cut -c1-200 gpurun_out/build_1b_w.jsonl
GKI_BUILD_DEBUG=1 timeout 300 python profiles/build_only.py 250000000 window,slab all1 2>&1 | cut -c1-200
