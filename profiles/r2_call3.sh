#!/bin/bash
python -m pytest tests/test_gpu_index.py tests/test_gpu_api_misc.py -x -q -k "partitioned or counter_has or device_copy or kmer_index2" > gpurun_out/tests3.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests3.log
python profiles/build_only.py 60000000 slab all1 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:slab_ -c 4 -o gpurun_out/prof_slab_v1 python profiles/build_only.py 60000000 slab all1 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
