#!/bin/bash
python profiles/build_only.py 60000000 window all1 > gpurun_out/plain11.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:window_build -c 1 -o gpurun_out/prof_window_v1 python profiles/build_only.py 60000000 window all1 > gpurun_out/ncu11.log 2>&1
tail -2 gpurun_out/ncu11.log
