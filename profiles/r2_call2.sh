#!/bin/bash
python profiles/calibrate_store_groups.py > gpurun_out/calibrate_store_groups.jsonl 2> gpurun_out/calibrate_store_groups.err; tail -2 gpurun_out/calibrate_store_groups.err
bash profiles/r2_build_call.sh
