#!/bin/bash
for h in 0 1 3 4 8 9; do echo "hint $h"; GKI_SLAB_HINT=$h python profiles/build_only.py 60000000 slab all1 2>&1 | cut -c1-140; done > gpurun_out/slab_hints.log 2>&1
cat gpurun_out/slab_hints.log
