#!/bin/bash
# round-2 evidence: full GPU test suite, smoke, the bench line (c2 + c3 record) and the reference arm, ncu launch list of the c2 step,
# ncu --set full of the count kernel (c2, c3), of K1 (both strands, forward strand) and of the slab build kernels (shuffled rows, 60 M)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_n1_r2_final.json 2> gpurun_out/bench_c2_n1_r2_final.err; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_c2_n1_r2_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm_c2_r2.json 2> gpurun_out/bench_reference_arm_c2_r2.err; echo "ref rc=$?"
CMD1="python bench.py --no-c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD1 > gpurun_out/plainA.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_r2.csv $CMD1 > gpurun_out/ncuA.log 2>&1
$CMD1 > gpurun_out/plainB.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c2_r2 $CMD1 > gpurun_out/ncuB.log 2>&1
CMD3="python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD3 > gpurun_out/plainC.log 2>&1 && ncu --set full --clock-control none -k regex:count_reads_kernel -s 3 -c 1 -o gpurun_out/prof_count_c3_r2 $CMD3 > gpurun_out/ncuC.log 2>&1
CMDK="python profiles/k1_only.py"
$CMDK > gpurun_out/k1_only_r2.jsonl 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hash_reads_kernel -s 1 -c 1 -o gpurun_out/prof_hash_r2_both $CMDK > gpurun_out/ncuD.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hash_reads_kernel -s 5 -c 1 -o gpurun_out/prof_hash_r2_fwd $CMDK > gpurun_out/ncuE.log 2>&1
CMDB="python profiles/build_only.py 60000000 slab all1"
$CMDB > gpurun_out/plainF.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"slab_scatter_kernel|slab_finish_kernel|slab_sample" -s 4 -c 4 -o gpurun_out/prof_build_shuffled_60m $CMDB > gpurun_out/ncuF.log 2>&1
for f in A B C D E F; do tail -n 1 gpurun_out/ncu$f.log; done
