#!/bin/bash
# Run under gpurun: plain run first, then the launch list and one full capture of the two top kernels.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'assemble_rows|spmv_kernel' -s 6 -c 4 -o gpurun_out/prof_top $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out
