#!/bin/bash
# Run under gpurun (one GPU): plain run first (must exit 0), then ONE `--set full` capture per kernel group.
# The launch list comes from a separate command:  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
#     --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
mkdir -p gpurun_out
CMD="python tools/prof_kernels.py"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
# setup launches first: 1 Jacobian assembly for the AMG set-up; then the step: skip to the second Newton iteration
ncu --set full --clock-control none --import-source on -k regex:'^(assemble_rows_kernel|spmv_kernel|multidot_kernel|multiaxpy_kernel)$' \
    --launch-skip 12 -c 10 -o gpurun_out/prof_krylov -f $CMD > gpurun_out/ncu_full1.log 2>&1
tail -n 2 gpurun_out/ncu_full1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:'^(ssell_kernel|ssell_mult_split_kernel|first_sweep_kernel|coarse_vcycle_kernel)$' \
    --launch-skip 60 -c 20 -o gpurun_out/prof_mg -f $CMD > gpurun_out/ncu_full2.log 2>&1
tail -n 2 gpurun_out/ncu_full2.log | cut -c1-200
ls -la gpurun_out | head -30
