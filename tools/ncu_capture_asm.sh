#!/bin/bash
# ONE `--set full` capture of the Jacobian / residual row kernels (run under gpurun after tools/prof_kernels.py exits 0)
mkdir -p gpurun_out
CMD="python tools/prof_kernels.py"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'^(assemble_rows_kernel|nodal_exp6_kernel|block_jacobi_kernel)$' \
    --launch-skip 2 -c 6 -o gpurun_out/prof_asm -f $CMD > gpurun_out/ncu_full3.log 2>&1
tail -n 2 gpurun_out/ncu_full3.log | cut -c1-200
