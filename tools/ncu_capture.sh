#!/bin/bash
# Run under gpurun: plain run first, then ONE full capture of the top kernels (assembly Jacobian, block SpMV, multidot).
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'assemble_rows_kernel.*StreamerModel|spmv_kernel<3|multidot_kernel|phi_residual' -c 10 \
    -o gpurun_out/prof_top $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 4 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
