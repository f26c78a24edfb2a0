#!/bin/bash
# Round-2 record run on ONE B200 (under gpurun): bench line, launch list of the same command, one --set full capture.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/prof_kernels.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'spmv_kernel|assemble_tiled_kernel' -s 2 -c 6 -o gpurun_out/r2_prof_final -f python tools/prof_kernels.py > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-160
