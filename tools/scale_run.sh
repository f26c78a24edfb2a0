#!/bin/bash
# Multi-GPU record run (under `gpurun --gpus 8`): partitioned parity at 8 ranks, strong-scaling bench lines at 8 / 4 / 2 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ $TR --nproc-per-node 8 --master-port 29521 tools/mgpu_check.py --golden --mode bench 2>&1 | grep "mgpu_check world"
  $TR --nproc-per-node 8 --master-port 29522 tools/mgpu_check.py --golden --mode exact 2>&1 | grep "mgpu_check world"
  $TR --nproc-per-node 8 --master-port 29524 tools/mgpu_check.py --golden --mode bench --precond ilu0 2>&1 | grep "mgpu_check world"
  $TR --nproc-per-node 4 --master-port 29523 tools/script_mgpu_check.py 2>&1 | grep "script_mgpu_check world"; } | tee gpurun_out/r2_mgpu_parity_8gpu.log
for N in 8 4 2; do
  $TR --nproc-per-node $N --master-port $((29530+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu.json').read().splitlines() if l.startswith('{')][-1]); print($N, round(d['value'],2), round(d['ms_per_step'],3), d['krylov_its_per_step'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['parity_vs_1gpu']['max_rel_diff'], d['setup_s'])"
done
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --profile-detail --no-parity > gpurun_out/r2_bench_8gpu_detail.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_8gpu_detail.json').read().splitlines() if l.startswith('{')][-1]); print('detail 8', round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
