summ() { python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
d=json.loads([l for l in t if l.startswith('{')][-1]); print(round(d['ms_per_step'],3), d['krylov_its_per_step'], d['newton_its_per_step'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('parity_vs_1gpu'))"; }
echo "== N=1 V(1,1)"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | summ
echo "== N=1 V(0,1)"; FEDMX_MG_V01=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | summ
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== N=2 check (V01 default)"; $TR tools/mgpu_check.py --golden --mode exact 2>&1 | grep mgpu_check; $TR tools/mgpu_check.py --golden --mode bench 2>&1 | grep mgpu_check
echo "== N=2 V(1,1)"; FEDMX_MG_V01=0 $TR bench.py --gpus 2 --steps 10 --warmup 3 --profile-detail --no-parity 2>/dev/null | summ
echo "== N=2 V(0,1)"; $TR bench.py --gpus 2 --steps 10 --warmup 3 --profile-detail 2>/dev/null | summ
