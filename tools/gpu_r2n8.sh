#!/bin/bash
# N-GPU bench (64^4 line with the res128 sub-record) + kernel tables; usage: tools/gpu_r2n8.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_64.json 2> gpurun_out/r2_bench_n${N}_64.err
echo "bench N=$N rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_n${N}_64.json | head -24
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${N}_64.json'))
c=d['config']
print('parity_vs_n1', c.get('parity_vs_n1'), 'tree', (c.get('tree_dot_mode') or {}).get('value'))
r=c.get('res128',{})
print('res128', {k:r[k] for k in r if k!='kernels'})
for k in r.get('kernels',[]): print('   ', k['kernel'][:36], k['launches_per_step'], round(k['avg_launch_ms'],4), round(k['share'],3), k.get('frac'))
PY
grep -iE "error|timed out|Traceback" gpurun_out/r2_bench_n${N}_64.err | head -5
