#!/bin/bash
# usage: tools/mgpu_run.sh N tag [res]   -- multi-GPU self-check + bench on N GPUs (peer mailboxes vs NCCL)
N=${1:-2}; TAG=${2:-x}; RES=${3:-64}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 600 $TR --master-port 29611 tools/multigpu_check.py 32 48 > gpurun_out/mg_check_${TAG}.json 2> gpurun_out/mg_check_${TAG}.err; echo "check rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/mg_check_${TAG}.json').read().strip().splitlines()[-1])
    print('check ok', d['ok'], 'iters', d['no_projection']['iters_single'], d['no_projection']['iters_sharded'], 'rel', d['no_projection']['rel_l2'], d['with_projection']['rel_l2'])
except Exception as e:
    print('check parse failed', e)
PY
timeout 600 $TR --master-port 29621 bench.py --gpus $N --steps 3 --res $RES --no-cpu-baseline > gpurun_out/bench_${TAG}_p2p.json 2> gpurun_out/bench_${TAG}_p2p.err; echo "bench p2p rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_p2p.json | head -${SHOW:-22}
if [ -z "$SKIP_NCCL" ]; then
FLOF_NO_P2P=1 timeout 600 $TR --master-port 29631 bench.py --gpus $N --steps 3 --res $RES --no-cpu-baseline > gpurun_out/bench_${TAG}_nccl.json 2> gpurun_out/bench_${TAG}_nccl.err; echo "bench nccl rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_nccl.json | head -${SHOW:-22}
fi
tail -n 3 gpurun_out/*_${TAG}*.err | grep -iE "error|timed out" || true
