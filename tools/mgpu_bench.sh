#!/bin/bash
# usage: tools/mgpu_bench.sh N tag res steps   -- one bench.py run on N GPUs (one process per GPU)
N=${1:-2}; TAG=${2:-x}; RES=${3:-64}; STEPS=${4:-3}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
  bench.py --gpus $N --steps $STEPS --res $RES --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench N=$N res=$RES rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}.json | head -${SHOW:-20}
grep -iE "error|timed out|Traceback" gpurun_out/bench_${TAG}.err | head -5
