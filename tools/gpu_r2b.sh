#!/bin/bash
# round-2 check: seq-sum tests, CG micro-benchmarks per apply variant, bench 64 / 128 (every step under its own timeout)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seqsum.py -x -q > gpurun_out/r2b_seqtests.txt 2>&1; echo "seqtests rc=$?"
tail -5 gpurun_out/r2b_seqtests.txt
for v in 7 5 3 1; do
  echo "== apply_variant $v"
  FLOF_APPLY_VARIANT=$v timeout 300 python tools/bench_kernel.py 64 cg 2>&1 | grep -v "^$"
done
echo "== 128 variant 7"; FLOF_APPLY_VARIANT=7 timeout 300 python tools/bench_kernel.py 128 cg 2>&1 | grep -v "^$"
echo "== 128 variant 5"; FLOF_APPLY_VARIANT=5 timeout 300 python tools/bench_kernel.py 128 cg 2>&1 | grep -v "^$"
echo "== 32"; timeout 300 python tools/bench_kernel.py 32 cg 2>&1 | grep -v "^$"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench64.json 2> gpurun_out/r2b_bench64.err; echo "bench64 rc=$?"
timeout 600 python bench.py --res 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench128.json 2> gpurun_out/r2b_bench128.err; echo "bench128 rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.txt 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2b_tests.txt
