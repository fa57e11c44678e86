#!/bin/bash
# round-2 state check: full GPU suite, bench 64 (+res128 record), CG micro-benchmarks, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2e_tests.txt 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2e_tests.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --write-n1-record > gpurun_out/r2e_bench64.json 2> gpurun_out/r2e_bench64.err; echo "bench64 rc=$?"
tail -3 gpurun_out/r2e_bench64.err
cp tests/golden/bench_n1_record.json gpurun_out/ 2>/dev/null
python tools/show_bench.py gpurun_out/r2e_bench64.json 2>&1 | head -60
for r in 64 128 32; do echo "== cg $r"; timeout 300 python tools/bench_kernel.py $r cg 2>&1 | grep -v "^$" | head -12; done
echo "== others 64";  timeout 300 python tools/bench_kernel.py 64 expol gauss advect project 2>&1 | grep -v "^$" | head -30
