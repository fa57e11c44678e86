#!/bin/bash
# round-2 ncu evidence for profiles/ (one GPU): launch list of the bench command + full captures of the top kernels.
# Numbers printed by runs under ncu are never bench values.
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 9560 -c 3200 --csv -f --log-file gpurun_out/${TAG}_launches_bench64.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-res128 --no-tree-record > gpurun_out/${TAG}_launches_bench64.out 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${TAG}_launches_bench64.csv)"
ncu --set full --import-source on --clock-control none -k regex:"k_cg_|k_dot_seq|k_seq_resolve" -s 70 -c 7 -f -o gpurun_out/${TAG}_full_cg \
    python tools/bench_kernel.py 64 cg > gpurun_out/${TAG}_full_cg.log 2>&1; echo "cg rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_cv_expol_items -s 30 -c 1 -f -o gpurun_out/${TAG}_full_expol64 \
    python tools/bench_kernel.py 64 expol > gpurun_out/${TAG}_full_expol64.log 2>&1; echo "expol64 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_cv_expol_items -s 30 -c 1 -f -o gpurun_out/${TAG}_full_expol128 \
    python tools/bench_kernel.py 128 expol > gpurun_out/${TAG}_full_expol128.log 2>&1; echo "expol128 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k_cg_apply" -s 10 -c 1 -f -o gpurun_out/${TAG}_full_apply128 \
    python tools/bench_kernel.py 128 cg > gpurun_out/${TAG}_full_apply128.log 2>&1; echo "apply128 rc=$?"
