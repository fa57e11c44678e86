#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of the bench command + full captures of the top kernels.
# Numbers printed by runs under ncu are never bench values.
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 1560 --csv -f --log-file gpurun_out/${TAG}_launches_bench64.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench64.out 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${TAG}_launches_bench64.csv)"
ncu --set full --import-source on --clock-control none -k regex:k_cg_ -s 30 -c 3 -f -o gpurun_out/${TAG}_full_cg \
    python tools/bench_kernel.py 64 cg > gpurun_out/${TAG}_full_cg.log 2>&1; echo "cg rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_gauss_blur4d_tiled -s 2 -c 1 -f -o gpurun_out/${TAG}_full_gauss \
    python tools/bench_kernel.py 64 gauss > gpurun_out/${TAG}_full_gauss.log 2>&1; echo "gauss rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_project_cells -s 1 -c 1 -f -o gpurun_out/${TAG}_full_project \
    python tools/bench_kernel.py 64 project > gpurun_out/${TAG}_full_project.log 2>&1; echo "project rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_cv_expol_items -s 30 -c 1 -f -o gpurun_out/${TAG}_full_expol \
    python tools/bench_kernel.py 64 expol > gpurun_out/${TAG}_full_expol.log 2>&1; echo "expol rc=$?"
