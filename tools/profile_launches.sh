#!/bin/bash
# ncu launch list of the bench command: one step's worth of launches (1538 per 64^4 step) after the warm-up steps
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 1560 --csv -f --log-file gpurun_out/${1:-r1}_launches_bench64.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${1:-r1}_launches_bench64.out 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${1:-r1}_launches_bench64.csv)"
