#!/bin/bash
# refresh of the full ncu captures of the CG kernels and the s=2 Gaussian blur (after kernel changes)
TAG=${1:-r1b}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_cg_ -s 30 -c 3 -f -o gpurun_out/${TAG}_full_cg \
    python tools/bench_kernel.py 64 cg > gpurun_out/${TAG}_full_cg.log 2>&1; echo "cg rc=$?"
ncu --set full --import-source on --clock-control none -k regex:k_gauss_blur4d_tiled -s 2 -c 1 -f -o gpurun_out/${TAG}_full_gauss \
    python tools/bench_kernel.py 64 gauss > gpurun_out/${TAG}_full_gauss.log 2>&1; echo "gauss rc=$?"
