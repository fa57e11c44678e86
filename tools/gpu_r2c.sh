#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seqsum.py -q > gpurun_out/r2d_seqtests.txt 2>&1; echo "seqtests rc=$?"
tail -5 gpurun_out/r2d_seqtests.txt
for v in 7 10 11 9 5 3 1; do
  echo "== apply_variant $v"
  FLOF_APPLY_VARIANT=$v timeout 300 python tools/bench_kernel.py 64 cg 2>&1 | grep -v "^$" | head -9
done
for v in 7 11 5; do
echo "== 128 variant $v"; FLOF_APPLY_VARIANT=$v timeout 300 python tools/bench_kernel.py 128 cg 2>&1 | grep -v "^$" | head -9
done
echo "== 32"; timeout 300 python tools/bench_kernel.py 32 cg 2>&1 | grep -v "^$" | head -9
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_dot_seq|k_seq_resolve" -s 60 -c 4 -f -o gpurun_out/r2d_full_seq python tools/bench_kernel.py 64 cg > gpurun_out/r2d_full_seq.log 2>&1; echo "ncu rc=$?"
