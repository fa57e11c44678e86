#!/bin/bash
# times the extrapolation-sweep kernels: dense (mode 2), Vec4 work list (mode 1), component planes (mode 0, variants)
RES=${1:-64}
FLOF_EXPOL_MODE=2 python tools/bench_kernel.py $RES expol 2>&1 | grep expol | sed 's/^/dense     /'
FLOF_EXPOL_MODE=1 python tools/bench_kernel.py $RES expol 2>&1 | grep expol | sed 's/^/vec4 list /'
for v in ${VARIANTS:-0 1 2}; do
  FLOF_EXPOL_MODE=0 FLOF_EXPOL_VARIANT=$v python tools/bench_kernel.py $RES expol 2>&1 | grep -E "expol" | sed "s/^/planes v$v /"
done
