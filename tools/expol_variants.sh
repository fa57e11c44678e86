#!/bin/bash
# times the work-list extrapolation kernel variants (FLOF_EXPOL_VARIANT) next to the dense kernel
RES=${1:-64}
FLOF_EXPOL_DENSE=1 python tools/bench_kernel.py $RES expol 2>&1 | grep expol | sed 's/^/dense   /'
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  FLOF_EXPOL_VARIANT=$v python tools/bench_kernel.py $RES expol 2>&1 | grep -E "expol" | sed "s/^/var $v   /"
done
