#!/bin/bash
# full GPU suite + default bench (64^4 with the res128 / tree_dot_mode sub-records)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.txt 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2g_tests.txt
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --write-n1-record > gpurun_out/r2g_bench64.json 2> gpurun_out/r2g_bench64.err; echo "bench64 rc=$?"
tail -3 gpurun_out/r2g_bench64.err
cp tests/golden/bench_n1_record.json gpurun_out/ 2>/dev/null
python tools/show_bench.py gpurun_out/r2g_bench64.json 2>&1 | head -24
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench64.json'))
c=d['config']
print('tree', c.get('tree_dot_mode'))
r=c.get('res128',{})
print('res128', {k:r[k] for k in r if k!='kernels'})
for k in r.get('kernels',[]): print('   ',k)
PY
