#!/bin/bash
# registers / spills / shared memory per kernel of one .cu file:  tools/ptxas_info.sh flof_solve.cu [filter]
cd "$(dirname "$0")/../ofblend_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xptxas -v -c "$1" -o /dev/null 2>&1 | python3 -c '
import sys, re, subprocess
name = None; rows = []
for line in sys.stdin:
    m = re.search(r"Compiling entry function .(\S+?). for", line)
    if m: name = m.group(1); spill = ""; continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: spill = "stack %s spill %s/%s" % m.groups(); continue
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        rows.append((name, int(m.group(1)), spill, sm.group(1) if sm else "0")); name = None
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for (n, regs, spill, sm), dn in zip(rows, names):
    dn = re.sub(r"\(.*", "", dn).replace("void ", "")
    if flt in dn: print("%-42s regs %3d  smem %6s  %s" % (dn, regs, sm, spill))
' "$2"
