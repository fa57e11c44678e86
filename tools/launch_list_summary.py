"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel:
   python tools/launch_list_summary.py gpurun_out/r1_launches_bench64.csv > profiles/r1_launches_summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        v = float(r[ival].replace(",", ""))
        u = r[iunit]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
        tot[r[iname]] += ms
        cnt[r[iname]] += 1
    total = sum(tot.values())
    print("total_ms %.3f launches %d" % (total, sum(cnt.values())))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-100s n=%5d %10.3f ms %6.1f%%" % (k[:100], cnt[k], v, 100 * v / total))


if __name__ == "__main__":
    main(sys.argv[1])
