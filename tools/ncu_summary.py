"""Prints the roofline-relevant metrics of an .ncu-rep (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'smsp__cycles_active.avg', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l and not l.startswith('==')]))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('--- %s  (id %s)' % (r[hdr.index('Kernel Name')][:90], r[0]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-68s %18s %s' % (w, r[i][:18], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and 'not_issued' not in h:
                try:
                    st.append((float(r[i]), h, units[i]))
                except ValueError:
                    pass
        for v, h, u in sorted(st, reverse=True)[:7]:
            print('  stall %-62s %18.3f %s' % (h.replace('smsp__average_warp', '').replace('smsp__', '')[:62], v, u))


if __name__ == '__main__':
    main(sys.argv[1])
