"""Summarise ncu reports into a small text file for profiles/:
    python tools/ncu_summary.py <out.txt> <lib.so> <kernel>=<report.ncu-rep> ...
Per kernel: duration, DRAM bytes, instruction counts, issue / occupancy / lane efficiency, and the source lines with
the most stall samples (tools/ncu_lines.py)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def main():
    out, lib = sys.argv[1], sys.argv[2]
    with open(out, 'w') as f:
        for spec in sys.argv[3:]:
            kernel, rep = spec.split('=')
            txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
            rows = list(csv.reader(txt.splitlines()))
            hdr, units, vals = rows[0], rows[1], rows[2]
            f.write('=== %s  (%s)\n' % (kernel, rep))
            for i, h in enumerate(hdr):
                if h in WANT:
                    f.write('  %-62s %14s %s\n' % (h, vals[i], units[i]))
            lines = subprocess.run([sys.executable, 'tools/ncu_lines.py', rep, lib, kernel, '14'], stdout=subprocess.PIPE, text=True).stdout
            f.write(lines + '\n')
    print(open(out).read())


if __name__ == '__main__':
    main()
