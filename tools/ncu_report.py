"""Text summary of an ncu report (all kernels in it) for profiles/: headline metrics per kernel and the source lines
with the largest share of instructions + stall samples.   python tools/ncu_report.py X.ncu-rep [top] > out.txt"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']

rep = sys.argv[1]
top = sys.argv[2] if len(sys.argv) > 2 else '16'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
kn = hdr.index('Kernel Name')
print('report: %s (ncu --set full --clock-control none; cold-cache, serialised launches)' % rep)
for v in rows[2:]:
    print('=== %s' % v[kn])
    for i, h in enumerate(hdr):
        if h in WANT:
            print('  %-62s %16s %s' % (h, v[i], units[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
open('/tmp/_ncu_src.csv', 'w').write(src)
print(subprocess.run([sys.executable, 'tools/ncu_src.py', '/tmp/_ncu_src.csv', top], stdout=subprocess.PIPE, text=True).stdout)
