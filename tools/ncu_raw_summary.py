"""Key metrics of an ncu report (raw page).  python tools/ncu_raw_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__cycles_active.avg']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            print('%-75s %s %s' % (k, d[k], units[hdr.index(k)]))
    for h in hdr:
        if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(d[h] or 0) > 0.02:
            print('%-75s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), d[h]))
    print('---')
