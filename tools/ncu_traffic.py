"""Per-launch DRAM traffic of the kernels in an `ncu --set full` report -> JSON that bench.py reads for
roofline.traffic.   python tools/ncu_traffic.py report.ncu-rep out.json"""
import csv
import json
import re
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('<unnamed>::', '').replace('bas_render_detail::', '')
    val = lambda k: float(d[k]) * scale[units[hdr.index(k)]]
    res[name] = {'dram_bytes_read': val('dram__bytes_read.sum'), 'dram_bytes_write': val('dram__bytes_write.sum'),
                 'duration_us_under_ncu': float(d['gpu__time_duration.sum']) * {'us': 1, 'ns': 1e-3, 'ms': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
                 'registers': int(d['launch__registers_per_thread']), 'grid': int(d['launch__grid_size'])}
json.dump({'source': sys.argv[1].split('/')[-1], 'command': 'bench.py --steps 4 --warmup 3 --no-cpu --in-flight 1 (ncu --set full --clock-control none)',
           'kernels': res}, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(res, indent=1))
