"""bench.py's reference arm and synthetic inputs on the CPU: the JSON line the driver parses, and that the arm does
not load the product library (the reference arm mapped libbas_b200.so)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line_without_the_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1']; "
            "runpy.run_path(%r, run_name='__main__'); "
            "import ctypes; maps = open('/proc/self/maps').read(); "
            "print('MAPPED' if 'libbas_b200' in maps or 'libbas_probe' in maps else 'CLEAN')" % os.path.join(ROOT, 'bench.py'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert out.stdout.strip().endswith('CLEAN')
    assert line['impl'] == 'reference' and line['higher_is_better'] is True and line['unit'] == 'sample-pairs/s'
    assert line['metric'] == 'binaural output sample-pairs/s' and line['value'] > 0
    assert line['config']['workload'].startswith('configs[2]: 64 independent 60 s sources')
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['gpu_launches'] == 0


def test_workload_names_follow_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert len(base['configs']) == 5
    for name, (idx, n_src, secs, fs, keep, ups, _) in bench.CONFIGS.items():
        assert bench.workload_name(name).startswith('configs[%d]' % idx)
    assert bench.CONFIGS['stress1024'][1:6] == (1024, 60, 44100, 512, 16)          # configs[4]: full-length IRs, N = 16 bank
    assert bench.CONFIGS['hour'][2:4] == (3600, 48000)                             # configs[3]
    assert bench.ALIASES == {'2': 'single', '3': 'mix64', '4': 'hour', '5': 'stress1024'}
