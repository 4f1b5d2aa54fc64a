"""Opcode histogram per kernel of the built libraries (cuobjdump -sass): what proves which hardware paths a kernel uses
(UTMALDG / UBLKCP = TMA, SYNCS = mbarrier, FFMA2 = packed FP32 FMA, UTC*MMA / LDTM = tcgen05 / TMEM, see B200_PROFILING.md).
    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARK = ('UTMALDG', 'UBLKCP', 'SYNCS', 'FFMA2', 'FFMA', 'UTCHMMA', 'UTCQMMA', 'LDTM', 'UTCBAR', 'LDGSTS', 'REDG', 'RED', 'ATOMG', 'HMMA', 'LDG', 'STG', 'LDS', 'STS', 'USETMAXREG', 'BAR', 'ACQBULK', 'MEMBAR', 'FENCE')
for libname in ('libbas_b200.so', 'libbas_probe.so'):
    path = os.path.join(ROOT, 'binaural-audio-synthesis_b200', libname)
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    print('=' * 100)
    print(libname)
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace('(anonymous namespace)::', '').replace('void ', '').replace('bas_render_detail::', '')
            name = re.sub(r'\(.*', '', name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ('SYNCS', 'UTMALDG', 'LDTM', 'UTCHMMA', 'UBLKCP'):
                cur[m.group(1) + m.group(2)] += 1
    for name, c in kernels.items():
        total = sum(v for k, v in c.items() if '.' not in k)
        marks = ', '.join('%s %d' % (k, c[k]) for k in MARK if c.get(k))
        detail = ', '.join('%s %d' % (k, v) for k, v in sorted(c.items()) if '.' in k)
        print('%-62s %6d instructions | %s%s' % (name[:62], total, marks, (' | ' + detail) if detail else ''))
