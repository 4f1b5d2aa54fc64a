"""Summarise an `ncu --page source --csv --print-source sass` export: stall totals, samples per
opcode, and samples per 100-instruction region.   python tools/ncu_src_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for k, hi in enumerate(hdr_idx[:1]):
    h = rows[hi]
    end = hdr_idx[k + 1] - 1 if k + 1 < len(hdr_idx) else len(rows)
    body = rows[hi + 1:end]
    ci = {n: i for i, n in enumerate(h)}
    num = lambda r, n: int(float(r[ci[n]] or 0))
    tot = sum(num(r, '# Samples') for r in body)
    print('instructions', len(body), 'samples', tot)
    agg = collections.Counter()
    for r in body:
        toks = [t for t in r[ci['Source']].split() if not t.startswith('@')]
        agg[toks[0] if toks else ''] += num(r, '# Samples')
    print('by opcode:', [(k2, v, round(100 * v / tot, 1)) for k2, v in agg.most_common(12)])
    for name in h:
        if name.startswith('stall_') and 'Not Issued' not in name:
            v = sum(num(r, name) for r in body)
            if v:
                print('%-24s %8d %5.1f%%' % (name, v, 100 * v / tot))
    reg = collections.Counter(); regn = collections.Counter(); regb = collections.Counter(); regl = collections.Counter()
    for i, r in enumerate(body):
        reg[i // 100] += num(r, '# Samples'); regn[i // 100] += num(r, 'stall_no_inst')
        regb[i // 100] += num(r, 'stall_barrier'); regl[i // 100] += num(r, 'stall_long_sb')
    print('region(100 instr): samples / no_inst / barrier / long_sb')
    for k2 in sorted(reg):
        print(k2, reg[k2], regn[k2], regb[k2], regl[k2], body[k2 * 100][ci['Source']][:50])
