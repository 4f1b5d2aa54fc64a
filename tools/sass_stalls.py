"""Decode control words of a `cuobjdump -sass` listing: per-opcode stall-count histogram and the sum
of stall counts over the FFMA2-dense region (a lower bound on one warp's issue time).
    python tools/sass_stalls.py file.sass"""
import collections
import re
import sys

lines = open(sys.argv[1]).read().split('\n')
ins = []
i = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            toks = [t for t in m.group(2).split() if not t.startswith('@')]
            ins.append(dict(addr=m.group(1), text=m.group(2).strip(), op=toks[0], stall=(w >> 105) & 0xf, yld=(w >> 109) & 1,
                            wbar=(w >> 110) & 7, rbar=(w >> 113) & 7, wait=(w >> 116) & 0x3f, reuse=(w >> 122) & 0xf))
            i += 2
            continue
    i += 1
ff = [k for k, x in enumerate(ins) if x['op'] == 'FFMA2']
lo, hi = ff[0], ff[-1]
body = ins[lo:hi + 1]
print('instructions', len(ins), 'FFMA2 region', lo, hi, 'len', len(body))
ops = collections.Counter(x['op'] for x in body)
print('ops in region', ops.most_common(12))
print('sum of stall counts in region', sum(x['stall'] for x in body), ' (FFMA-pipe floor: %d)' % (2 * (ops['FFMA2'] + ops['FADD2'] + ops['FMUL2'])))
for op in ('FFMA2', 'FADD2', 'LDS.128', 'MOV'):
    print(op, 'stall hist', sorted(collections.Counter(x['stall'] for x in body if x['op'] == op).items()),
          'with wait mask', sum(1 for x in body if x['op'] == op and x['wait']))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    for x in body[k:k + 60]:
        print(x['addr'], x['text'][:72].ljust(72), 's%d y%d wb%d rb%d w%s r%s' % (x['stall'], x['yld'], x['wbar'], x['rbar'], bin(x['wait']), bin(x['reuse'])))
