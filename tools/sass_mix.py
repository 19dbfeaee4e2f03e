#!/usr/bin/env python3
"""Executed-instruction mix by opcode from an `ncu -i rep --page source --csv --print-source cuda,sass` export
(first file section).  usage: tools/sass_mix.py src.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
ends = [i for i, r in enumerate(rows) if r and r[0] == 'File Path' and i > hi]
end = ends[0] if ends else len(rows)
hdr = rows[hi]; samp = hdr.index('# Samples'); inst = hdr.index('Instructions Executed')
def I(x):
    try: return int(x)
    except ValueError: return 0
uniq = {}
for r in rows[hi + 1:end]:
    if len(r) < len(hdr): continue
    if r[0] in ('', '-') and r[2] not in uniq: uniq[r[2]] = (r[3].strip(), I(r[inst]), I(r[samp]))
ops = collections.Counter(); sm = collections.Counter(); tot = 0; ts = 0
for a, (src, n, s) in uniq.items():
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src
    ops[op] += n; sm[op] += s; tot += n; ts += s
print('instructions executed (warp-level):', tot, ' samples:', ts)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%-10s %6.2f%% inst  %6.2f%% samples' % (op, 100 * n / tot, 100 * sm[op] / ts))
