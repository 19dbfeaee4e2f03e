#!/usr/bin/env python3
"""Per-opcode and per-stall-reason summary of an `ncu -i rep --page source --csv --print-source sass` export.
usage: tools/sass_profile.py sass.csv [top_n_hot_instructions]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
def I(x):
    try: return int(x)
    except ValueError: return 0
ops = collections.Counter(); smp = collections.Counter(); thr = collections.Counter(); tot = ts = 0
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
hot = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[col['Source']].strip()
    n, s, t = I(r[col['Instructions Executed']]), I(r[col['# Samples']]), I(r[col['Thread Instructions Executed']])
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src
    if op in ('LDG', 'LDS', 'LD', 'ST', 'STG', 'STS', 'LDL', 'STL'):
        op = '.'.join(m.group(2).split('.')[:1]) + ('.128' if '.128' in src else '.64' if '.64' in src else '')
    ops[op] += n; smp[op] += s; thr[op] += t; tot += n; ts += s
    for h in stall_cols: stalls[h] += I(r[col[h]])
    hot.append((s, n, t / n if n else 0, r[col['Address']][-5:], src))
print('warp-level instructions executed:', tot, ' samples:', ts)
print('%-10s %8s %8s %8s' % ('opcode', 'inst %', 'samp %', 'lanes'))
for op, n in ops.most_common(28):
    print('%-10s %7.2f%% %7.2f%% %8.1f' % (op, 100 * n / tot, 100 * smp[op] / ts, thr[op] / n if n else 0))
print('stall reasons (all samples):')
for h, v in stalls.most_common(10): print('  %-24s %6.2f%%' % (h, 100 * v / ts))
if len(sys.argv) > 2:
    print('hottest instructions:')
    for s, n, l, a, src in sorted(hot, reverse=True)[:int(sys.argv[2])]:
        print('  %5.2f%% samp  exec %10d  lanes %4.1f  %s  %s' % (100 * s / ts, n, l, a, src[:90]))
