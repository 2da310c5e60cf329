#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` : opcode mix + stall samples + hottest lines.
usage: python tools/ncu_sass_summary.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]
si, wi, ei = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
ops, stall, lines = collections.Counter(), collections.Counter(), []
for r in rows[h + 1:]:
    if len(r) <= max(si, wi, ei):
        continue
    toks = r[si].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    try:
        n, s = float(r[ei] or 0), float(r[wi] or 0)
    except ValueError:
        continue
    ops[op.split('.')[0]] += n
    stall[op.split('.')[0]] += s
    lines.append((s, n, r[0], r[si]))
tot, ts = sum(ops.values()) or 1, sum(stall.values()) or 1
print(f'total warp-instructions {tot:.0f}, stall samples {ts:.0f}')
for op, n in ops.most_common(top_n):
    print(f'{op:14s} {n / tot * 100:6.2f}% inst   {stall[op] / ts * 100:6.2f}% stall samples')
print('--- hottest SASS lines by stall samples')
for s, n, addr, src in sorted(lines, reverse=True)[:top_n]:
    print(f'{s / ts * 100:6.2f}%  exec {n:9.0f}  {addr}  {src[:110]}')
