#!/usr/bin/env python
"""Developer tool: executed warp-instructions of an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) per block of 40 SASS instructions in
address order, with the source lines they come from.  usage: python tools/ncu_regions.py file.csv [block]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
iaddr, isass = 2, 3
iex = hdr.index("Instructions Executed")
ith = hdr.index("Thread Instructions Executed")
cur = None
fname = ""
ins = []
for r in rows[:3] + rows[3:]:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1].replace('tcrt_render', 'r').replace('.cuh', '').replace('.cu', '')
        continue
    if r and r[0] != '':
        cur = f"{fname}:{r[0]}" if r[0].isdigit() else cur
        continue
    if not r:
        continue
    if r[iaddr].startswith('0x'):
        ins.append((int(r[iaddr], 16), cur, r[isass].strip(), int(r[iex]), int(r[ith])))
ins.sort()
tot = sum(i[3] for i in ins)
print(len(ins), "SASS instructions,", tot, "executed warp-instructions")
blk = collections.OrderedDict()
for k, (a, l, s, e, t) in enumerate(ins):
    d = blk.setdefault(k // B, [0, 0, collections.Counter()])
    d[0] += e
    d[1] += t
    d[2][l] += e
cum = 0
for b, (e, t, c) in blk.items():
    cum += e
    top = ', '.join(f"{l}:{100 * v / max(e, 1):.0f}%" for l, v in c.most_common(5))
    print(f"{b * B:5d} {100 * e / tot:5.2f}% cum {100 * cum / tot:5.1f}%  thr/inst {t / max(e, 1):5.1f}  lines {top}")
