#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed-instruction histogram per opcode, stall
reasons, hottest instructions.  usage: ncu -i X.ncu-rep --page source --csv > s.csv; python tools/ncu_sass_summary.py s.csv"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
thr = collections.Counter()
stalls = collections.Counter()
total = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["Instructions Executed"]])
        t = int(r[col["Thread Instructions Executed"]])
    except ValueError:
        continue
    src = r[col["Source"]].strip()
    parts = src.split()
    op = parts[0]
    if op.startswith("@"):
        op = parts[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("MUFU", "LDS", "LDG", "STL", "LDL")) else "")
    ops[op] += n
    thr[op] += t
    total += n
    samples = int(r[col["# Samples"]] or 0)
    lines.append((samples, n, t, src))
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                stalls[h] += int(r[col[h]] or 0)
            except ValueError:
                pass
print(f"total warp-instructions {total}")
print("opcode            warp-inst    share   avg-threads")
for op, n in ops.most_common(40):
    print(f"{op:16s} {n:11d}  {100.0 * n / total:6.2f}%   {thr[op] / max(1, n):5.1f}")
ts = sum(stalls.values())
print("\nstall samples:")
for k, v in stalls.most_common(12):
    print(f"  {k:28s} {100.0 * v / max(1, ts):6.2f}%")
print("\nhottest instructions by samples:")
for s, n, t, src in sorted(lines, reverse=True)[:25]:
    print(f"  {s:6d} {n:10d}  {src[:90]}")
