"""Hot spots of an `ncu --page source --csv` export (SASS view): samples by region and opcode."""
import csv, sys, re
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
totinst = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
print("total samples", tot, "warp-inst executed", totinst)
ops_i, ops_s = Counter(), Counter()
for r in body:
    src = re.sub(r'^\s*(@!?U?P\w+\s+)?', '', r[ix['Source']])
    op = src.split()[0].split('.')[0].rstrip(';') if src.split() else '?'
    ops_i[op] += int(r[ix['Instructions Executed']] or 0)
    ops_s[op] += int(r[ix['# Samples']] or 0)
print("opcode: executed share / stall-sample share")
for op, c in ops_i.most_common(22):
    print(f"  {op:10s} {100*c/totinst:6.2f}%  {100*ops_s[op]/tot:6.2f}%")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
print(f"samples per window of {N} SASS instructions (idx: samples%, executed%)")
for k in range(0, len(body), N):
    w = body[k:k + N]
    s = sum(int(r[ix['# Samples']] or 0) for r in w)
    e = sum(int(r[ix['Instructions Executed']] or 0) for r in w)
    print(f"  {k:5d}: {100*s/tot:6.2f}%  {100*e/totinst:6.2f}%   {w[0][ix['Source']].strip()[:50]}")
