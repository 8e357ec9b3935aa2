#!/usr/bin/env python3
"""Hot-spot view of an ncu source-page CSV (SASS view): executed-instruction histogram by opcode, local-memory
instructions, and the top stall-sample addresses."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = []
for r in rows[2:]:  # a multi-launch export repeats the header rows: keep the first launch
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        ins.append(r)
tot = sum(int(r[iex]) for r in ins)
ops = Counter()
for r in ins:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[iex])
print("total warp inst", tot)
for op, n in ops.most_common(28):
    print(f"  {op:12s} {n:14d} {100*n/tot:5.1f}%")
print("local-memory instructions (executed > 0):")
for k, r in enumerate(ins):
    if ("LDL" in r[isrc] or "STL" in r[isrc]) and int(r[iex]) > tot / 20000:
        print(f"  #{k:6d} {r[isrc][:60]:60s} exec {int(r[iex]):12d} samples {r[isamp]}")
print("top stall-sample instructions:")
for k, r in sorted(enumerate(ins), key=lambda kr: -int(kr[1][isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  #{k:6d} {r[isrc][:70]:70s} exec {int(r[iex]):12d} samples {r[isamp]}")

# stall reasons aggregated by opcode (all samples)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_op = {}
for r in ins:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    d = by_op.setdefault(op, Counter())
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            d[hdr[i]] += v
print("stall samples by opcode (top reasons):")
for op, d in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    tot_s = sum(d.values())
    print(f"  {op:10s} {tot_s:8d}  " + "  ".join(f"{k[6:]}={v}" for k, v in d.most_common(5)))
