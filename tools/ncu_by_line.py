#!/usr/bin/env python3
"""Developer tool: executed instructions / stall samples of one kernel by SOURCE line.
usage: ncu_by_line.py <ncu --page source --csv export> <all.sass from `nvdisasm --print-line-info-inline cubin`> <mangled kernel name> [anchor file]
Joins the ncu SASS rows (in address order) with nvdisasm's line-info chains (same order) and aggregates on the frame of
`anchor file` (default kernels.cuh) nearest to the kernel body, else the outermost frame."""
import collections
import csv
import re
import sys

csv_path, sass_path, target = sys.argv[1:4]
anchor = sys.argv[4] if len(sys.argv) > 4 else "kernels.cuh"
ins, on, block, chain = [], False, [], []
for l in open(sass_path):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        on = m.group(1) == target
        continue
    if not on:
        continue
    if l.startswith("\t//## File"):
        m = re.match(r'\t//## File "([^"]+)", line (\d+)', l)
        block.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if block:
            chain, block = block, []
        ins.append((m.group(2).strip(), chain))
rows = list(csv.reader(open(csv_path)))
demangled_hint = re.sub(r"^_ZN3mcp\d+", "", target)[:12]
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
body = None
for s in starts:
    hdr = rows[s + 1]
    b = []
    for r in rows[s + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            b.append(r)
    if len(b) == len(ins):
        body = b
        break
assert body is not None, f"no kernel in the export has {len(ins)} instructions"
iex, ith, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
K, KS, KT = collections.Counter(), collections.Counter(), collections.Counter()
for (txt, ch), r in zip(ins, body):
    key = ch[-1] if ch else ("?", 0)
    for f, ln in reversed(ch):
        if f == anchor:
            key = (f, ln)
            break
    K[key] += int(r[iex]); KS[key] += int(r[isamp]); KT[key] += int(r[ith])
tot, tots = sum(K.values()), sum(KS.values())
print("total warp inst", tot, "samples", tots)
for key, e in K.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 50):
    print(f"{key[0]}:{key[1]:5d}  inst {100 * e / tot:5.1f}%  samples {100 * KS[key] / tots:5.1f}%  lanes {KT[key] / max(e, 1):5.1f}")
