#!/usr/bin/env python3
"""Regenerate profiles/inst_per_fit.json from an ncu raw-page CSV of the headline capture (ols_lean_kernel + ols_solve_kernel
of ONE bench step of `studies` studies): executed warp / thread instructions per fit.  The algorithmic and essential
 counts are kept.  usage: make_inst_per_fit.py <raw.csv> <studies per step> <source note>"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rows = list(csv.reader(open(sys.argv[1])))
studies = int(sys.argv[2])
hdr = rows[0]
ik, iw, it = hdr.index("Kernel Name"), hdr.index("smsp__inst_executed.sum"), hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")
seen, warp, thread, names = set(), 0.0, 0.0, []
for r in rows[2:]:
    name = r[ik].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    names.append(name)
    warp += float(r[iw])
    thread += float(r[iw]) * float(r[it])  # warp instructions x average active threads
path = ROOT / "profiles" / "inst_per_fit.json"
prof = json.loads(path.read_text())
prof["warp_inst_per_fit"] = warp / studies
prof["thread_inst_per_fit"] = thread / studies
prof["source"] = sys.argv[3] + " (ncu smsp__inst_executed.sum of " + " + ".join(names) + f" / {studies} studies, n=500); algorithmic figure: SURVEY.md section 8(d); essential figure: DESIGN.md section 3.3"
path.write_text(json.dumps(prof, indent=1) + "\n")
print(json.dumps(prof, indent=1))
