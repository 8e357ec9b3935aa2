#!/bin/bash
# developer tool: sweep lean-kernel knobs.  usage: tune_lean.sh "<stagger list>" "<tune list>"  (tune needs -DMCP_LEAN_TUNE)
for t in ${2:-default}; do for s in ${1:-default}; do
  echo -n "tune=$t stagger=$s "
  ( [ "$t" != default ] && export MCPOWER_B200_LEAN_TUNE=$t; [ "$s" != default ] && export MCPOWER_B200_LEAN_STAGGER=$s;
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sims-per-step 2097152 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['power_x1_x2'])" )
done; done
