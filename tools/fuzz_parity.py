#!/usr/bin/env python3
"""Developer tool: the randomised differential parity of tests/test_gpu_fuzz_parity.py over an arbitrary case range.
usage: fuzz_parity.py <first case> <last case>   (needs a GPU; prints the failing labels)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import test_gpu_fuzz_parity as F  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for case in range(lo, hi):
    try:
        F.test_gpu_equals_oracle_on_a_random_spec(case)
    except AssertionError as e:
        bad += 1
        print("FAIL", case, str(e)[:300], flush=True)
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("ERROR", case, type(e).__name__, str(e)[:300], flush=True)
    if (case - lo + 1) % 25 == 0:
        print(f"... {case - lo + 1} cases done, {bad} failures", flush=True)  # a run cut off by a time limit still says how far it got
print(f"cases {lo}..{hi - 1}: {bad} failures")
