#!/usr/bin/env python3
"""Enumerate staging geometries (hoff, row, study) of the lane-quad kernels and their shared-memory bank behaviour; prints the
best candidate per M = ceil((P + 1) / 2).  The table in csrc/plan.h (mcp_stage_geom) is this script's output.

Bank model: 32 banks x 4 B; a 16-byte access is served per quarter-warp, an 8-byte access per half-warp; lanes that read the same
address share one access.  Lane q of quad i reads the low half (h = 0) or the reversed high half (h = 1) of staged row r:
p-half h = (q in {1, 2}), r-half h = (q >= 2); lane q writes row q."""


def nway(accesses):
    cnt = {}
    for a, w in set(accesses):
        for k in range(w):
            for b in (0, 1):
                bank = (2 * (a + k) + b) % 32
                cnt[bank] = cnt.get(bank, 0) + 1
    return max(cnt.values())


def evaluate(study, hoff, row, m):
    mh = (m + 1) & ~1

    def half(q, which):
        return (1 if q in (1, 2) else 0) if which == "p" else (1 if q >= 2 else 0)

    l64 = l128 = st = 1
    for which in "pr":
        for r in range(4):
            for e in range(0, mh, 2):
                if e + 1 >= m:  # the compiler narrows the last 16-byte load of an odd half to 8 bytes
                    for hw in (0, 1):
                        l64 = max(l64, nway([(i * study + r * row + half(q, which) * hoff + e, 1)
                                             for i in range(4 * hw, 4 * hw + 4) for q in range(4)]))
                else:
                    for qw in range(4):
                        l128 = max(l128, nway([(i * study + r * row + half(q, which) * hoff + e, 2)
                                               for i in range(2 * qw, 2 * qw + 2) for q in range(4)]))
    for qw in range(4):
        for base in (0, hoff):
            for e in range(0, mh, 2):
                st = max(st, nway([(i * study + q * row + base + e, 2) for i in range(2 * qw, 2 * qw + 2) for q in range(4)]))
    return l64, l128, st


if __name__ == "__main__":
    for m in range(1, 10):
        mh = (m + 1) & ~1
        res = []
        for hoff in (mh, mh + 2, mh + 4):
            for row in range(2 * hoff, 2 * hoff + 9, 2):
                for extra in range(0, 16, 2):
                    ev = evaluate(4 * row + extra, hoff, row, m)
                    res.append((max(ev), sum(ev), 4 * row + extra, hoff, row, ev))
        res.sort()
        old = evaluate(8 * mh + 2, mh, 2 * mh, m)
        _, _, study, hoff, row, ev = res[0]
        print(f"M={m}: (hoff, row, study) = ({hoff}, {row}, {study}) ways (ld8, ld16, st16) = {ev}; former rule {old}")
