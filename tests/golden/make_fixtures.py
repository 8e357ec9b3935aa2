#!/usr/bin/env python3
"""Extract the reference's frozen validation datasets into portable fixtures.

Run HERE (the container that has /root/reference) — the GPU box never sees the
reference tree, so the extracted vectors are committed next to this script.

Sources (all under /root/reference):
  validation/data/<case>.rds         sim-0 design matrix + outcome of the shipped
                                     engine at base_seed = splitmix64(spec.seed)
                                     (engine-orchestrator/src/debug.rs:343-349,
                                      engine-core/src/introspect.rs:104-155)
  validation/data/<case>.golden.rds  frozen R lm/glm/lmer fits on those bytes
                                     (validation/common.R:534-543)

Output: tests/golden/validation_cases.json.gz (one entry per case: spec, columns,
design as f32 bit patterns column-major, outcome as f32 bit patterns, cluster
ids, golden beta/se/stat/crit as f64 hex).  Every design/outcome value is
f32-representable (asserted), so u32 bit patterns lose nothing.

The RDS reader below handles just the R serialisation subset those files use
(XDR format v2/v3: NILVALUE, pairlist/attributes, CHARSXP, LGL/INT/REAL/STR/VEC).
"""
import gzip
import json
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/validation/data")
OUT = Path(__file__).resolve().parent / "validation_cases.json.gz"


class RdsReader:
    def __init__(self, data: bytes):
        self.b = data
        self.p = 0
        self.refs = []

    def _take(self, n):
        s = self.b[self.p:self.p + n]
        self.p += n
        return s

    def i32(self):
        return struct.unpack(">i", self._take(4))[0]

    def f64(self):
        return struct.unpack(">d", self._take(8))[0]

    def header(self):
        assert self._take(2) == b"X\n", "not XDR rds"
        version = self.i32()
        self.i32()
        self.i32()
        if version >= 3:
            n = self.i32()
            self._take(n)

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if ty == 254 or ty == 253 or ty == 242:  # NILVALUE / GLOBALENV / EMPTYENV
            return None
        if ty == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if ty == 2:  # LISTSXP (pairlist): attr?, tag?, car, cdr
            out = {}
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out[tag if tag is not None else len(out)] = car
                flags = self.i32()
                ty2 = flags & 0xFF
                if ty2 == 254:
                    return out
                assert ty2 == 2, f"pairlist cdr type {ty2}"
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
        if ty == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self._take(n).decode("utf-8")
        val = None
        if ty == 10 or ty == 13:  # LGLSXP / INTSXP
            n = self.i32()
            val = list(struct.unpack(f">{n}i", self._take(4 * n)))
            if ty == 10:
                val = [None if v == -2147483648 else bool(v) for v in val]
        elif ty == 14:  # REALSXP
            n = self.i32()
            val = np.frombuffer(self._take(8 * n), dtype=">f8").astype(np.float64)
        elif ty == 16:  # STRSXP
            n = self.i32()
            val = [self.item() for _ in range(n)]
        elif ty == 19:  # VECSXP
            n = self.i32()
            val = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"SEXP type {ty} at {self.p}")
        if has_attr:
            attrs = self.item()
            if ty == 19 and attrs and "names" in attrs:
                return dict(zip(attrs["names"], val))
            if attrs and "dim" in attrs and ty == 14:
                return {"_dim": attrs["dim"], "_data": val}
            if attrs and "names" in attrs and ty in (14, 13, 10, 16):
                return {"_names": attrs["names"], "_data": val}
        return val


def read_rds(path):
    r = RdsReader(gzip.open(path, "rb").read())
    r.header()
    return r.item()


def scalar(x):
    if x is None:
        return None
    if isinstance(x, dict) and "_data" in x:
        x = x["_data"]
    if isinstance(x, np.ndarray):
        return float(x[0]) if x.size == 1 else [float(v) for v in x]
    if isinstance(x, list):
        return x[0] if len(x) == 1 else x
    return x


def plain(x):
    """JSON-friendly deep conversion (used for the free-form `spec` block)."""
    if isinstance(x, dict):
        if "_data" in x and "_names" in x:
            return dict(zip(x["_names"], plain(x["_data"])))
        if "_data" in x and "_dim" in x:
            return {"dim": list(x["_dim"]), "data": [float(v) for v in x["_data"]]}
        return {str(k): plain(v) for k, v in x.items()}
    if isinstance(x, np.ndarray):
        return [float(v) for v in x]
    if isinstance(x, list):
        return [plain(v) for v in x]
    return x


def f32_bits(a):
    a = np.asarray(a, dtype=np.float64)
    f = a.astype(np.float32)
    assert np.array_equal(f.astype(np.float64), a), "fixture value not f32-representable"
    return [int(v) for v in f.view(np.uint32)]


def f64_hex(a):
    if isinstance(a, dict):  # a named numeric vector
        a = a["_data"]
    return [float(v).hex() for v in np.atleast_1d(np.asarray(a, dtype=np.float64))]


def main():
    cases = {}
    families = ("ols_", "glm_", "probit_", "poisson_", "lme_", "lmm_")
    for path in sorted(REF.glob("*.rds")):
        name = path.name[:-4]
        if name.endswith(".golden") or not name.startswith(families):
            continue
        if name.endswith("_extra_slopes"):
            continue  # slopes on extra groupings: outside the restated path
        d = read_rds(path)
        gpath = REF / f"{name}.golden.rds"
        g = read_rds(gpath) if gpath.exists() else None
        design = d["design"]
        n, p = (int(v) for v in design["_dim"])
        entry = {
            "label": scalar(d["label"]),
            "spec": plain(d["spec"]),
            "columns": d["columns"],
            "n": n,
            "p": p,
            "design_f32_colmajor": f32_bits(design["_data"]),
            "outcome_f32": f32_bits(d["outcome"]),
            "cluster_ids": d.get("cluster_ids"),
            "extra_grouping_ids": d.get("extra_grouping_ids"),
            "mcpower_version": scalar(d.get("mcpower")),
        }
        if g is not None and "z" in g:  # lmer fits with variance components (validation/common.R: lmm goldens)
            entry["golden"] = {"beta": f64_hex(g["beta"]), "z": f64_hex(g["z"]), "vars": f64_hex(g["vars"]),
                               "corr": f64_hex(g["corr"]), "sigma2": f64_hex(g["sigma2"])}
        elif g is not None:
            gb = g["beta"]
            entry["golden"] = {
                "beta": f64_hex(gb["_data"] if isinstance(gb, dict) else gb),
                "se": f64_hex(g["se"]["_data"] if isinstance(g["se"], dict) else g["se"]),
                "stat": f64_hex(g["stat"]["_data"] if isinstance(g["stat"], dict) else g["stat"]),
                "crit": f64_hex(g["crit"]["_data"] if isinstance(g["crit"], dict) else g["crit"]),
            }
        cases[name] = entry
    with gzip.open(OUT, "wt", compresslevel=9) as f:
        f.write(json.dumps(cases, separators=(",", ":")))
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB, {len(cases)} cases)")
    for k, v in cases.items():
        print(" ", k, v["n"], v["p"], v["columns"], json.dumps(v["spec"])[:200])


if __name__ == "__main__":
    sys.exit(main())
