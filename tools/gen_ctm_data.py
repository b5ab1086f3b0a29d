#!/usr/bin/env python3
"""Build-time generator: continuum coefficient DATA -> build/jrb_ctm_data.h (git-ignored).

The CO2/H2O/N2/O2 continuum coefficients are data arrays embedded in the reference's GPL sources
(src/ctmco2.tbl, ctmh2o.tbl, ctmn2.tbl, ctmo2.tbl; `#include`d at src/jr_common.h:317,335,366,380).
They are input data of the hot path, so they are NOT copied into this repository: this script reads
them from the reference checkout (env JURASSIC_REF, default /root/reference) at build time and
writes a plain header of `static const double` arrays into build/.  On a machine without the
reference checkout the previously built libraries are used as they are.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("JURASSIC_REF", "/root/reference")
OUT = os.path.join(ROOT, "build", "jrb_ctm_data.h")

FILES = {
    "ctmco2.tbl": [("co2296", 2001), ("co2260", 2001), ("co2230", 2001)],
    "ctmh2o.tbl": [("h2o296", 2001), ("h2o260", 2001), ("h2ofrn", 2001)],
    "ctmn2.tbl": [("ba", 98), ("betaa", 98)],
    "ctmo2.tbl": [("ba", 90), ("betaa", 90)],
}
PREFIX = {"ctmco2.tbl": "", "ctmh2o.tbl": "", "ctmn2.tbl": "n2_", "ctmo2.tbl": "o2_"}


def parse(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"static\s+double\s+const\s+\((\w+)\)\[(\d+)\]\s*=\s*\{([^}]*)\}", txt):
        name, n, body = m.group(1), int(m.group(2)), m.group(3)
        vals = [v.strip() for v in body.replace("\n", " ").split(",") if v.strip()]
        if len(vals) != n:
            raise SystemExit(f"{path}: array {name} has {len(vals)} values, expected {n}")
        out[name] = vals
    return out


def main():
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        if os.path.exists(OUT):
            print(f"gen_ctm_data: {src} absent, keeping existing {OUT}")
            return 0
        print(f"gen_ctm_data: reference sources not found at {src}", file=sys.stderr)
        return 1
    lines = ["/* GENERATED at build time by tools/gen_ctm_data.py from the reference's src/ctm*.tbl -- do not commit. */",
             "#ifndef JRB_CTM_DATA_H", "#define JRB_CTM_DATA_H"]
    for fn, arrays in FILES.items():
        got = parse(os.path.join(src, fn))
        for name, n in arrays:
            if name not in got or len(got[name]) != n:
                raise SystemExit(f"{fn}: array {name}[{n}] not found")
            lines.append(f"static const double jrb_{PREFIX[fn]}{name}[{n}] = {{")
            vals = got[name]
            for i in range(0, n, 8):
                lines.append("  " + ", ".join(vals[i:i + 8]) + ",")
            lines.append("};")
    lines.append("#endif")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    new = "\n".join(lines) + "\n"
    if not os.path.exists(OUT) or open(OUT).read() != new:
        open(OUT, "w").write(new)
    print(f"gen_ctm_data: wrote {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
