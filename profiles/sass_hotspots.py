#!/usr/bin/env python
"""Opcode mix and top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv > file.csv`:
    python profiles/sass_hotspots.py file.csv [n_top]
"""
import csv
import sys
from collections import Counter


def main(path, ntop=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_barrier", "stall_wait", "stall_math", "stall_short_sb",
                                      "stall_mio", "stall_not_selected", "stall_selected", "stall_branch_resolving")}
    data = [r for r in rows[2:] if len(r) > max(cols.values()) and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in data)
    totex = sum(int(r[iex]) for r in data)
    print("samples", tot, "warp-instructions", totex, "sass lines", len(data))
    print("stall totals:", {k: f"{sum(int(r[i]) for r in data) / tot * 100:.1f}%" for k, i in cols.items()})
    op = Counter()
    for r in data:
        s = r[ia].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op[s.split()[0].split(".")[0]] += int(r[iex])
    print("opcode mix:", ", ".join(f"{k} {v / totex * 100:.1f}%" for k, v in op.most_common(24)))
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
        print(f"{int(r[isamp]) / tot * 100:5.2f}%  lsb={r[cols['stall_long_sb']]:>6s} bar={r[cols['stall_barrier']]:>6s} "
              f"wait={r[cols['stall_wait']]:>6s} math={r[cols['stall_math']]:>6s} ex={r[iex]:>9s}  {r[ia].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
