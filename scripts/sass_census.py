#!/usr/bin/env python3
"""Per-kernel census of the SASS mnemonics that prove the Blackwell data path (B200_PROFILING.md): UTMALDG / UTMAPF (TMA tensor
loads / L2 prefetch), SYNCS (mbarrier), and the fp64 / shared-memory instruction counts, from libmg_b200.so:
    python scripts/sass_census.py > profiles/r2_sass_tma.txt"""
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pde_multigrid_b200", "libmg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, rows = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        rows[cur] = dict.fromkeys(("UTMALDG", "UTMAPF", "SYNCS", "LDS", "STS", "DADD", "DFMA", "DMUL", "LDG", "STG", "BAR"), 0)
        rows[cur]["total"] = 0
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        continue
    rows[cur]["total"] += 1
    for k in rows[cur]:
        if k != "total" and re.search(r"\b%s\b" % k, line.split("/*")[1] if False else line):
            rows[cur][k] += 1
print("arch of the cubins:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
print("%-62s %7s %7s %6s %6s %5s %5s %5s %5s %5s %5s %4s %6s" % ("kernel", "UTMALDG", "UTMAPF", "SYNCS", "LDS", "STS", "DADD", "DFMA", "DMUL", "LDG", "STG", "BAR", "total"))
for k in sorted(rows, key=lambda k: (-rows[k]["UTMALDG"], k)):
    r = rows[k]
    if r["UTMALDG"] or "relax" in k or "residual" in k or "interp" in k:
        print("%-62s %7d %7d %6d %6d %5d %5d %5d %5d %5d %5d %4d %6d" % (k[:62], r["UTMALDG"], r["UTMAPF"], r["SYNCS"], r["LDS"], r["STS"], r["DADD"],
                                                                        r["DFMA"], r["DMUL"], r["LDG"], r["STG"], r["BAR"], r["total"]))
print("TOTAL UTMALDG %d, SYNCS %d over %d kernels" % (sum(r["UTMALDG"] for r in rows.values()), sum(r["SYNCS"] for r in rows.values()), len(rows)))
