#!/usr/bin/env python
"""Dynamic instruction mix of the first kernel in an .ncu-rep (source page): warp instructions per thread."""
import collections, csv, subprocess, sys
rep, cells, cpt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
by, tot = collections.Counter(), 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        e = int(r[iE])
    except Exception:
        continue
    op = r[iS].split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    by[o] += e
    tot += e
threads = cells / cpt / 32
print(rows[0][1], "total warp instr", tot, "per thread", round(tot / threads, 1))
for o, c in by.most_common(28):
    print(f"{o:10s} {c / threads:8.2f}")
