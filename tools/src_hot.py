#!/usr/bin/env python3
"""Per-instruction view of an ncu source-page CSV: address, SASS, samples, executed, top stall reasons.
usage: src_hot.py source.csv [first_row] [last_row]   (rows are counted over SASS instructions)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
out = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[idx["# Samples"]].isdigit():
        continue
    top = sorted(((int(r[idx[n]]) if r[idx[n]].isdigit() else 0, n) for n in stalls), reverse=True)[:2]
    out.append((r[idx["Address"]][-5:], r[idx["Source"]].strip()[:64], int(r[idx["# Samples"]]), r[idx["Instructions Executed"]],
                r[idx["Avg. Threads Executed"]], [(n[6:], v) for v, n in top if v]))
a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = int(sys.argv[3]) if len(sys.argv) > 3 else len(out)
tot = sum(o[2] for o in out)
print("instructions %d, samples %d" % (len(out), tot))
for i, o in enumerate(out[a:b]):
    print("%4d %s %-64s %7d %11s %3s %s" % (a + i, o[0], o[1], o[2], o[3], o[4], o[5]))
