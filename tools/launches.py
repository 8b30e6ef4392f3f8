#!/usr/bin/env python3
"""Reduces an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X) to one line per kernel of this
repository: launches, total and mean duration.  usage: launches.py launches.csv [out.txt]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hd = rows[h]; ki = hd.index("Kernel Name"); vi = hd.index("Metric Value"); ui = hd.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki]
    if "ffv1" not in name and not name.startswith(("k_", "void k_")):
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    short = name.split("(")[0].replace("void ", "").replace("ffv1::", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = ["%-36s %8s %12s %10s %7s" % ("kernel", "launches", "total ms", "mean ms", "share")]
for k, (n, t) in agg.items():
    out.append("%-36s %8d %12.3f %10.3f %6.1f%%" % (k, n, t, t / n, 100 * t / tot if tot else 0))
out.append("%-36s %8s %12.3f" % ("all", "", tot))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
