#!/usr/bin/env python
"""ncu --metrics gpu__time_duration.sum --csv log -> per-kernel launch count / total time / share table.
usage: tools/ncu_launch_list.py <launches.csv> <out.txt> <title line> [<command line>]"""
import csv
import re
import sys
from collections import defaultdict

src, out, title = sys.argv[1:4]
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
H = rows[0]
ik, iv, iu, im = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Metric Name")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik]).strip()
    name = re.sub(r"^void ", "void ", name)[:110]
    tot[name] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    cnt[name] += 1
T = sum(tot.values())
lines = ["# " + title] + (["# " + sys.argv[4]] if len(sys.argv) > 4 else []) + ["# cold-cache, serialised: compare SHARES, not absolutes",
                                                                                 f"{'kernel':72s} {'n':>5s} {'total_ms':>10s} {'share':>7s}"]
for k in sorted(tot, key=lambda k: -tot[k]):
    lines.append(f"{k:72s} {cnt[k]:5d} {tot[k]:10.3f} {100 * tot[k] / T:6.2f}%")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
