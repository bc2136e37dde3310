#!/usr/bin/env python
"""Aggregate an ncu source-page capture of k_run by CUDA source line (needs -lineinfo in the build).

usage: tools/ncu_by_line.py <report.ncu-rep> <libgrasp_engine.so> [file-substring] [top-n]
Joins per-SASS-address counters from `ncu --page source --csv` with the line table printed by `nvdisasm -gi`.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    txt = ""
    for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):  # one cubin per translation unit (engine variants)
        t = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout
        if any(l.startswith(".text.") and kernel in l for l in t.splitlines()):
            txt = t
            break
    table, on, cur = {}, False, ("?", 0)
    for line in txt.splitlines():
        if line.startswith(".text."):
            on = kernel in line
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, so = sys.argv[1], sys.argv[2]
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(so, os.environ.get("NCU_KERNEL", "k_run"))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hi]
    ia, ie, isamp = H.index("Address"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
    base = None
    agg = collections.defaultdict(lambda: [0, 0])
    for r in rows[hi + 1:]:
        if len(r) < len(H):
            continue
        try:
            a = int(r[ia], 16)
        except ValueError:
            continue
        if base is None:
            base = a
        key = table.get(a - base, ("?", 0))
        agg[key][0] += int(float(r[ie] or 0))
        agg[key][1] += int(float(r[isamp] or 0))
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    items = [(k, v) for k, v in agg.items() if filt in k[0]]
    print(f"{'file:line':34s} {'warp-inst':>12s} {'%inst':>6s} {'samples':>9s} {'%time':>6s}")
    for k, v in sorted(items, key=lambda x: -x[1][1])[:topn]:
        print(f"{k[0] + ':' + str(k[1]):34s} {v[0]:12d} {100 * v[0] / ti:6.2f} {v[1]:9d} {100 * v[1] / ts:6.2f}")


if __name__ == "__main__":
    main()
