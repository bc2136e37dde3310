#!/usr/bin/env python
"""profiles/k_run_traffic.json from an `ncu --set full` capture of one k_run launch of bench.py (read here, no GPU needed):
DRAM bytes of the launch + the sha256 of the engine library the capture was taken with; bench.py reports `roofline.traffic` only when
the library it runs is that one and the launch shape (envs, chunk) matches.
usage: tools/write_traffic_json.py <k_run.ncu-rep> <envs> <chunk> "<capture command>" """
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

rep, envs, chunk, cmd = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[0]
vals = {}
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
    i = H.index(name)
    unit, v = rows[1][i], float(rows[2][i].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[unit]
    vals[name] = v * scale
lib = os.path.join(root, "mujoco_rl_ur5_b200", "csrc", "libgrasp_engine.so")
j = {"lib_sha16": hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16], "envs": envs, "chunk": chunk,
     "dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
     "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
     "launch_ms_under_ncu": vals["gpu__time_duration.sum"], "capture": cmd}
json.dump(j, open(os.path.join(root, "profiles", "k_run_traffic.json"), "w"), indent=1)
print(json.dumps(j))
