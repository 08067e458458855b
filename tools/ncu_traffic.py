"""profiles/r02_traffic.json from an `ncu --set full` report: DRAM bytes per launch of every kernel in it.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep "<the command that was profiled>" [out.json]

bench.py reads the file for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch of the
dominant kernel) — nothing about traffic is hard-coded there."""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    u = unit.strip().lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_us(v, unit):
    x = float(v.replace(",", ""))
    return x * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit.strip().lower(), 1)


kern = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
    k = kern.setdefault(name, {"launches": 0, "read": 0.0, "write": 0.0, "us": 0.0})
    k["launches"] += 1; k["read"] += rd; k["write"] += wr; k["us"] += us
res = {"source": "ncu --set full --clock-control none, report %s; command: %s" % (os.path.basename(rep), cmd), "kernels": {}}
for name, k in kern.items():
    n = k["launches"]
    res["kernels"][name] = {"launches": n, "dram_read_bytes_per_launch": k["read"] / n, "dram_write_bytes_per_launch": k["write"] / n,
                            "dram_bytes_per_launch": (k["read"] + k["write"]) / n, "avg_us_under_ncu": k["us"] / n}
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps(res, indent=1))
