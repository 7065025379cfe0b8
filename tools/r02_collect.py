#!/usr/bin/env python
"""tools/r02_collect.py TAG -- after tools/r02_final.sh TAG ran under gpurun: copy the records into profiles/ (bench lines,
launch list, parity logs), summarise every ncu capture (tools/ncu_summary.py), rebuild profiles/ncu_traffic.json (DRAM bytes
per launch, keyed by workload, with the kernel name bench.py reports for it) and dump the SASS histogram of the shipped
k_pool / k_check / k_feas cubins."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02z"
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return int(round(float(v) * scale))


traffic = {}
for w in ("c3", "c5", "c4", "c2", "c1", "feas"):
    rep = os.path.join(O, f"{TAG}_prof_{w}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{TAG}_ncu_{w}.md"), "w").write(
        f"# ncu --set full --clock-control none, one launch of the check kernel of `bench.py --workload {w}` "
        f"(tools/r02_final.sh {TAG}; k_feas: tools/feas_bench.py 22)\n\n" + md)
    rows, units = raw(rep)
    d = rows[0]
    rd = to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
    wr = to_bytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
    bench = os.path.join(O, f"{TAG}_bench_{w}.json")
    kernel = json.load(open(bench))["roofline"]["kernel"] if os.path.exists(bench) else None
    if w != "feas":
        traffic[w] = {"kernel": kernel, "ncu_kernel_name": d["Kernel Name"], "dram_bytes_read": rd, "dram_bytes_write": wr,
                      "dram_bytes_per_launch": rd + wr,
                      "source": f"profiles/{TAG}_ncu_{w}.md (ncu --set full --clock-control none, one launch: dram__bytes_read.sum + dram__bytes_write.sum)"}
    print(w, d["Kernel Name"][:60], d["gpu__time_duration.sum"], units["gpu__time_duration.sum"], rd + wr)
if traffic:
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for f in sorted(os.listdir(O)):
    if f.startswith(TAG + "_") and (f.endswith(".json") or f.endswith(".csv") or f.endswith(".log")) and "ncu_full" not in f and "ncu_list" not in f:
        if os.path.getsize(os.path.join(O, f)) > 0:
            shutil.copy(os.path.join(O, f), os.path.join(P, f))
lib = os.path.join(ROOT, "rapidquadcoptercollisiondetection_b200", "librqcd.so")
with open(os.path.join(P, f"{TAG}_sass_kernels.txt"), "w") as fh:
    for sub in ("k_pool", "k_check", "k_feas"):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_histogram.py"), lib, sub], capture_output=True, text=True).stdout
        keep, on = [], True
        for block in txt.split("\n## "):
            if "rqcd_fma::" in block.split("\n")[0]:
                continue  # the RQCD_F_FAST_FMA build of the same kernels
            keep.append(block)
        fh.write("\n## ".join(keep) + "\n")
print("done")
