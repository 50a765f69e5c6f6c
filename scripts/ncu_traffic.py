"""DRAM traffic of the captured sweep kernels -> profiles/r02_sweep_traffic.json
    python scripts/ncu_traffic.py growth_1M=a.ncu-rep,b.ncu-rep relu_1M=c.ncu-rep ...
Several captures per workload are summed (the sweep of points with extra lanes
is two kernels, list_cubes + interact_lists). bench.py reports the number as
roofline.traffic (bytes per sweep)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r02_sweep_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for arg in sys.argv[1:]:
    workload, reps = arg.split("=")
    kernels = []
    for rep in reps.split(","):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        m = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}

        def num(key):
            value, unit = m[key]
            return float(value.replace(",", "")) * SCALE.get(unit, 1)

        kernels.append({
            "kernel": m["Kernel Name"][0].split("(")[0],
            "dram_bytes_read": num("dram__bytes_read.sum"),
            "dram_bytes_write": num("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(
                m["gpu__time_duration.sum"][0].replace(",", "")),
            "source": os.path.basename(rep) + " (ncu --set full --clock-control "
                      "none, one launch after warm-up)"})
    table[workload] = {
        "dram_bytes_per_launch": sum(k["dram_bytes_read"] + k["dram_bytes_write"]
                                     for k in kernels),
        "kernels": kernels}
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(table, indent=1))
