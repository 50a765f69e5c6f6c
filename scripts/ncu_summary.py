"""Summarise an .ncu-rep (first kernel): key metrics + per-source-line hot spots.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k][0]:>20s} {m[k][1]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(m[h][0]) if m[h][0] else 0
        if v > 0.15:
            print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.2f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
h = None
agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        h = r
        continue
    if h and len(r) == len(h):
        d = dict(zip(h, r))
        if not d["Line No"]:
            continue
        try:
            ie = int(d["Instructions Executed"])
        except ValueError:
            continue
        if ie == 0:
            continue
        key = (cur.split("/")[-1], int(d["Line No"]))
        a = agg.setdefault(key, [0, 0, 0.0, r[1]])
        a[0] += ie
        try:
            a[1] += int(d["# Samples"])
            a[2] = float(d["Avg. Threads Executed"] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
smp = sum(a[1] for a in agg.values())
print(f"total inst (source-attributed) {tot/1e6:.1f}M, samples {smp}")
byfile = {}
for (f, l), a in agg.items():
    b = byfile.setdefault(f, [0, 0])
    b[0] += a[0]
    b[1] += a[1]
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {f:28s} inst {100*b[0]/tot:5.1f}%  samples {100*b[1]/max(smp,1):5.1f}%")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:22s} L{l:4d} inst {100*a[0]/tot:5.1f}% smp {100*a[1]/max(smp,1):5.1f}% | {a[3].strip()[:80]}")
