"""Benchmark of the ya||a hot path: Heun-step cell-updates/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload growth_1M|relu_1M|epithelium_1M|...]

One "step" is one model step of BASELINE.json's configs[1] (examples/
passive_growth.cu at 1 M cells): Solution<Po_cell, Grid_solver>::take_step
<relu_w_epithelium>(dt, reset_nbs) -- grid build, 27-cube pairwise sweep, Heun
update, twice -- followed by the proliferate kernel (cell division, dynamic n).
Everything goes through the C ABI of include/yalla_b200.h.

  value     sum over steps of n_t / device time (CUDA events on the launching
            stream, state resident in HBM), whole job over all ranks
  e2e       same metric through yb_sim_step_host: host buffers in and out, the
            H2D and D2H copies inside the timed region (host clock + sync)
  roofline  the pairwise sweep kernel, timed live with CUDA events (product arm)
  cpu_baseline  the CPU oracle on a bounded sample of the same workload
  --impl reference  the UNMODIFIED reference headers compiled for sm_100a
            (oracle/_ref/libyalla_ref.so) on the same workload: ya||a has no CPU
            path, its own CUDA build is the baseline (BASELINE.json north_star).
            Falls back to the CPU oracle port if that library is absent.

With N > 1 (torchrun, one rank per GPU) every rank integrates its own tissue of
the same size -- the models shard by independent tissues, there is no data-path
collective -- and the ranks are bracketed by barriers; time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import yalla_b200 as yb  # noqa: E402
from yalla_b200 import workloads  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libyalla_oracle.so")

# name -> model, cells, nearest-neighbour distance, dt, extra
WORKLOADS = {
    # configs[1]: passive_growth at 1 M cells, n_max 2 M, division every step
    "growth_1M": dict(model="growth", n=1_000_000, n_max=2_097_152, d=0.75,
                      dt=0.2, params={"prolif_rate": 0.006, "mean_dist": 0.75,
                                      "seed": 2}, typed=True),
    # configs[2]: epithelium at 1 M cells (bending forces, no friction)
    "epithelium_1M": dict(model="epithelium", n=1_000_000, n_max=1_000_000,
                          d=0.8, dt=0.05, params={}, typed=False),
    # float3 relu_force tissue (the relaxation functor; configs[4]'s kind)
    "relu_1M": dict(model="relu_grid", n=1_000_000, n_max=1_000_000, d=0.8,
                    dt=0.1, params={}, typed=False),
    "relu_10M": dict(model="relu_grid", n=10_000_000, n_max=10_000_000, d=0.8,
                     dt=0.1, params={}, typed=False),
    "growth_100k": dict(model="growth", n=100_000, n_max=262_144, d=0.75,
                        dt=0.2, params={"prolif_rate": 0.006, "mean_dist": 0.75,
                                        "seed": 2}, typed=True),
}
LANES_BYTES = {3: 12, 5: 20, 7: 28}


def make_state(spec, seed):
    rng = np.random.default_rng(seed)
    lanes = yb.MODEL_LANES[spec["model"]]
    if lanes == 3:
        X = workloads.lattice_ball(spec["n"], spec["d"], rng)
    else:
        X = workloads.polarized_ball(spec["n"], spec["d"], rng, lattice=True,
                                     noise=0.0 if spec["typed"] else 0.5)
    types = None
    if spec["typed"]:
        types = workloads.shell_types(X)
        X[types == 0, 3:5] = 0  # mesenchyme carries no polarity
    growth = spec["n_max"] / spec["n"]
    gs = workloads.grid_size_for(spec["n"], spec["d"], growth=growth)
    return X, types, gs


def new_sim(lib, spec, X, types, gs):
    sim = lib.sim(spec["model"], spec["n_max"], gs, 1.0)
    for key, value in spec["params"].items():
        sim.set_param(key, value)
    if types is not None:
        sim.set_ints("type", types)
    sim.set_state(X)
    return sim


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)],
                stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        self.file.seek(0)
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.file:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(sm_max)) if sm_max else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(spec, steps=2, sample_cells=200_000):
    """The oracle (a CPU port of the reference's algorithm) on a bounded sample
    of the workload: the same kind of tissue at sample_cells cells, division
    switched off (no curand on the host)."""
    lib = yb.load(ORACLE_LIB)
    small = dict(spec, n=min(spec["n"], sample_cells))
    small["n_max"] = small["n"]
    small["params"] = dict(spec["params"])
    if "prolif_rate" in small["params"]:
        small["params"]["prolif_rate"] = 0.0
    X, types, gs = make_state(small, seed=99)
    with new_sim(lib, small, X, types, gs) as sim:
        ms, updates = sim.step_timed(spec["dt"], steps)
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    sample = (f"{small['model']} model, {small['n']} cells, {steps} steps, "
              f"division off; OpenMP over cells")
    return {"value": updates / (ms * 1e-3), "unit": "cell-updates/s",
            "cores": cores, "kind": "port", "sample": sample}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=30)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", default="product",
                        choices=["product", "reference"])
    parser.add_argument("--workload", default="growth_1M", choices=sorted(WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    warmup = max(args.warmup, 3)
    spec = WORKLOADS[args.workload]

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    is_reference = args.impl == "reference"

    if is_reference and rank != 0:
        return  # the reference arm runs on rank 0 alone

    use_dist = world > 1 and not is_reference
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank if not is_reference else 0)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    fallback_to_port = False
    if is_reference:
        if os.path.exists(yb.REFERENCE_LIB):
            lib = yb.reference()
        else:
            lib = yb.load(ORACLE_LIB)  # no reference build here: the CPU port
            fallback_to_port = True
            spec = dict(spec, params=dict(spec["params"], prolif_rate=0.0))
    else:
        lib = yb.product()

    X, types, gs = make_state(spec, seed=1000 + rank)
    lanes = X.shape[1]
    steps = args.steps if not fallback_to_port else min(args.steps, 2)
    sim = new_sim(lib, spec, X, types, gs)
    sim.step(spec["dt"], warmup if not fallback_to_port else 0)
    sim.sync()

    # ---- device-resident throughput ---------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms, updates = sim.step_timed(spec["dt"], steps)
    barrier()
    clocks = sampler.stop()
    n_end = sim.n()
    if use_dist:
        stats = torch.tensor([ms, float(updates)], dtype=torch.float64, device="cuda")
        worst = stats.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        ms, updates = float(worst[0]), int(stats[1])
    value = updates / (ms * 1e-3)

    # ---- end to end: host buffers through the C ABI ---------------------------
    e2e = None
    if not is_reference:
        host_in = torch.from_numpy(X).pin_memory().numpy() if X.size else X
        host_out = np.zeros((spec["n_max"], lanes), dtype=np.float32)
        e2e_steps = max(3, steps // 3)
        n_in = len(host_in)
        sim.close()
        sim = new_sim(lib, spec, X, types, gs)
        sim.step_host(host_in, spec["dt"], 1, host_out)  # warm
        barrier()
        start = time.perf_counter()
        cells, h2d, d2h = 0, 0, 0
        current, n_current = host_in, n_in
        for _ in range(e2e_steps):
            n_out = sim.step_host(current[:n_current], spec["dt"], 1, host_out)
            cells += n_current
            h2d += n_current * lanes * 4
            d2h += n_out * lanes * 4
            # next step's input is this step's output; the ABI consumes the
            # input before it writes the output, so the buffer may be shared
            current, n_current = host_out, n_out
        barrier()
        seconds = time.perf_counter() - start
        if use_dist:
            t = torch.tensor([seconds, float(cells)], dtype=torch.float64,
                             device="cuda")
            worst = t.clone()
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            seconds, cells = float(worst[0]), int(t[1])
        e2e = {"value": cells / seconds, "unit": "cell-updates/s",
               "h2d_bytes_per_step": h2d // e2e_steps,
               "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps}

    # ---- roofline of the dominant kernel (product arm) ---------------------------
    roofline = None
    launches_per_step = None
    if not is_reference:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_kind = json.load(open(peaks_path))["hbm_gbs"], "measured"
        else:
            peak, peak_kind = 6650.0, "fallback"
        sim.close()
        sim = new_sim(lib, spec, X, types, gs)
        sim.step(spec["dt"], 2)
        sim.profile_sweeps(True)
        n_before = sim.n()
        sim.step(spec["dt"], 4)
        sweep_ms, sweep_launches = sim.read_sweep_profile()
        n_after = sim.n()
        sim.profile_sweeps(False)
        cells_per_launch = 0.5 * (n_before + n_after)
        # one sweep launch reads state + old velocities and writes dX:
        # 2 * sizeof(Pt) + 12 bytes per cell (kernels)
        bytes_per_launch = cells_per_launch * (2 * LANES_BYTES[lanes] + 12)
        avg_ms = sweep_ms / max(sweep_launches, 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "sweep_cubes",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_kind": peak_kind,
                    "traffic": None, "avg_launch_ms": avg_ms,
                    "share_of_step": 2 * avg_ms / (ms / steps),
                    "step_frac": value / world * (9 * LANES_BYTES[lanes] + 36)
                    / 1e9 / peak,
                    "note": "instruction-issue bound, not HBM bound; see "
                            "DESIGN.md and profiles/"}
        # kernels of this repo per model step: stage 1 bin_cells, scan_bins,
        # place_ids, reorder_cells, sweep_cubes, predictor_step; stage 2 the
        # same minus bin_cells (fused into the predictor), corrector_step
        launches_per_step = 11 + (1 if spec["model"] == "growth" else 0)
    sim.close()

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    line = {
        "metric": "Heun-step cell-updates/s (Grid_solver)",
        "value": value, "unit": "cell-updates/s", "n_gpus": world if not
        is_reference else 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": spec["model"],
                   "cells_start": spec["n"], "cells_end": n_end,
                   "n_max": spec["n_max"], "grid_size": gs, "dt": spec["dt"],
                   "tissue": "jittered FCC ball, shuffled order, seeded",
                   "replicas": world if not is_reference else 1,
                   "l2": "working set (>300 MB per replica) exceeds the 126 MB L2"},
        "clocks": clocks,
    }
    if is_reference:
        line["impl"] = "reference"
        kind = "port" if fallback_to_port else "reference"
        line["cpu_baseline"] = {
            "value": value, "unit": "cell-updates/s", "kind": kind,
            "cores": (os.cpu_count() if fallback_to_port else 0),
            "sample": ("CPU oracle port, division off" if fallback_to_port else
                       "the reference's own CUDA build (unmodified headers, "
                       "sm_100a) on the full workload; ya||a has no CPU path")}
        line["e2e"] = {"value": value, "unit": "cell-updates/s",
                       "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        line["gpu_launches"] = 0
    else:
        line["e2e"] = e2e
        line["roofline"] = roofline
        line["gpu_launches"] = launches_per_step * steps
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(spec)
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
