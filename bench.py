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
    # configs[3]'s ingredients at 1 M cells: 7-float branching cell (Turing
    # reaction-diffusion + bending + atomic neighbour counters), and protrusion
    # links as generic force (one link per cell)
    "branching_1M": dict(model="branching", n=1_000_000, n_max=1_000_000, d=0.75,
                         dt=0.2, params={}, typed=True),
    "protrusions_1M": dict(model="protrusions", n=1_000_000, n_max=1_000_000,
                           d=0.8, dt=0.1, params={"link_strength": 0.2},
                           typed=False, links_per_cell=1),
    # configs[3] at its named size on one GPU (Cell = 28 B, n_max 12.5 M)
    "branching_10M": dict(model="branching", n=10_000_000, n_max=10_000_000,
                          d=0.75, dt=0.2, params={}, typed=True),
    # configs[3] for real: branching cell + division + one protrusion per cell
    # rewired every step (Grid::build + curand) pulling through link_forces
    "branching_growth_1M": dict(model="branching_growth", n=1_000_000,
                                n_max=2_097_152, d=0.75, dt=0.2,
                                params={"mes_rate": 0.006, "epi_rate": 0.006,
                                        "seed": 4}, typed=True),
    "branching_growth_10M": dict(model="branching_growth", n=10_000_000,
                                 n_max=12_582_912, d=0.75, dt=0.2,
                                 params={"mes_rate": 0.006, "epi_rate": 0.006,
                                         "seed": 4}, typed=True),
    # Gabriel_solver: the relu force on Gabriel neighbours only
    "gabriel_1M": dict(model="relu_gabriel", n=1_000_000, n_max=1_000_000, d=0.8,
                       dt=0.1, params={}, typed=False),
    "growth_100k": dict(model="growth", n=100_000, n_max=262_144, d=0.75,
                        dt=0.2, params={"prolif_rate": 0.006, "mean_dist": 0.75,
                                        "seed": 2}, typed=True),
}
# One tissue cut into slabs, one per GPU (yalla_b200/dd.py): weak scaling,
# 12.5 M float3 cells per rank; at 8 GPUs this is configs[4], the 100 M sphere.
DD_WORKLOADS = {
    "sphere_dd": dict(model="relu_grid", cells_per_gpu=12_500_000, d=0.8, dt=0.1),
    "sphere_dd_2M": dict(model="relu_grid", cells_per_gpu=2_000_000, d=0.8, dt=0.1),
}
LANES_BYTES = {3: 12, 5: 20, 7: 28}


def make_state(spec, seed):
    rng = np.random.default_rng(seed)
    lanes = yb.MODEL_LANES[spec["model"]]
    if lanes == 3:
        X = workloads.lattice_ball(spec["n"], spec["d"], rng)
    else:
        X = np.zeros((spec["n"], lanes), dtype=np.float32)
        X[:, :5] = workloads.polarized_ball(spec["n"], spec["d"], rng, lattice=True,
                                            noise=0.0 if spec["typed"] else 0.5)
        if lanes > 5:  # morphogen concentrations u, v
            X[:, 5:] = rng.random((spec["n"], lanes - 5)).astype(np.float32) * 0.2
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
    if spec.get("links_per_cell"):
        sim.set_links(workloads.random_links(
            X, spec["links_per_cell"] * len(X), 2.0, np.random.default_rng(77)))
    sim.set_state(X)
    return sim


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read in-process
    through NVML (nvidia_ml_py) from a background thread. NVML is initialised
    in the constructor, i.e. before the warm-up: starting an nvidia-smi process
    next to the timed region stalls concurrent cudaMalloc/cudaFree calls -- the
    reference build does four of each per step -- by tens of milliseconds."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40,
               "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device_index, interval_s=0.5):
        import threading
        self.interval_s = float(os.environ.get("YALLA_BENCH_CLOCK_S", interval_s))
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.handle = None
        self.stop_flag = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = device_index
            if visible:
                index = int(visible.split(",")[device_index])
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(
                self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def _sample(self):
        try:
            self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(
                self.handle, self.nvml.NVML_CLOCK_SM)))
            mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            for name, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        # sparse on purpose: every NVML query contends with the driver lock that
        # cudaMalloc/cudaFree (reference arm: several per step) also need
        self.stop_flag.wait(0.02)
        while not self.stop_flag.is_set():
            self._sample()
            self.stop_flag.wait(self.interval_s)

    def start(self):
        if self.handle is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        self._sample()  # at least one sample from inside the region
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.sm_max, "samples": len(self.sm),
                "reasons": sorted(self.reasons)}


# FP32 lane-instructions of the pairwise functors per accepted pair (SURVEY.md
# 8(d): counted from the SASS of the reference build; bending-type functors
# include their MUFU expansions). The growth functor bends only epithelium-
# epithelium pairs, a thin shell of the tissue.
FUNCTOR_LANE_INSTR = {"relu_grid": 12, "relu_gabriel": 12, "spring_grid": 8, "protrusions": 12,
                      "growth": 20, "epithelium": 250, "branching": 270,
                      "branching_growth": 270}


def pair_statistics(X, cube_size=1.0, sample=200_000):
    """Exact mean number of candidates (cells in the 27 surrounding cubes, self
    included) and of accepted pairs (distance < cube_size, self included) per
    cell -- the C and N of SURVEY.md 8(d). Candidates from the cube histogram
    of the whole tissue, accepted pairs on a random sample of cells."""
    from scipy.ndimage import uniform_filter
    from scipy.spatial import cKDTree
    pos = X[:, :3].astype(np.float64)
    cube = np.floor(pos / cube_size).astype(np.int64)
    cube -= cube.min(axis=0) - 1
    shape = tuple(cube.max(axis=0) + 2)
    counts = np.zeros(shape, dtype=np.float64)
    np.add.at(counts, (cube[:, 0], cube[:, 1], cube[:, 2]), 1.0)
    around = uniform_filter(counts, size=3, mode="constant") * 27.0
    candidates = float(np.rint(around[cube[:, 0], cube[:, 1], cube[:, 2]]).mean())
    tree = cKDTree(pos)
    pick = np.random.default_rng(5).choice(len(pos), size=min(sample, len(pos)),
                                           replace=False)
    accepted = tree.query_ball_point(pos[pick], cube_size * (1 - 1e-7),
                                     return_length=True, workers=-1)
    return candidates, float(np.mean(accepted))


def sweep_traffic(workload):
    """DRAM bytes per sweep (sweep_cubes, or list_cubes + interact_lists) from
    the committed `ncu --set full` captures of this workload
    (profiles/r02_sweep_traffic.json, else round 1's), or None."""
    for name in ("r02_sweep_traffic.json", "r01_sweep_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            entry = json.load(open(path)).get(workload)
            if entry is not None:
                return entry["dram_bytes_per_launch"]
    return None


def cpu_baseline(spec, steps=10, sample_cells=1_000_000):
    """The oracle (a CPU port of the reference's algorithm) on a bounded sample
    of the workload: the same kind of tissue at up to sample_cells cells for a
    few steps (about 10-20 s of host time with 16 cores), division switched off
    (no curand on the host)."""
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


def pin_to_gpu_cpus(local_rank):
    """Bind this rank to the CPU cores next to its GPU (NVML's affinity mask),
    before any pinned host buffer is allocated: first touch then places the
    buffers on the GPU's NUMA node, and the ranks of a multi-GPU run stop
    sharing one socket's memory controllers for their host copies."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def run_decomposed(steps, warmup, spec, workload, rank, local_rank, world,
                   cells_total=None, e2e=True):
    """One tissue over `world` GPUs: bricks (slabs for 2, 2x2x1 for 4, 2x2x2 for
    8), halo exchange, migration and drift sum by the library's own kernels over
    peer memory (include/b200/domain.cuh); torch.distributed only carries the
    CUDA IPC handles at set-up. The process group must be up when world > 1.
    Returns the bench record on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from yalla_b200 import dd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d, dt = spec["d"], spec["dt"]
    n_target = cells_total or spec["cells_per_gpu"] * world
    density = np.sqrt(2.0) / d ** 3  # FCC with nearest-neighbour distance d
    radius = (n_target / density * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    gs = int(np.ceil(2 * (radius + d))) + 4
    gs += gs % 2
    bricks = dd.brick_grid_for(world)
    cuts = dd.ball_brick_cuts(radius, bricks)
    # a face is at most the ball's cross-section divided among the bricks that
    # tile it; its halo strip is 1.5 cubes thick
    across = sorted(bricks)
    face_cells = int(np.pi * radius ** 2 / (across[0] * across[1]) * 1.5 * density)
    n_faces = sum(1 for b in bricks if b > 1) * (2 if max(bricks) > 2 else 1)
    n_max = int(n_target / world * 1.08) + n_faces * int(face_cells * 1.2) + 4096
    lib = yb.product()
    transport = ("bricks %dx%dx%d, halo exchange + migration + drift sum by kernels "
                 "over peer memory (NVLink P2P), no NCCL in a step" % bricks)
    try:
        domain = dd.BrickDomain(lib, spec["model"], n_max, gs, 1.0, bricks, cuts,
                                rank, world, face_capacity=int(face_cells * 1.3) + 4096)
        domain.connect_over_ipc()
        n_seeded = domain.seed_lattice_ball(radius, d, seed=20261017)
        assert n_seeded <= n_max, (n_seeded, n_max)
    except yb.YallaError as error:
        # no peer mapping between the ranks on this box (raised on every rank):
        # round 1's transport, z-slabs over torch.distributed send/recv
        if rank == 0:
            print(f"bench: {error}; falling back to slabs over NCCL", file=sys.stderr)
        transport = (f"z-slabs x{world}, halo exchange + migration over NCCL "
                     "send/recv, drift all-reduce (fallback: no CUDA IPC; tissue "
                     "generated on the host)")
        bounds = [-np.inf] + dd.ball_slab_cuts(radius, world) + [np.inf]
        mine = dd.lattice_ball_slab(radius, d, bounds[rank], bounds[rank + 1],
                                    np.random.default_rng(1000 + rank))
        slab_face = int(1.5 * np.pi * radius ** 2 * density)
        slabs = dd.SlabDomain(lib, spec["model"], int(len(mine) * 1.05) + 2 * slab_face
                              + 1024, gs, 1.0, bounds[rank], bounds[rank + 1], "cuda",
                              halo_capacity=slab_face * 13 // 10 + 4096)
        slabs.set_cells(mine)

        class Stepped:  # the brick driver's interface on the slab driver
            sim = slabs.sim
            counts = staticmethod(slabs.counts)
            owned_state = staticmethod(slabs.owned_state)
            set_cells = staticmethod(slabs.set_cells)
            close = staticmethod(slabs.close)

            @staticmethod
            def step(dt_, n_steps=1):
                for _ in range(n_steps):
                    slabs.step(dt_)
        domain = Stepped
    sampler = ClockSampler(local_rank, 0.05)
    domain.step(dt, warmup)

    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    start.record()
    domain.step(dt, steps)
    stop.record()
    barrier()
    clocks = sampler.stop()
    ms = start.elapsed_time(stop)
    n_mine, with_ghosts, problems = domain.counts()

    # end to end: host buffers in and out every step
    e2e_seconds, e2e_steps, n_out = 0.0, 0, 0
    if e2e:
        host_X = domain.owned_state()[0].cpu().pin_memory()
        e2e_steps = max(2, steps // 4)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            domain.set_cells(host_X)
            domain.step(dt)
            host_out = domain.owned_state()[0].cpu()
        barrier()
        e2e_seconds = time.perf_counter() - t0
        n_out = len(host_out)

    # the dominant kernel, timed alone
    domain.sim.profile_sweeps(True)
    domain.step(dt, 3)
    sweep_ms, sweep_launches = domain.sim.read_sweep_profile()
    try:
        phases = {name: value / 3 for name, value in
                  domain.sim.dom_read_profile().items()}
    except yb.YallaError:
        phases = None
    domain.sim.profile_sweeps(False)

    stats = torch.tensor([ms, e2e_seconds, float(n_mine), float(n_out),
                          float(with_ghosts - n_mine), float(problems)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        worst = stats.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        ms, e2e_seconds = float(worst[0]), float(worst[1])
    cells_total = int(stats[2])
    ghosts_total, problems = int(stats[4]), int(stats[5])
    barrier()
    domain.close()
    if rank != 0:
        return None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(
        peaks_path) else 6650.0
    avg_ms = sweep_ms / max(sweep_launches, 1)
    achieved = with_ghosts * (2 * 12 + 12) / (avg_ms * 1e-3) / 1e9
    value = cells_total * steps / (ms * 1e-3)
    record = 4 * (yb.MODEL_LANES[spec["model"]] + 3)
    line = {
        "metric": "Heun-step cell-updates/s (Grid_solver)", "value": value,
        "unit": "cell-updates/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "model": spec["model"],
                   "cells_total": cells_total,
                   "cells_per_gpu": cells_total // world, "grid_size": gs,
                   "dt": dt, "parallelism": transport,
                   "tissue": "jittered FCC ball, seeded on the device",
                   "l2": "working set (>1 GB per rank) exceeds the 126 MB L2"},
        "clocks": clocks,
        "ghost_cells": ghosts_total,
        "phase_ms_per_step_rank0": phases,
        # two halo rounds per step: every ghost record crosses NVLink twice
        "nvlink_bytes_per_step": 2 * ghosts_total * record,
        "roofline": {"bound": "hbm", "kernel": "sweep_cubes",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None,
                     "avg_launch_ms": avg_ms,
                     "share_of_step": 2 * avg_ms / (ms / steps),
                     "note": "rank 0; instruction-issue bound, see DESIGN.md"},
        # per step and rank: 3 x (dd_select, dd_wait), 2 x dd_append_ghosts,
        # 2 x (bin, scan, place, settle, sweep, dd_allreduce_drift, update),
        # dd_merge, commit
        "gpu_launches": 24 * steps * world,
        "problems": problems,
    }
    if e2e:
        line["e2e"] = {"value": cells_total * e2e_steps / e2e_seconds,
                       "unit": "cell-updates/s",
                       "h2d_bytes_per_step": cells_total * 12,
                       "d2h_bytes_per_step": cells_total * 12, "steps": e2e_steps}
    return line


def run_decomposed_model(steps, warmup, workload, rank, local_rank, world):
    """BASELINE.json configs[3] on `world` GPUs: ONE tissue of the typed model
    (branching cell + Property arrays + division + protrusions rewired per step)
    cut into bricks -- strong scaling, the cell count is fixed. Property arrays,
    cell identities and links travel with the cells (Solution::dom_register_array,
    include/b200/brick_links.cuh). Returns the record on rank 0."""
    import torch
    import torch.distributed as dist
    from yalla_b200 import dd

    spec = WORKLOADS[workload]
    X, types, gs = make_state(spec, seed=1000)  # the same tissue on every rank
    halo = 2.5 if spec["model"] == "branching_growth" else 1.5
    bricks = dd.brick_grid_for(world)
    radius = float(np.max(np.linalg.norm(X[:, :3], axis=1)))
    density = np.sqrt(2.0) / spec["d"] ** 3
    across = sorted(bricks)
    face_cells = int(np.pi * radius ** 2 / (across[0] * across[1]) * halo * density)
    n_faces = sum(1 for b in bricks if b > 1) * (2 if max(bricks) > 2 else 1)
    share = spec["n_max"] / spec["n"]
    n_max = int(len(X) / world * share * 1.05) + n_faces * int(face_cells * 1.3) + 8192
    lib = yb.product()
    domain = dd.BrickDomain(lib, spec["model"], n_max, gs, 1.0, bricks,
                            dd.ball_brick_cuts(radius, bricks), rank, world,
                            face_capacity=int(face_cells * 1.4) + 8192, halo=halo)
    domain.connect_over_ipc()
    for key, value in spec["params"].items():
        domain.sim.set_param(key, value)
    mine = domain.owns(X)
    domain.set_cells(X[mine])
    domain.sim.set_ints("type", types[mine])
    del X
    dt = spec["dt"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    domain.step(dt, warmup)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n_before = domain.counts()[0]
    start.record()
    domain.step(dt, steps)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    n_after, with_ghosts, problems = domain.counts()
    unresolved = 0
    if spec["model"] == "branching_growth":
        unresolved = int(domain.sim.get_ints("unresolved_links")[0])
    stats = torch.tensor([ms, 0.5 * (n_before + n_after), float(n_after),
                          float(with_ghosts - n_after), float(problems),
                          float(unresolved)], dtype=torch.float64, device="cuda")
    if world > 1:
        worst = stats.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        ms = float(worst[0])
    barrier()
    domain.close()
    if rank != 0:
        return None
    return {"value": float(stats[1]) * steps / (ms * 1e-3), "unit": "cell-updates/s",
            "n_gpus": world, "steps": steps, "ms_per_step": ms / steps,
            "scaling": "strong", "cells_end": int(stats[2]),
            "ghost_cells": int(stats[3]), "problems": int(stats[4]),
            "unresolved_link_stages": int(stats[5]),
            "config": {"workload": workload, "model": spec["model"],
                       "cells_start": spec["n"], "bricks": list(bricks),
                       "halo": halo, "dt": dt, "grid_size": gs,
                       "parallelism": "bricks %dx%dx%d over peer memory; Property "
                       "arrays, cell identities and links travel with the cells"
                       % bricks}}


def time_reference(lib, spec, X, types, gs, steps, warmup, repeats):
    """Best of `repeats` runs of the reference build; every run starts from the
    same state as the product arm (state and types reloaded, `warmup` untimed
    steps, then `steps` timed ones). The reference allocates and frees Thrust
    temporaries every step and single timings scatter by up to 10x on this
    pool, so the best run is the conservative bar."""
    best = None
    with new_sim(lib, spec, X, types, gs) as sim:
        for _ in range(repeats):
            if types is not None:
                sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(spec["dt"], warmup)
            sim.sync()
            ms, updates = sim.step_timed(spec["dt"], steps)
            if best is None or updates / ms > best[1] / best[0]:
                best = (ms, updates, sim.n())
    return best


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=30)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", default="product",
                        choices=["product", "reference"])
    parser.add_argument("--workload", default="growth_1M",
                        choices=sorted(WORKLOADS) + sorted(DD_WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-decomposed", action="store_true",
                        help="skip the slab-decomposed sphere record")
    args = parser.parse_args()
    warmup = max(args.warmup, 3)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    is_reference = args.impl == "reference"

    if is_reference and rank != 0:
        return  # the reference arm runs on rank 0 alone
    cpus = pin_to_gpu_cpus(local_rank if not is_reference else 0)

    use_dist = world > 1 and not is_reference
    torch.cuda.set_device(local_rank if not is_reference else 0)
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload in DD_WORKLOADS:
        dd_spec = DD_WORKLOADS[args.workload]
        if not is_reference:
            line = run_decomposed(args.steps, warmup, dd_spec, args.workload,
                                  rank, local_rank, world)
            if rank == 0:
                print(json.dumps(line))
            if use_dist:
                dist.destroy_process_group()
            return
        # the reference is single-GPU: it integrates one rank's share
        spec = dict(model=dd_spec["model"], n=dd_spec["cells_per_gpu"],
                    n_max=dd_spec["cells_per_gpu"], d=dd_spec["d"],
                    dt=dd_spec["dt"], params={}, typed=False)
    else:
        spec = WORKLOADS[args.workload]

    X, types, gs = make_state(spec, seed=1000 + rank)
    lanes = X.shape[1]
    steps = args.steps
    config = {"workload": args.workload, "model": spec["model"],
              "cells_start": spec["n"], "n_max": spec["n_max"], "grid_size": gs,
              "dt": spec["dt"],
              "tissue": "jittered FCC ball, shuffled order, seeded",
              "l2": "working set (>300 MB per tissue) exceeds the 126 MB L2",
              "parallelism": "1 GPU" if (is_reference or world == 1) else
              f"independent tissues x{world} (+ one decomposed tissue, see "
              "`decomposed`)"}

    # ======================= the reference arm ============================
    if is_reference:
        sampler = ClockSampler(0, 0.5)  # sparse: NVML stalls cudaMalloc/cudaFree
        if os.path.exists(yb.REFERENCE_LIB):
            builds = {"-O3 (README flags, asserts on)": yb.REFERENCE_LIB}
            if os.path.exists(yb.REFERENCE_NDEBUG_LIB):
                builds["-O3 -DNDEBUG"] = yb.REFERENCE_NDEBUG_LIB
            sampler.start()
            timings = {}
            for flags, path in builds.items():
                timings[flags] = time_reference(yb.load(path), spec, X, types, gs,
                                                steps, warmup, repeats=3)
            clocks = sampler.stop()
            flags = max(timings, key=lambda f: timings[f][1] / timings[f][0])
            ms, updates, n_end = timings[flags]
            kind, cores = "reference", 0
            sample = ("the reference's own CUDA build (unmodified headers, "
                      f"sm_100a, {flags}: the faster of {len(builds)} builds) "
                      "on the full workload; ya||a has no CPU path")
            builds_report = {f: t[1] / (t[0] * 1e-3) for f, t in timings.items()}
        else:
            # no reference build travelled with the snapshot: the CPU port
            port = dict(spec, params=dict(spec["params"], prolif_rate=0.0)) \
                if "prolif_rate" in spec["params"] else spec
            steps = min(steps, 2)
            sampler.start()
            with new_sim(yb.load(ORACLE_LIB), port, X, types, gs) as sim:
                ms, updates = sim.step_timed(spec["dt"], steps)
                n_end = sim.n()
            clocks = sampler.stop()
            kind, cores = "port", os.cpu_count()
            sample, builds_report = "CPU oracle port, division off", None
        value = updates / (ms * 1e-3)
        line = {
            "metric": "Heun-step cell-updates/s (Grid_solver)", "value": value,
            "unit": "cell-updates/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "impl": "reference", "cells_end": n_end,
            "cpu_baseline": {"value": value, "unit": "cell-updates/s",
                             "kind": kind, "cores": cores, "sample": sample},
            "e2e": {"value": value, "unit": "cell-updates/s",
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "reference_timing": "best of 3 runs per build, each from the "
                                "product arm's start state (allocation jitter)",
            "reference_builds": builds_report,
        }
        print(json.dumps(line))
        return

    # ======================= the product arm ==============================
    lib = yb.product()
    assert lib.build_info.startswith("yalla-b200"), lib.build_info
    assert os.path.realpath(lib.path) == os.path.realpath(os.path.join(
        ROOT, "yalla_b200", "_lib", "libyalla_b200.so")), lib.path
    # the product arm's steps make no driver calls NVML could stall: 10 ms
    sampler = ClockSampler(local_rank, 0.01)
    sim = new_sim(lib, spec, X, types, gs)
    sim.step(spec["dt"], warmup)
    sim.sync()

    # ---- device-resident throughput ---------------------------------------
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
    # Every step is one independent batch: a tissue is uploaded from pinned host
    # memory, integrated for one step and downloaded again, all inside the timed
    # region. yb_sim_step_host_async pipelines the batches through ONE model
    # instance: two copy streams and two device-side staging slots let the
    # upload of batch k + 1 and the download of batch k - 1 overlap the kernels
    # of batch k. Checked afterwards against the same batches run one by one.
    host_in = torch.from_numpy(X).pin_memory().numpy()
    n_in = len(host_in)
    out_cells = min(spec["n_max"], n_in + n_in // 32)  # room for division
    outs = [torch.zeros((out_cells, lanes), dtype=torch.float32
                        ).pin_memory().numpy() for _ in range(2)]
    counts = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
    e2e_steps = max(4, steps // 2)

    def batch(k):
        sim.step_host_async(host_in, spec["dt"], 1, outs[k % 2], out_cells,
                            counts[k % 2].data_ptr())

    if types is not None:
        sim.set_ints("type", types)
    sim.set_state(X)       # zero velocities: the batches start from a known state
    for k in range(2):
        batch(k)           # warm the pipeline
    sim.host_drain()
    barrier()
    start = time.perf_counter()
    for k in range(e2e_steps):
        batch(k)
    sim.host_drain()
    barrier()
    seconds = time.perf_counter() - start
    last = (e2e_steps - 1) % 2
    pipelined = outs[last][:n_in].copy()
    assert int(counts[last][0]) >= n_in and np.all(np.isfinite(pipelined))
    # the same sequence of batches, one at a time: the cells that existed at the
    # start of a batch must come out bit-identical (daughters land in slots
    # chosen by atomics, in either mode)
    if types is not None:
        sim.set_ints("type", types)
    sim.set_state(X)
    check = np.zeros((spec["n_max"], lanes), dtype=np.float32)
    for k in range(e2e_steps + 2):
        sim.step_host(host_in, spec["dt"], 1, check)
    e2e_verified = bool(np.array_equal(check[:n_in], pipelined))
    # what the host link gives this rank while all ranks copy at once: the same
    # two transfers per batch, nothing else (bounds e2e from above)
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    pinned_in = torch.from_numpy(host_in)
    device_in = torch.empty_like(pinned_in, device="cuda")
    device_out = torch.zeros((out_cells, lanes), dtype=torch.float32, device="cuda")
    pinned_out = torch.from_numpy(outs[0])
    barrier()
    copy_start = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(up_stream):
            device_in.copy_(pinned_in, non_blocking=True)
        with torch.cuda.stream(down_stream):
            pinned_out.copy_(device_out, non_blocking=True)
    up_stream.synchronize()
    down_stream.synchronize()
    barrier()
    copy_seconds = (time.perf_counter() - copy_start) / 10
    cells = n_in * e2e_steps
    if use_dist:
        t = torch.tensor([seconds, float(cells), copy_seconds],
                         dtype=torch.float64, device="cuda")
        worst = t.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        seconds, cells, copy_seconds = float(worst[0]), int(t[1]), float(worst[2])
    e2e = {"value": cells / seconds, "unit": "cell-updates/s",
           "h2d_bytes_per_step": n_in * lanes * 4,
           "d2h_bytes_per_step": out_cells * lanes * 4 + 4,
           "steps": e2e_steps, "matches_unpipelined_run": e2e_verified,
           # a batch's two copies alone, all ranks at once: e2e cannot beat
           # n_in / this; it falls with N when the ranks share the host's links
           "copies_alone_ms_per_batch": copy_seconds * 1e3,
           "copies_alone_bound": n_in * world / copy_seconds,
           "how": "independent batches pipelined through one model instance: "
                  "2 copy streams + 2 device staging slots, pinned host buffers"}
    sim.close()
    sim = new_sim(lib, spec, X, types, gs)

    # ---- roofline of the dominant kernel --------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_kind = json.load(open(peaks_path))["hbm_gbs"], "measured"
    else:
        peak, peak_kind = 6650.0, "fallback"
    sim.step(spec["dt"], 2)
    sim.profile_sweeps(True)
    n_before = sim.n()
    sim.step(spec["dt"], 4)
    sweep_ms, sweep_launches = sim.read_sweep_profile()
    n_after = sim.n()
    sim.profile_sweeps(False)
    sim.close()
    cells_per_launch = 0.5 * (n_before + n_after)
    avg_ms = sweep_ms / max(sweep_launches, 1)
    # HBM: one sweep launch reads state + old velocities and writes dX:
    # 2 * sizeof(Pt) + 12 bytes per cell (kernels)
    bytes_per_launch = cells_per_launch * (2 * LANES_BYTES[lanes] + 12)
    hbm_achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    hbm = {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s",
           "frac": hbm_achieved / peak, "peak_kind": peak_kind,
           "algorithmic_bytes_per_launch": bytes_per_launch,
           "step_frac": value / world * (9 * LANES_BYTES[lanes] + 36) / 1e9 / peak}
    roofline = None
    if rank == 0:
        # The binding ceiling: algorithmic FP32 lane-instructions
        # per sweep and cell, I = C * 7 + N * (13 + f_pw), against the
        # microbenchmarked FP32 issue rate of the chip. The HBM figure is
        # reported next to it: this path is not bandwidth bound.
        candidates, accepted = pair_statistics(X)
        per_cell = candidates * 7 + accepted * (
            13 + FUNCTOR_LANE_INSTR[spec["model"]])
        fp32_peak = json.load(open(os.path.join(
            ROOT, "profiles", "r01_microbench_peaks.json")))[
                "fp32_lane_instr_per_s_T"]
        fp32_achieved = per_cell * cells_per_launch / (avg_ms * 1e-3) / 1e12
        roofline = {
            "bound": "fp32_issue",
            "kernel": "sweep_cubes" if lanes <= 4 else
            "list_cubes + interact_lists (the sweep of points with extra lanes)",
            "achieved": fp32_achieved, "peak": fp32_peak,
            "unit": "T lane-instr/s", "frac": fp32_achieved / fp32_peak,
            "peak_kind": "microbenchmark (profiles/r01_microbench_peaks.json)",
            "candidates_per_cell": candidates, "accepted_per_cell": accepted,
            "lane_instr_per_cell_and_sweep": per_cell,
            "traffic": sweep_traffic(args.workload),
            "avg_launch_ms": avg_ms,
            "share_of_step": 2 * avg_ms / (ms / steps),
            "hbm": hbm,
            "note": "the sweep is instruction-issue / latency bound, not HBM "
                    "bound (DESIGN.md section 3); both ceilings are reported"}

    # ---- the same N GPUs on ONE tissue: the slab-decomposed sphere ---------------
    # (configs[4]: 12.5 M float3 cells per GPU, 100 M at 8 GPUs; on one GPU also
    # the whole 100 M-cell sphere, the T1 of the strong-scaling figure)
    decomposed = None
    if not args.no_decomposed:
        try:
            dd_spec = DD_WORKLOADS["sphere_dd"]
            dd_steps = max(3, min(steps, 10))
            record = run_decomposed(dd_steps, 3, dd_spec, "sphere_dd", rank,
                                    local_rank, world, e2e=False)
            if rank == 0:
                decomposed = {key: record[key] for key in (
                    "value", "unit", "n_gpus", "steps", "ms_per_step", "ghost_cells",
                    "nvlink_bytes_per_step", "phase_ms_per_step_rank0", "problems",
                    "config")}
                decomposed["sweep_ms_per_launch"] = record["roofline"]["avg_launch_ms"]
            if world == 1 and os.environ.get("YALLA_BENCH_STRONG", "1") != "0":
                whole = run_decomposed(3, 3, dd_spec, "sphere_dd", 0, local_rank, 1,
                                       cells_total=100_000_000, e2e=False)
                decomposed["whole_sphere_on_one_gpu"] = {
                    key: whole[key] for key in ("value", "ms_per_step", "steps",
                                                "config")}
        except Exception as error:  # the headline line must not depend on it
            if rank == 0:
                print(f"bench: decomposed record failed: {error!r}", file=sys.stderr)
                decomposed = dict(decomposed or {}, error=repr(error))
        # configs[3] on the same N GPUs: the 10 M-cell branching tissue with
        # division and protrusions, one tissue cut into bricks (strong scaling)
        if os.environ.get("YALLA_BENCH_CONFIG3", "1") != "0":
            try:
                model_record = run_decomposed_model(
                    5, 3, "branching_growth_10M", rank, local_rank, world)
                if rank == 0:
                    decomposed = dict(decomposed or {}, configs3=model_record)
            except Exception as error:
                if rank == 0:
                    print(f"bench: decomposed configs[3] failed: {error!r}",
                          file=sys.stderr)
                    decomposed = dict(decomposed or {}, configs3={"error": repr(error)})

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    # kernels of this repo per model step: stage 1 bin_cells, scan_bins,
    # place_ids, reorder_cells, sweep_cubes, predictor_step; stage 2 the same
    # minus bin_cells (fused into the predictor), corrector_step; models with a
    # counter reset add zero_cells per stage, the growth model snapshot_count
    # and proliferate
    launches_per_step = 11 + {"growth": 4, "branching": 2,
                              "branching_growth": 18}.get(spec["model"], 0)
    line = {
        "metric": "Heun-step cell-updates/s (Grid_solver)",
        "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "clocks": clocks, "cells_end": n_end, "cpus": cpus,
        "e2e": e2e, "roofline": roofline,
        "gpu_launches": launches_per_step * steps * world,  # all ranks
        "decomposed": decomposed,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(spec)
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
