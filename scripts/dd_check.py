"""Multi-GPU check and timing of the slab decomposition (run under torchrun):

    torchrun --nproc-per-node N scripts/dd_check.py check            # parity
    torchrun --nproc-per-node N scripts/dd_check.py bench <cells_per_gpu> [steps]

check: a 300k-cell tissue is integrated by N slabs and, on rank 0, by one
solver; the cell sets must agree. bench: weak scaling of a float3 relu_force
ball with cells_per_gpu cells per rank (N = 8, 12.5 M -> the 100 M-cell sphere
of BASELINE.json configs[4]).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import dd, workloads  # noqa: E402


def setup():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def fcc_radius(n, d):
    return (n * d ** 3 / np.sqrt(2.0) * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)


def check(rank, world):
    from scipy.spatial import cKDTree
    n, d, dt, steps = 300_000, 0.8, 0.1, 5
    X = workloads.lattice_ball(n, d, np.random.default_rng(5))
    gs = workloads.grid_size_for(n, d)
    cuts = dd.ball_slab_cuts(np.abs(X[:, 2]).max(), world)
    bounds = [-np.inf] + cuts + [np.inf]
    lib = yb.product()
    mine = X[(X[:, 2] >= bounds[rank]) & (X[:, 2] < bounds[rank + 1])]
    domain = dd.SlabDomain(lib, "relu_grid", n, gs, 1.0, bounds[rank],
                           bounds[rank + 1], "cuda")
    domain.set_cells(mine)
    for _ in range(steps):
        domain.step(dt)
    got = domain.gather_all()
    owned, total, problems = domain.counts()
    ghosts, migrated = total - owned, abs(owned - len(mine))
    assert problems == 0, problems
    domain.close()
    if rank == 0:
        with lib.sim("relu_grid", n, gs, 1.0) as sim:
            sim.set_state(X)
            sim.step(dt, steps)
            want = sim.get_state()
        distance, index = cKDTree(want).query(got, k=1)
        ok = (len(np.unique(index)) == n and distance.max() < 1e-4)
        print(json.dumps({"check": "slabs vs single solver", "world": world,
                          "cells": n, "steps": steps, "ok": bool(ok),
                          "max_deviation": float(distance.max()),
                          "ghosts_rank0": ghosts, "net_migration_rank0": migrated}))
        assert ok


def bench(rank, world, cells_per_gpu, steps, warmup=3):
    d, dt = 0.8, 0.1
    n_total = cells_per_gpu * world
    radius = fcc_radius(n_total, d)
    gs = int(np.ceil(2 * (radius + d))) + 4
    gs += gs % 2
    cuts = dd.ball_slab_cuts(radius, world)
    bounds = [-np.inf] + cuts + [np.inf]
    rng = np.random.default_rng(100 + rank)
    t0 = time.time()
    mine = dd.lattice_ball_slab(radius, d, bounds[rank], bounds[rank + 1], rng)
    halo_cells = int(2 * 1.5 * np.pi * radius ** 2 * np.sqrt(2) / d ** 3)
    n_max = int(len(mine) * 1.05) + halo_cells + 1024
    lib = yb.product()
    domain = dd.SlabDomain(lib, "relu_grid", n_max, gs, 1.0, bounds[rank],
                           bounds[rank + 1], "cuda",
                           halo_capacity=halo_cells // 2 * 13 // 10 + 4096)
    domain.set_cells(mine)
    n_mine = len(mine)
    setup_s = time.time() - t0
    for _ in range(warmup):
        domain.step(dt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    start = time.perf_counter()
    for _ in range(steps):
        domain.step(dt)
    barrier()
    cells = steps * n_mine  # migration moves a few hundred cells at most
    seconds = time.perf_counter() - start
    stats = torch.tensor([seconds, float(cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        worst = stats.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        seconds, cells = float(worst[0]), float(stats[1])
    total = domain.total_cells()
    owned, with_ghosts, problems = domain.counts()
    ghosts = with_ghosts - owned
    assert problems == 0, problems
    domain.close()
    if rank == 0:
        print(json.dumps({
            "bench": "slab decomposition, float3 relu_force ball, weak scaling",
            "n_gpus": world, "cells_total": total, "cells_per_gpu": total // world,
            "grid_size": gs, "steps": steps, "ms_per_step": 1e3 * seconds / steps,
            "cell_updates_per_s": cells / seconds, "ghosts_rank0": ghosts,
            "setup_s": round(setup_s, 1)}))


def main():
    rank, local, world = setup()
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    try:
        if mode == "check":
            check(rank, world)
        else:
            cells = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
            steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
            bench(rank, world, cells, steps)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
