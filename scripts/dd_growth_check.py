"""The growth model (Po_cell + Property arrays + curand division, BASELINE.json
configs[1] cut into bricks) across real GPUs against the single-domain run.

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node N --master-addr 127.0.0.1 --master-port 29519 \
        scripts/dd_growth_check.py [n_cells] [steps] [timed_steps]

One rank per GPU and brick. The cell type travels with the cells and with their
ghost copies, the neighbour counters and the curand states with the cells
(Solution::dom_register_array). Part 1, division off: cell by cell the same
positions, polarities, types and neighbour counters as one domain. Part 2,
division on: the tissue grows at the same rate (every brick draws from its own
curand streams, so the comparison is statistical). Part 3: ms per step of the
decomposed growing tissue.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import dd, workloads  # noqa: E402

DT = 0.1


def bricks_run(lib, X, types, n_max, gs, rate, steps, rank, world, timed=0):
    bricks = dd.brick_grid_for(world)
    radius = float(np.max(np.linalg.norm(X[:, :3], axis=1)))
    domain = dd.BrickDomain(lib, "growth", n_max, gs, 1.0, bricks,
                            dd.ball_brick_cuts(radius, bricks), rank, world,
                            face_capacity=max(n_max // 4, 4096))
    domain.connect_over_ipc()
    domain.sim.set_param("prolif_rate", rate)
    domain.sim.set_param("seed", 5)
    mine = domain.owns(X)
    domain.set_cells(X[mine])
    domain.sim.set_ints("type", types[mine])
    domain.step(DT, steps)
    torch.cuda.synchronize()
    ms = None
    if timed:
        dist.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
        start.record()
        domain.step(DT, timed)
        stop.record()
        torch.cuda.synchronize()
        ms = torch.tensor([start.elapsed_time(stop) / timed], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
    owned, with_ghosts, problems = domain.counts()
    part = (domain.owned_state()[0].cpu().numpy(), domain.sim.get_ints("type"),
            domain.sim.get_ints("mes_nbs"), domain.sim.get_ints("epi_nbs"),
            int(mine.sum()), owned, with_ghosts, problems)
    parts = [None] * world
    dist.all_gather_object(parts, part)
    dist.barrier()
    domain.close()
    return parts, ms


def one_domain(lib, X, types, n_max, gs, rate, steps):
    with lib.sim("growth", n_max, gs, 1.0) as sim:
        sim.set_param("prolif_rate", rate)
        sim.set_param("seed", 5)
        sim.set_ints("type", types)
        sim.set_state(X)
        sim.step(DT, steps)
        return (sim.get_state(), sim.get_ints("type"), sim.get_ints("mes_nbs"),
                sim.get_ints("epi_nbs"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    timed = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(
        "cuda", int(os.environ["LOCAL_RANK"])))
    lib = yb.product()
    X = workloads.polarized_ball(n, 0.75, np.random.default_rng(6), lattice=True)
    X[:, :3] *= 0.95  # squeezed: the tissue expands, cells cross the cuts
    X = X.astype(np.float32)
    types = workloads.shell_types(X)
    gs = workloads.grid_size_for(n, 0.75, growth=2.5)
    report = {"world": world, "bricks": dd.brick_grid_for(world), "cells": n}

    # 1. division off: cell by cell against one domain
    parts, _ = bricks_run(lib, X, types, n, gs, 0.0, steps, rank, world)
    if rank == 0:
        from scipy.spatial import cKDTree
        got = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        want = one_domain(lib, X, types, n, gs, 0.0, steps)
        distance, index = cKDTree(want[0][:, :3]).query(got[0][:, :3], k=1)
        error = np.abs(got[0] - want[0][index]).max(axis=1)
        report["no_division"] = {
            "cells": int(len(got[0])), "unique_matches": int(len(np.unique(index))),
            "max_deviation": float(error.max()),
            "cells_beyond_1e-4": int(np.sum(error > 1e-4)),
            "type_mismatches": int(np.sum(got[1] != want[1][index])),
            "mes_nbs_mismatches": int(np.sum(got[2] != want[2][index])),
            "epi_nbs_mismatches": int(np.sum(got[3] != want[3][index])),
            "migrated": int(sum(abs(p[4] - p[5]) for p in parts)),
            "ghosts": int(sum(p[6] - p[5] for p in parts)),
            "problems": int(sum(p[7] for p in parts))}

    # 2. + 3. division on: growth statistics and speed
    grow_steps = 4 * steps
    parts, ms = bricks_run(lib, X, types, 3 * n, gs, 0.01, grow_steps, rank, world,
                           timed=timed)
    if rank == 0:
        got_n = sum(len(p[0]) for p in parts)
        got_X = np.concatenate([p[0] for p in parts])
        want = one_domain(lib, X, types, 3 * n, gs, 0.01, grow_steps + timed)
        radius_of_gyration = [float(np.sqrt(np.mean(np.sum(
            (s[:, :3] - s[:, :3].mean(axis=0)) ** 2, axis=1)))) for s in (got_X, want[0])]
        report["division"] = {
            "steps": grow_steps + timed, "cells_bricks": int(got_n),
            "cells_one_domain": int(len(want[0])),
            "relative_difference": float(got_n / len(want[0]) - 1.0),
            "radius_of_gyration": radius_of_gyration,
            "epithelial_fraction": [float(np.mean(np.concatenate(
                [p[1] for p in parts]))), float(np.mean(want[1]))],
            "finite": bool(np.all(np.isfinite(got_X))),
            "problems": int(sum(p[7] for p in parts)),
            "ms_per_step": ms,
            "cell_updates_per_s": float(got_n / (ms * 1e-3))}
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
