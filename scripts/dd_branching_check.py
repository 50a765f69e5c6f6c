"""BASELINE.json configs[3] -- branching cell + division + one protrusion per cell
rewired every step -- cut into bricks across real GPUs.

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
        scripts/dd_branching_check.py [n_cells] [steps] [timed_steps]

Links are kept as cell identities (include/b200/brick_links.cuh) and resolved per
stage over owned + ghost cells; a link across a cut pulls on its `a` end on the
brick that owns `a` and on its `b` end on the brick that owns `b`. Curand draws
are per brick, so the comparison with one domain is statistical: cell count,
radius of gyration, linked fraction, link lengths. Exact: every link end names a
live cell, and (rates 0, link strength 0) positions equal the one-domain run.
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
HALO = 2.5  # protrusions reach r_protrusion = 2 cube sizes


def tissue(n, seed):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 7), dtype=np.float32)
    X[:, :5] = workloads.polarized_ball(n, 0.75, rng, lattice=True, noise=0.0)
    X[:, 5:] = rng.random((n, 2)).astype(np.float32) * 0.2
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    return X, types


def set_params(sim, rate, strength):
    for name, value in (("seed", 9), ("mes_rate", rate), ("epi_rate", rate),
                        ("link_strength", strength)):
        sim.set_param(name, value)


def bricks_run(lib, X, types, n_max, gs, rate, strength, steps, rank, world, timed=0):
    bricks = dd.brick_grid_for(world)
    radius = float(np.max(np.linalg.norm(X[:, :3], axis=1)))
    domain = dd.BrickDomain(lib, "branching_growth", n_max, gs, 1.0, bricks,
                            dd.ball_brick_cuts(radius, bricks), rank, world,
                            face_capacity=max(n_max // 3, 4096), halo=HALO)
    domain.connect_over_ipc()
    set_params(domain.sim, rate, strength)
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
    part = (domain.owned_state()[0].cpu().numpy(), domain.sim.get_ints("identity"),
            domain.sim.get_ints("partner"),
            int(domain.sim.get_ints("unresolved_links")[0]), int(mine.sum()), owned,
            with_ghosts, problems)
    parts = [None] * world
    dist.all_gather_object(parts, part)
    dist.barrier()
    domain.close()
    return parts, ms


def one_domain(lib, X, types, n_max, gs, rate, strength, steps):
    with lib.sim("branching_growth", n_max, gs, 1.0) as sim:
        set_params(sim, rate, strength)
        sim.set_ints("type", types)
        sim.set_state(X)
        sim.step(DT, steps)
        return sim.get_state(), sim.get_links()


def link_statistics(X, a, b):
    live = a != b
    length = np.linalg.norm(X[a[live], :3] - X[b[live], :3], axis=1)
    centred = X[:, :3] - X[:, :3].mean(axis=0)
    return {"cells": int(len(X)), "linked_fraction": float(live.mean()),
            "mean_link_length": float(length.mean()) if live.any() else 0.0,
            "max_link_length": float(length.max()) if live.any() else 0.0,
            "radius_of_gyration": float(np.sqrt((centred ** 2).sum(axis=1).mean()))}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    timed = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(
        "cuda", int(os.environ["LOCAL_RANK"])))
    lib = yb.product()
    X, types = tissue(n, 61)
    X[:, :3] *= 0.95
    gs = workloads.grid_size_for(n, 0.75, growth=2.5)
    report = {"world": world, "bricks": dd.brick_grid_for(world), "cells": n,
              "halo": HALO}

    # 1. exact: no division, links that do not pull -> the positions of one domain
    parts, _ = bricks_run(lib, X, types, n, gs, 0.0, 0.0, 5, rank, world)
    if rank == 0:
        from scipy.spatial import cKDTree
        got = np.concatenate([p[0] for p in parts])
        want, _ = one_domain(lib, X, types, n, gs, 0.0, 0.0, 5)
        distance, index = cKDTree(want[:, :3]).query(got[:, :3], k=1)
        error = np.abs(got - want[index]).max(axis=1)
        identity = np.concatenate([p[1] for p in parts])
        report["no_pull"] = {
            "cells": int(len(got)), "unique_matches": int(len(np.unique(index))),
            "max_deviation": float(error.max()),
            "cells_beyond_1e-4": int(np.sum(error > 1e-4)),
            "unique_identities": int(len(np.unique(identity))),
            "migrated": int(sum(abs(p[4] - p[5]) for p in parts)),
            "problems": int(sum(p[7] for p in parts))}

    # 2. + 3. the model as it is: statistics against one domain, and speed
    total = steps + timed
    parts, ms = bricks_run(lib, X, types, 3 * n, gs, 0.01, 0.2, steps, rank, world,
                           timed=timed)
    if rank == 0:
        got = np.concatenate([p[0] for p in parts])
        identity = np.concatenate([p[1] for p in parts])
        partner = np.concatenate([p[2] for p in parts])
        order = np.argsort(identity)
        at = np.searchsorted(identity[order], partner)
        at = np.clip(at, 0, len(identity) - 1)
        dangling = identity[order][at] != partner
        ends = order[at]
        here = np.arange(len(identity))
        ends[dangling] = here[dangling]
        mine = link_statistics(got, here, ends)
        want, links = one_domain(lib, X, types, 3 * n, gs, 0.01, 0.2, total)
        theirs = link_statistics(want, links[:len(want), 0], links[:len(want), 1])
        report["model"] = {
            "steps": total, "bricks": mine, "one_domain": theirs,
            "unique_identities": bool(len(np.unique(identity)) == len(identity)),
            "dangling_links": int(dangling.sum()),
            "unresolved_link_stages": int(sum(p[3] for p in parts)),
            "ghosts": int(sum(p[6] - p[5] for p in parts)),
            "finite": bool(np.all(np.isfinite(got))),
            "problems": int(sum(p[7] for p in parts)),
            "ms_per_step": ms,
            "cell_updates_per_s": float(len(got) / (ms * 1e-3))}
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
