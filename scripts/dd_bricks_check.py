"""Brick decomposition across real GPUs: parity with the single-domain run.

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        scripts/dd_bricks_check.py [n_cells] [steps]

One rank per GPU; the bricks exchange halos, migrants and drift sums through
each other's memory (CUDA IPC, include/b200/domain.cuh). Rank 0 also integrates
the whole tissue in one domain and prints the max deviation.
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device(
        "cuda", int(os.environ["LOCAL_RANK"])))
    lib = yb.product()
    X = (workloads.lattice_ball(n, 0.8, np.random.default_rng(5)) * 0.9).astype(
        np.float32)
    gs = workloads.grid_size_for(n, 0.8) + 4
    bricks = dd.brick_grid_for(world)
    radius = float(np.max(np.linalg.norm(X, axis=1)))
    domain = dd.BrickDomain(lib, "relu_grid", n, gs, 1.0, bricks,
                            dd.ball_brick_cuts(radius, bricks), rank, world,
                            face_capacity=n // 2)
    domain.connect_over_ipc()
    mine = X[domain.owns(X)]
    domain.set_cells(mine)
    domain.step(0.1, steps)
    torch.cuda.synchronize()
    owned, with_ghosts, problems = domain.counts()
    state = domain.owned_state()[0].cpu().numpy()
    parts = [None] * world
    dist.all_gather_object(parts, (state, len(mine), owned, with_ghosts, problems))
    if rank == 0:
        from scipy.spatial import cKDTree
        got = np.concatenate([p[0] for p in parts])
        with lib.sim("relu_grid", n, gs, 1.0) as sim:
            sim.set_state(X)
            sim.step(0.1, steps)
            want = sim.get_state()
        distance, index = cKDTree(want).query(got, k=1)
        print(json.dumps({
            "world": world, "bricks": bricks, "cells": int(len(got)),
            "cells_expected": n, "unique_matches": int(len(np.unique(index))),
            "max_deviation": float(distance.max()),
            "migrated": int(sum(abs(p[1] - p[2]) for p in parts)),
            "ghosts": int(sum(p[3] - p[2] for p in parts)),
            "problems": int(sum(p[4] for p in parts))}))
    dist.barrier()
    domain.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
