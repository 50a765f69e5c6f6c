"""Domain decomposition (yalla_b200/dd.py): the N > 1 path.

CPU: two gloo ranks drive the CPU oracle through exactly the orchestration code
the GPUs use (halo exchange, global drift all-reduce, migration) and must
reproduce the single-domain result. GPU (-m gpu): the same with the product
library on one device, world size 1 and, when launched under torchrun with
several GPUs (scripts/dd_check.py), across devices.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import yalla_b200 as yb
from yalla_b200 import dd, workloads
from conftest import ORACLE_LIB


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def match_cells(got, want, tol):
    """Every cell of `got` has exactly one partner in `want` within tol."""
    from scipy.spatial import cKDTree
    assert got.shape == want.shape
    distance, index = cKDTree(want[:, :3]).query(got[:, :3], k=1)
    assert len(np.unique(index)) == len(want), "cells lost or duplicated"
    assert distance.max() < tol, f"max deviation {distance.max():.3e}"
    return np.max(np.abs(got - want[index]))


def single_domain(lib, model, X, dt, steps, gs):
    with lib.sim(model, len(X), gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(dt, steps)
        return sim.get_state()


def run_rank(rank, world, port, model, X, cuts, dt, steps, gs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = yb.load(ORACLE_LIB)
        bounds = [-np.inf] + list(cuts) + [np.inf]
        z_lo, z_hi = bounds[rank], bounds[rank + 1]
        mine = X[(X[:, 2] >= z_lo) & (X[:, 2] < z_hi)]
        domain = dd.SlabDomain(lib, model, len(X), gs, 1.0, z_lo, z_hi, "cpu")
        domain.set_cells(mine)
        migrated, before = 0, len(mine)
        for _ in range(steps):
            domain.step(dt)
            now = domain.n_owned
            migrated += abs(now - before)
            before = now
        assert domain.total_cells() == len(X)
        assert domain.counts()[2] == 0
        result = domain.gather_all()
        if rank == 0:
            np.save(out, result)
            np.save(out + ".migrated.npy", np.array([migrated]))
        domain.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,lanes,world", [("relu_grid", 3, 2),
                                               ("relu_grid", 3, 3),
                                               ("epithelium", 5, 2)])
def test_two_slabs_match_single_domain_cpu(oracle, tmp_path, model, lanes, world):
    rng = np.random.default_rng(21)
    n, steps, dt = 3000, 4, 0.1 if lanes == 3 else 0.05
    if lanes == 3:
        X = workloads.lattice_ball(n, 0.8, rng)
    else:
        X = workloads.polarized_ball(n, 0.8, rng, lattice=True)
    gs = 30
    radius = np.abs(X[:, 2]).max()
    cuts = dd.ball_slab_cuts(radius, world)
    want = single_domain(oracle, model, X, dt, steps, gs)
    out = str(tmp_path / "result.npy")
    mp.spawn(run_rank, args=(world, free_port(), model, X, cuts, dt, steps, gs, out),
             nprocs=world, join=True)
    got = np.load(out)
    error = match_cells(got, want, tol=1e-4)
    assert error < 1e-5 * steps * max(np.abs(want).max(), 1.0)


def test_cells_migrate_between_slabs_cpu(oracle, tmp_path):
    # a cut through a tissue that is still relaxing: cells cross it
    rng = np.random.default_rng(22)
    X = workloads.random_ball(2000, 0.8, rng)
    want = single_domain(oracle, "relu_grid", X, 0.1, 6, 30)
    out = str(tmp_path / "result.npy")
    mp.spawn(run_rank, args=(2, free_port(), "relu_grid", X, [0.0], 0.1, 6, 30, out),
             nprocs=2, join=True)
    match_cells(np.load(out), want, tol=5e-4)


def test_slab_helpers():
    cuts = dd.ball_slab_cuts(100.0, 8)
    assert len(cuts) == 7 and cuts == sorted(cuts)
    assert abs(cuts[3]) < 1e-9 and all(c == round(c) for c in cuts)
    # equal volumes: cap fractions
    R = 100.0
    edges = np.array([-R] + cuts + [R])
    vol = (R ** 2 * edges - edges ** 3 / 3)
    fractions = np.diff(vol) / (4 * R ** 3 / 3)
    assert np.all(np.abs(fractions - 1 / 8) < 0.01)
    rng = np.random.default_rng(1)
    parts = [dd.lattice_ball_slab(12.0, 0.8, lo, hi, rng)
             for lo, hi in ((-np.inf, -3.0), (-3.0, 4.0), (4.0, np.inf))]
    whole = np.concatenate(parts)
    assert np.all(parts[1][:, 2] >= -3.0) and np.all(parts[1][:, 2] < 4.0)
    expected = 4 / 3 * np.pi * 12.0 ** 3 * np.sqrt(2) / 0.8 ** 3
    assert abs(len(whole) - expected) < 0.03 * expected
    assert np.linalg.norm(whole, axis=1).max() < 12.0 + 0.1


@pytest.mark.gpu
def test_single_slab_matches_plain_step_gpu(product):
    # world size 1: the decomposition building blocks alone must reproduce
    # take_step (same kernels, externally supplied drift)
    rng = np.random.default_rng(23)
    X = workloads.lattice_ball(50000, 0.8, rng)
    gs = workloads.grid_size_for(len(X), 0.8)
    want = single_domain(product, "relu_grid", X, 0.1, 5, gs)
    domain = dd.SlabDomain(product, "relu_grid", len(X), gs, 1.0, -np.inf, np.inf,
                           "cuda")
    domain.set_cells(X)
    for _ in range(5):
        domain.step(0.1)
    got = domain.owned_state()[0].cpu().numpy()
    domain.close()
    # (migration rounds re-store the cells in cube order: compare as sets)
    assert match_cells(got, want, 1e-4) < 1e-5 * 5 * np.abs(want).max()


@pytest.mark.gpu
def test_ghosts_are_partners_only_gpu(product, oracle):
    # one device, two solvers: emulate two slabs by hand and compare with the
    # oracle doing the same
    rng = np.random.default_rng(24)
    X = workloads.lattice_ball(20000, 0.8, rng)
    gs = workloads.grid_size_for(len(X), 0.8)
    lower, upper = X[X[:, 2] < 0], X[X[:, 2] >= 0]
    ghosts = upper[upper[:, 2] < 1.5]
    sums = {}
    for name, lib, dev in (("product", product, "cuda"), ("oracle", oracle, "cpu")):
        Xo = torch.from_numpy(lower).to(dev)
        vo = torch.zeros_like(Xo)
        Xg = torch.from_numpy(ghosts).to(dev)
        vg = torch.zeros_like(Xg)
        out = torch.zeros(4, device=dev)
        with lib.sim("relu_grid", len(X), gs, 1.0) as sim:
            sim.dd_load(0, Xo.data_ptr(), vo.data_ptr(), len(Xo), Xg.data_ptr(),
                        vg.data_ptr(), len(Xg))
            sim.dd_forces(0, out.data_ptr())
            mean = (out[:3] / out[3]).contiguous()
            sim.dd_update(0, 0.1, mean.data_ptr())
            X1 = torch.empty_like(Xo)
            sim.dd_read(1, X1.data_ptr(), len(Xo))
            if dev == "cuda":
                torch.cuda.synchronize()
            sums[name] = (out.cpu().numpy().copy(), X1.cpu().numpy().copy())
    assert sums["product"][0][3] == len(lower)
    assert np.allclose(sums["product"][0], sums["oracle"][0], rtol=1e-4, atol=1e-3)
    assert np.max(np.abs(sums["product"][1] - sums["oracle"][1])) < 1e-5 * 30


def step_slabs_in_process(domains, dt):
    """One decomposed Heun step of several slabs living in ONE process (same
    device): SlabDomain.step with the neighbour exchange and the all-reduce
    replaced by tensor copies -- every kernel of the multi-GPU path runs."""
    def exchange():
        for below, above in zip(domains[:-1], domains[1:]):
            above.recv[0].copy_(below.send[1])
            below.recv[1].copy_(above.send[0])

    for what in (0, 1, 2):
        for d in domains:
            d.sim.slab_pack(what, d.send[0].data_ptr(), d.send[1].data_ptr())
        exchange()
        for d in domains:
            d.sim.slab_unpack(what, d.recv[0].data_ptr(), d.recv[1].data_ptr())
        if what == 2:
            break
        for d in domains:
            d.sim.dd_forces(what, d.sums.data_ptr())
        total = torch.stack([d.sums for d in domains]).sum(dim=0)
        for d in domains:
            d.sums.copy_(total)
            d.sim.slab_update(what, dt, d.sums.data_ptr())


@pytest.mark.gpu
@pytest.mark.parametrize("model,lanes", [("relu_grid", 3), ("epithelium", 5)])
def test_three_slabs_on_one_device_gpu(product, model, lanes):
    # all decomposition kernels of the product (slab_select for halo and
    # migration, ghost append, merge) for both record widths, against the
    # single-domain product run
    rng = np.random.default_rng(31)
    n, steps, dt = 60000, 6, 0.05
    if lanes == 3:
        X = workloads.lattice_ball(n, 0.8, rng)
    else:
        X = workloads.polarized_ball(n, 0.8, rng, lattice=True)
    gs = workloads.grid_size_for(n, 0.8)
    want = single_domain(product, model, X, dt, steps, gs)
    bounds = [-np.inf, -4.0, 3.0, np.inf]
    domains = []
    for z_lo, z_hi in zip(bounds[:-1], bounds[1:]):
        d = dd.SlabDomain(product, model, n, gs, 1.0, z_lo, z_hi, "cuda")
        d.set_cells(X[(X[:, 2] >= z_lo) & (X[:, 2] < z_hi)])
        domains.append(d)
    start = [d.n_owned for d in domains]
    for _ in range(steps):
        step_slabs_in_process(domains, dt)
    parts = [d.owned_state()[0].cpu().numpy() for d in domains]
    problems = [d.counts()[2] for d in domains]
    for d in domains:
        d.close()
    assert problems == [0, 0, 0]
    got = np.concatenate(parts)
    assert len(got) == n and sum(start) == n
    for part, (z_lo, z_hi) in zip(parts, zip(bounds[:-1], bounds[1:])):
        assert np.all(part[:, 2] >= z_lo) and np.all(part[:, 2] < z_hi)  # migrated
    deviation = match_cells(got, want, 1e-4)
    assert deviation < 1e-5 * steps * max(np.abs(want).max(), 1.0) * 4
