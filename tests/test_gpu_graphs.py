"""Captured steps, cached link topology, pipelined host batches and concurrent
models (pytest -m gpu): every fast path must give the very bits of the plain one.

Reference behaviour being preserved: Heun_solver::take_step with generic forces
(solvers.cuh:226-275: zero dX, callback, pairwise sum, twice) and link_forces
(links.cuh:128-140).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from helpers import assert_states_close
from yalla_b200 import workloads

pytestmark = pytest.mark.gpu


def typed_tissue(n, seed, lanes=5):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, lanes), dtype=np.float32)
    X[:, :5] = workloads.polarized_ball(n, 0.75, rng, lattice=True, noise=0.0)
    if lanes > 5:
        X[:, 5:] = rng.random((n, lanes - 5)).astype(np.float32) * 0.2
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    return X, types


RUN_MODELS = """
import sys
import numpy as np
sys.path.insert(0, {root!r})
import yalla_b200 as yb
from yalla_b200 import workloads
sys.path.insert(0, {root!r} + "/tests")
from test_gpu_graphs import typed_tissue
lib = yb.product()
out = {{}}
X, types = typed_tissue(30000, 3)
gs = workloads.grid_size_for(30000, 0.75, growth=2.0)
with lib.sim("growth", 60000, gs, 1.0) as sim:
    sim.set_param("prolif_rate", 0.0)
    sim.set_ints("type", types)
    sim.set_state(X)
    sim.step(0.1, 6)
    out["growth"] = sim.get_state()
    out["growth_mes"] = sim.get_ints("mes_nbs")
X7, types = typed_tissue(30000, 4, lanes=7)
with lib.sim("branching", 30000, gs, 1.0) as sim:
    sim.set_ints("type", types)
    sim.set_state(X7)
    sim.step(0.1, 5)
    out["branching"] = sim.get_state()
    out["branching_epi"] = sim.get_ints("epi_nbs")
with lib.sim("branching_growth", 60000, gs, 1.0) as sim:
    sim.set_param("mes_rate", 0.0)
    sim.set_param("epi_rate", 0.0)
    sim.set_ints("type", types)
    sim.set_state(X7)
    sim.step(0.1, 5)
    out["config3"] = sim.get_state()
    out["config3_links"] = sim.get_links()
Xp = workloads.lattice_ball(20000, 0.8, np.random.default_rng(5))
links = workloads.random_links(Xp, 30000, 2.0, np.random.default_rng(6))
with lib.sim("protrusions", 20000, 40, 1.0) as sim:
    sim.set_param("link_strength", 0.2)
    sim.set_links(links)
    sim.set_state(Xp)
    sim.step(0.1, 4)
    sim.set_links(links[::-1].copy())   # new topology mid-run: cache must notice
    sim.step(0.1, 4)
    out["protrusions"] = sim.get_state()
np.savez(sys.argv[1], **out)
"""


def test_captured_steps_equal_direct_steps(tmp_path):
    """growth (whole iteration recorded as one graph), branching and protrusions
    (solver-level graphs with capturable generic forces, cached link index) vs
    the same models with YALLA_B200_NO_GRAPH=1: identical bits."""
    results = {}
    for mode, env in (("graph", {}), ("direct", {"YALLA_B200_NO_GRAPH": "1"})):
        path = str(tmp_path / f"{mode}.npz")
        run = subprocess.run(
            [sys.executable, "-c", RUN_MODELS.format(root=ROOT), path],
            env=dict(os.environ, **env), capture_output=True, text=True,
            timeout=900)
        assert run.returncode == 0, run.stderr[-3000:]
        results[mode] = np.load(path)
    for key in results["graph"].files:
        assert np.array_equal(results["graph"][key], results["direct"][key]), key


def test_protrusions_with_changing_links_match_oracle(product, oracle):
    Xp = workloads.lattice_ball(4000, 0.8, np.random.default_rng(7))
    links_a = workloads.random_links(Xp, 6000, 2.0, np.random.default_rng(8))
    links_b = workloads.random_links(Xp, 3000, 2.0, np.random.default_rng(9))
    out = []
    for lib in (product, oracle):
        with lib.sim("protrusions", 4000, 30, 1.0) as sim:
            sim.set_param("link_strength", 0.2)
            sim.set_links(links_a)
            sim.set_state(Xp)
            sim.step(0.1, 3)
            sim.set_links(links_b)
            sim.step(0.1, 3)
            sim.set_param("link_strength", 0.1)  # baked into captured launches
            sim.step(0.1, 2)
            out.append(sim.get_state())
    assert_states_close(out[0], out[1], 8, "protrusions, links changed twice", 4.0)


def test_pipelined_host_batches_equal_blocking_ones(product):
    import torch
    n, n_max = 50_000, 100_000
    X, types = typed_tissue(n, 11)
    gs = workloads.grid_size_for(n, 0.75, growth=2.0)
    host_in = torch.from_numpy(X).pin_memory().numpy()
    outs = [torch.zeros((n_max, 5)).pin_memory().numpy() for _ in range(2)]
    counts = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
    batches = 5
    with product.sim("growth", n_max, gs, 1.0) as sim:
        sim.set_param("prolif_rate", 0.01)
        sim.set_ints("type", types)
        sim.set_state(X)
        for k in range(batches):
            sim.step_host_async(host_in, 0.1, 1, outs[k % 2], n_max,
                                counts[k % 2].data_ptr())
        sim.host_drain()
        pipelined = outs[(batches - 1) % 2][:n].copy()
        n_out = int(counts[(batches - 1) % 2][0])
    assert n < n_out <= n_max
    with product.sim("growth", n_max, gs, 1.0) as sim:
        sim.set_param("prolif_rate", 0.01)
        sim.set_ints("type", types)
        sim.set_state(X)
        blocking = np.zeros((n_max, 5), dtype=np.float32)
        for k in range(batches):
            sim.step_host(host_in, 0.1, 1, blocking)
    # the cells that existed when a batch started are deterministic; daughters
    # land in slots picked by atomicAdd in either mode
    assert np.array_equal(pipelined, blocking[:n])


def test_two_typed_models_on_two_streams_do_not_mix(product):
    """growth models find their Property arrays through process-global
    __device__ pointers; two instances on two non-blocking streams, stepped
    alternately, must behave like each one alone (the rebind used to
    race with the other instance's kernels)."""
    import torch
    n = 40_000
    gs = workloads.grid_size_for(n, 0.75)
    tissues = [typed_tissue(n, seed) for seed in (21, 22)]
    alone = []
    for X, types in tissues:
        with product.sim("growth", n, gs, 1.0) as sim:
            sim.set_param("prolif_rate", 0.0)
            sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(0.1, 6)
            alone.append((sim.get_state(), sim.get_ints("mes_nbs")))
    streams = [torch.cuda.Stream() for _ in tissues]
    sims = [product.sim("growth", n, gs, 1.0) for _ in tissues]
    for sim, stream, (X, types) in zip(sims, streams, tissues):
        sim.set_stream(stream.cuda_stream)
        sim.set_param("prolif_rate", 0.0)
        sim.set_ints("type", types)
        sim.set_state(X)
    for _ in range(6):
        for sim in sims:
            sim.step(0.1, 1)
    for sim, (want_X, want_nbs) in zip(sims, alone):
        assert np.array_equal(sim.get_state(), want_X)
        assert np.array_equal(sim.get_ints("mes_nbs"), want_nbs)
        sim.close()


def test_state_copies_follow_the_models_stream(product):
    # get_state right after steps on a non-blocking stream must see their result
    import torch
    n = 200_000
    X = workloads.lattice_ball(n, 0.8, np.random.default_rng(23))
    gs = workloads.grid_size_for(n, 0.8)
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(0.1, 8)
        want = sim.get_state()
    stream = torch.cuda.Stream()
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_stream(stream.cuda_stream)
        sim.set_state(X)
        sim.step(0.1, 8)
        got = sim.get_state()
        ms, updates = sim.step_timed(0.1, 3)
        assert updates == 3 * n and ms > 0.05
    assert np.array_equal(got, want)
