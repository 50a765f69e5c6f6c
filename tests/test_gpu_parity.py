"""Parity of the CUDA path, through the C ABI, on a real B200 (pytest -m gpu).

Every test compares libyalla_b200.so -- this repo's kernels -- against
  * the committed golden vectors of the reference's own sm_100a build,
  * the CPU oracle on the same seeded inputs (sizes it finishes in seconds),
  * the reference library itself when it travelled with the snapshot,
and, at the benchmark's full size, against size-independent properties
(determinism, momentum conservation, grid invariants, link-force symmetry).
Integer and index results must be bit-exact; floating point within 1e-5
relative per step on the max-norm (north_star).
"""
import os

import numpy as np
import pytest

import make_golden
from conftest import ROOT, golden
from helpers import assert_states_close, model_cases, run_case
from yalla_b200 import workloads

pytestmark = pytest.mark.gpu


def device_grid(lib, X, gs, cs):
    import torch
    n, lanes = X.shape
    d_X = torch.from_numpy(np.ascontiguousarray(X, np.float32)).cuda()
    arrays = [torch.full((n,), -7, dtype=torch.int32, device="cuda"),
              torch.full((n,), -7, dtype=torch.int32, device="cuda"),
              torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda"),
              torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda")]
    lib.grid_build(d_X.data_ptr(), n, lanes, gs, cs,
                   *[a.data_ptr() for a in arrays])
    return [a.cpu().numpy() for a in arrays]


def host_grid(lib, X, gs, cs):
    n, lanes = X.shape
    X = np.ascontiguousarray(X, np.float32)
    arrays = [np.zeros(n, np.int32), np.zeros(n, np.int32),
              np.zeros(gs ** 3, np.int32), np.zeros(gs ** 3, np.int32)]
    lib.grid_build(X.ctypes.data, n, lanes, gs, cs,
                   *[a.ctypes.data for a in arrays])
    return arrays


GRID_KEYS = ("cube_id", "point_id", "cube_start", "cube_end")


# ---- the library in use is the CUDA one -----------------------------------------
def test_product_is_the_cuda_build(product):
    import torch
    assert torch.cuda.is_available()
    assert product.build_info.startswith("yalla-b200")
    assert product.path.endswith("yalla_b200/_lib/libyalla_b200.so")


# ---- integer work: bit-exact -------------------------------------------------------
def test_nhood_table(product, oracle):
    for gs in (5, 50, 128):
        assert np.array_equal(product.nhood(gs), oracle.nhood(gs))
        assert np.array_equal(product.nhood(gs), golden("nhood")[f"gs{gs}"])


@pytest.mark.parametrize("name", sorted(make_golden.grid_cases()))
def test_grid_build_golden(product, name):
    X, gs, cs = make_golden.grid_cases()[name]
    want = golden("grid_" + name)
    for got, key in zip(device_grid(product, X, gs, cs), GRID_KEYS):
        assert np.array_equal(got, want[key]), key


@pytest.mark.parametrize("n,lanes,gs,cs", [
    (1, 3, 50, 1.0), (2, 3, 10, 1.0), (33, 4, 50, 1.0), (5000, 3, 50, 1.0),
    (5000, 5, 50, 2.0), (20000, 4, 64, 0.7), (100000, 7, 128, 1.0),
    (300000, 3, 96, 1.0)])
def test_grid_build_vs_oracle(product, oracle, n, lanes, gs, cs):
    rng = np.random.default_rng(n + lanes)
    X = np.zeros((n, lanes), dtype=np.float32)
    X[:, :3] = workloads.random_ball(n, 0.8, rng)
    X[:, 3:] = rng.random((n, lanes - 3))
    for got, want, key in zip(device_grid(product, X, gs, cs),
                              host_grid(oracle, X, gs, cs), GRID_KEYS):
        assert np.array_equal(got, want), key


def test_grid_build_crowded_cubes(product, oracle):
    # many cells per cube, all cells in a handful of cubes, ties on faces
    rng = np.random.default_rng(3)
    X = (rng.random((4000, 3)).astype(np.float32) - 0.5) * 2.0
    X[:500] = np.round(X[:500])  # exactly on cube faces
    X[500:1500] = 0.25           # a thousand cells in one spot
    for got, want, key in zip(device_grid(product, X, 8, 1.0),
                              host_grid(oracle, X, 8, 1.0), GRID_KEYS):
        assert np.array_equal(got, want), key


def test_grid_invariants_full_size(product):
    # 1M cells, the benchmark's grid: sortedness, permutation, range coverage
    n = 1_000_000
    gs = workloads.grid_size_for(n, 0.8)
    X = workloads.lattice_ball(n, 0.8, np.random.default_rng(4))
    cube_id, point_id, start, end = device_grid(product, X, gs, 1.0)
    assert np.all(np.diff(cube_id) >= 0)
    assert np.array_equal(np.sort(point_id), np.arange(n, dtype=np.int32))
    half = gs // 2
    cubes = (np.floor(X).astype(np.int64) + half) @ np.array([1, gs, gs * gs])
    assert np.array_equal(cubes[point_id], cube_id)
    same_cube = np.diff(cube_id) == 0          # stable: ascending id in a cube
    assert np.all(np.diff(point_id)[same_cube] > 0)
    occupied = start >= 0
    assert np.all(end[~occupied] == -2) and np.all(start[~occupied] == -1)
    assert int(np.sum(end[occupied] - start[occupied] + 1)) == n
    assert np.array_equal(cube_id[start[occupied]], np.nonzero(occupied)[0])


# ---- models: golden vectors, oracle, reference library ----------------------------------
@pytest.mark.parametrize("name", sorted(model_cases()))
def test_model_golden(product, name):
    case = model_cases()[name]
    want = golden("model_" + name)
    got = run_case(product, case)
    assert_states_close(got["X_out"], want["X_out"], case["steps"], name)
    assert_states_close(got["v_out"], want["v_out"], case["steps"],
                        name + " velocities", 10)
    for key in ("mes_nbs", "epi_nbs"):
        if key in want:
            assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", sorted(model_cases()))
def test_model_vs_oracle(product, oracle, name):
    case = model_cases()[name]
    got, want = run_case(product, case), run_case(oracle, case)
    factor = 4.0 if case["X"].shape[1] > 3 else 1.0
    assert_states_close(got["X_out"], want["X_out"], case["steps"], name, factor)
    for key in ("mes_nbs", "epi_nbs"):
        if key in want:
            assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("model,n,lanes,dt,steps", [
    ("relu_grid", 50000, 3, 0.1, 10), ("epithelium", 30000, 5, 0.05, 10),
    ("relu_tile", 1500, 3, 0.1, 5), ("branching", 20000, 7, 0.1, 5)])
def test_model_vs_reference_library(product, reference, model, n, lanes, dt, steps):
    rng = np.random.default_rng(n)
    X = np.zeros((n, lanes), dtype=np.float32)
    if lanes == 3:
        X[:] = workloads.lattice_ball(n, 0.8, rng)
    else:
        X[:, :5] = workloads.polarized_ball(n, 0.8, rng, lattice=True)
        X[:, 5:] = rng.random((n, lanes - 5)) * 0.2
    case = dict(model=model, X=X, dt=dt, steps=steps,
                grid_size=workloads.grid_size_for(n, 0.8), params={},
                types=workloads.shell_types(X) if lanes == 7 else None,
                links=None)
    got, want = run_case(product, case), run_case(reference, case)
    assert_states_close(got["X_out"], want["X_out"], steps, model)
    if lanes == 7:
        assert np.array_equal(got["mes_nbs"], want["mes_nbs"])
        assert np.array_equal(got["epi_nbs"], want["epi_nbs"])


def test_gabriel_solver_vs_reference_library(product, reference):
    # Gabriel_solver (reference solvers.cuh:505-644): neighbours within the
    # cut-off whose Gabriel sphere holds no third cell
    n, steps = 20000, 5
    X = workloads.lattice_ball(n, 0.8, np.random.default_rng(33))
    case = dict(model="relu_gabriel", X=X, dt=0.1, steps=steps,
                grid_size=workloads.grid_size_for(n, 0.8), params={}, types=None,
                links=None)
    got, want = run_case(product, case), run_case(reference, case)
    assert_states_close(got["X_out"], want["X_out"], steps, "gabriel")


def test_tile_and_grid_agree(product):
    # tests/test_solvers.cu:102-125 of the reference
    X = workloads.random_ball(50, 0.733333, np.random.default_rng(6))
    out = []
    for model in ("spring_tile", "spring_grid"):
        with product.sim(model, 50) as sim:
            sim.set_state(X)
            sim.step(0.1, 2)
            out.append(sim.get_state())
    assert_states_close(out[1], out[0], 2, "grid vs tile")


@pytest.mark.parametrize("model,n,dt", [
    ("springs", 800, 0.001), ("springs", 3, 0.001), ("spring_tile", 33, 0.05),
    ("relu_tile", 2048, 0.05), ("relu_tile", 2049, 0.05), ("spring_tile", 5000, 0.05)])
def test_tile_split_pairs_matches_oracle(product, oracle, model, n, dt):
    # Tile_computer::split_pairs (extension): several lanes share one cell and
    # sum in a different, fixed order -- same tolerance as every other path,
    # and the same bits from run to run.
    X = workloads.random_ball(n, 0.8, np.random.default_rng(n))
    with oracle.sim(model, n) as sim:
        sim.set_state(X)
        sim.step(dt, 5)
        want = sim.get_state()
    got = []
    for split in (1, 1, 0):  # on by default for these models; 0 = one thread per cell
        with product.sim(model, n) as sim:
            sim.set_param("split_pairs", split)
            sim.set_state(X)
            sim.step(dt, 5)
            got.append(sim.get_state())
    assert_states_close(got[0], want, 5, f"split {model} n={n}", 4.0)
    assert_states_close(got[2], want, 5, f"plain {model} n={n}", 4.0)
    assert np.array_equal(got[0], got[1])


def test_cube_size_limits_interactions(product, oracle):
    # tests/test_solvers.cu:318-336: cube_size 0.5 hides a neighbour at 0.75
    X = np.array([[0, 0, 0], [0.75, 0, 0]], dtype=np.float32)
    with product.sim("spring_grid", 2, 50, 0.5) as sim:
        sim.set_state(X)
        sim.step(0.1, 1)
        assert sim.get_state()[0, 0] == 0
    with product.sim("spring_grid", 2, 50, 1.0) as sim:
        sim.set_state(X)
        sim.step(0.1, 1)
        assert sim.get_state()[0, 0] != 0


# ---- edge cases ------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["relu_grid", "relu_tile", "epithelium"])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 257])
def test_ragged_sizes(product, oracle, model, n):
    rng = np.random.default_rng(n)
    lanes = 5 if model == "epithelium" else 3
    X = np.zeros((n, lanes), dtype=np.float32)
    if lanes == 3:
        X[:] = workloads.random_ball(n, 0.8, rng)
    else:
        X[:] = workloads.polarized_ball(n, 0.8, rng)
    case = dict(model=model, X=X, dt=0.05, steps=3, grid_size=20, params={},
                types=None, links=None)
    got, want = run_case(product, case), run_case(oracle, case)
    assert_states_close(got["X_out"], want["X_out"], 3, f"{model} n={n}", 4.0)


def test_n_smaller_than_capacity(product, oracle):
    # n_max 4096, 700 live cells: kernels must honour the device-side count
    X = workloads.lattice_ball(700, 0.8, np.random.default_rng(8))
    out = []
    for lib in (product, oracle):
        with lib.sim("relu_grid", 4096, 30, 1.0) as sim:
            sim.set_state(X)
            sim.step(0.1, 4)
            assert sim.n() == 700
            out.append(sim.get_state())
    assert_states_close(out[0], out[1], 4, "partial fill")


def test_crowded_cells_multi_round_and_list_overflow(product, oracle):
    # 3000 cells in a ball of radius 2.2: ~700 candidates per cell, far more
    # than the staging window and the per-thread list hold at once
    rng = np.random.default_rng(9)
    X = (workloads.random_ball(3000, 0.8, rng) * 0.25).astype(np.float32)
    case = dict(model="relu_grid", X=X, dt=0.001, steps=2, grid_size=20,
                params={}, types=None, links=None)
    got, want = run_case(product, case), run_case(oracle, case)
    assert_states_close(got["X_out"], want["X_out"], 2, "crowded", 4.0)


def test_cells_at_the_grid_boundary(product, oracle):
    # cells in the outermost cube layer: neighbour cubes fall outside the grid
    # (the reference reads out of bounds there; product and oracle clamp)
    gs = 8
    rng = np.random.default_rng(10)
    X = ((rng.random((600, 3)) - 0.5) * (gs - 0.01)).astype(np.float32)
    case = dict(model="relu_grid", X=X, dt=0.01, steps=2, grid_size=gs,
                params={}, types=None, links=None)
    got, want = run_case(product, case), run_case(oracle, case)
    assert_states_close(got["X_out"], want["X_out"], 2, "boundary", 4.0)


# ---- links and polarity ------------------------------------------------------------------
def test_link_forces_golden_and_oracle(product, oracle):
    import torch
    want = golden("link_forces")
    X, links = want["X_in"], want["links"]
    d_X = torch.from_numpy(X).cuda()
    d_dX = torch.zeros_like(d_X)
    d_links = torch.from_numpy(links).cuda()
    product.link_forces(d_X.data_ptr(), d_dX.data_ptr(), 3, len(X),
                        d_links.data_ptr(), len(links), 0.2)
    got = d_dX.cpu().numpy()
    assert np.max(np.abs(got - want["dX"])) < 1e-6
    assert np.max(np.abs(got.sum(axis=0))) < 1e-4  # action = reaction


def test_link_forces_hub_and_determinism(product):
    # one cell with 4000 links (a long segment), run twice: identical bits
    import torch
    rng = np.random.default_rng(12)
    n = 5000
    X = workloads.lattice_ball(n, 0.8, rng)
    links = np.zeros((6000, 2), dtype=np.int32)
    links[:4000, 0] = 17
    links[:4000, 1] = rng.integers(18, n, size=4000)
    links[4000:, 0] = rng.integers(0, n, size=2000)
    links[4000:, 1] = rng.integers(0, n, size=2000)
    d_X = torch.from_numpy(X).cuda()
    d_links = torch.from_numpy(links).cuda()
    results = []
    for _ in range(2):
        d_dX = torch.zeros_like(d_X)
        product.link_forces(d_X.data_ptr(), d_dX.data_ptr(), 3, n,
                            d_links.data_ptr(), len(links), 0.2)
        results.append(d_dX.cpu().numpy())
    assert np.array_equal(results[0], results[1])
    r = X[links[:, 0]] - X[links[:, 1]]
    live = links[:, 0] != links[:, 1]
    unit = r[live] / np.linalg.norm(r[live], axis=1, keepdims=True)
    want = np.zeros((n, 3))
    np.add.at(want, links[live, 0], -0.2 * unit)
    np.add.at(want, links[live, 1], 0.2 * unit)
    assert np.max(np.abs(results[0] - want)) < 2e-4


def test_polarity_forces(product, oracle):
    want = golden("polarity_pairs")
    for fn, key in ((product.bending_force, "bending"),
                    (product.polarization_force, "polarization")):
        got = fn(want["Xi"], want["Xj"])
        scale = np.maximum(np.abs(want[key]), 1.0)
        assert np.max(np.abs(got - want[key]) / scale) < 1e-6, key
    # reference tests/test_polarity.cu:78-94 known answer, on the device
    Xi = np.array([[0.935, 0.675, 0.649, 0.793, 0.073]], dtype=np.float32)
    Xj = np.array([[0.566, 0.809, 0.533, 0.297, 0.658]], dtype=np.float32)
    dF = product.bending_force(Xi, Xj)[0]
    assert np.allclose(dF, (0.214, -0.971, -1.802, -0.339, 0.453), rtol=1e-2)


# ---- dynamic cell count -----------------------------------------------------------------
def test_growth_statistics_match_reference(product, reference):
    # With noise the two builds are compared on ensemble statistics from the
    # same curand seeds: the cell count after K steps of division.
    rng = np.random.default_rng(14)
    X = workloads.polarized_ball(4000, 0.8, rng, lattice=True)
    types = workloads.shell_types(X)
    counts = {}
    for name, lib in (("product", product), ("reference", reference)):
        ns = []
        for seed in range(4):
            with lib.sim("growth", 16000, 40, 1.0) as sim:
                sim.set_param("prolif_rate", 0.02)
                sim.set_param("seed", seed)
                sim.set_ints("type", types)
                sim.set_state(X)
                sim.step(0.1, 20)
                ns.append(sim.n())
                state = sim.get_state()
                assert np.all(np.isfinite(state))
        counts[name] = np.array(ns, dtype=np.float64)
    assert np.all(counts["product"] > 4000)
    # same seeds, same division rule: the mean growth agrees within 3 %
    assert abs(counts["product"].mean() - counts["reference"].mean()) < \
        0.03 * counts["reference"].mean()


def test_first_division_round_is_bit_exact_vs_reference(product, reference):
    """Division counts are integer work (north_star): with the same curand seed
    the first round of divisions must pick exactly the same mothers in both
    builds -- the same number of daughters, and the same multiset of daughter
    positions (which slot a daughter lands in depends on atomic order in both
    builds, and with it the RNG stream later rounds draw from, so later rounds
    are compared statistically above)."""
    rng = np.random.default_rng(16)
    n = 6000
    X = workloads.polarized_ball(n, 0.8, rng, lattice=True, noise=0.0)
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    out = {}
    for name, lib in (("product", product), ("reference", reference)):
        with lib.sim("growth", 4 * n, 40, 1.0) as sim:
            sim.set_param("prolif_rate", 0.05)
            sim.set_param("seed", 9)
            sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(0.1, 1)
            state = sim.get_state()
            out[name] = (state, sim.get_ints("type"))
    (a, types_a), (b, types_b) = out["product"], out["reference"]
    assert len(a) == len(b) > n                      # same number of divisions
    assert np.array_equal(np.sort(types_a[n:]), np.sort(types_b[n:]))
    # mothers moved by the same step (tolerance), daughters are the same set
    assert_states_close(a[:n], b[:n], 1, "mothers after one step")
    order_a = np.lexsort(a[n:, :3].T[::-1])
    order_b = np.lexsort(b[n:, :3].T[::-1])
    assert_states_close(a[n:][order_a], b[n:][order_b], 1, "daughters (sorted)")


def test_growth_appended_cells_are_integrated(product):
    rng = np.random.default_rng(15)
    X = workloads.polarized_ball(2000, 0.8, rng, lattice=True)
    with product.sim("growth", 8000, 32, 1.0) as sim:
        sim.set_param("prolif_rate", 0.05)
        sim.set_ints("type", np.zeros(2000, dtype=np.int32))
        sim.set_state(X)
        sim.step(0.1, 10)
        n = sim.n()
        state = sim.get_state()
    assert 2000 < n <= 8000 and len(state) == n
    assert np.all(np.isfinite(state))
    # daughters are placed mean_dist / 4 = 0.1875 from their mothers
    # (passive_growth.cu:82-84) and pushed apart by the following steps: only
    # the last step's daughters may still sit that close
    from scipy.spatial import cKDTree
    nearest = cKDTree(state[:, :3]).query(state[:, :3], k=2)[0][:, 1]
    assert nearest.min() > 0.05
    assert np.median(nearest) > 0.3


# ---- full-size properties (the benchmark configuration) ----------------------------------
# ---- seeded device-side initial conditions (extension, b200/seeded_inits.cuh) ----------
def test_seeded_sphere_distribution_and_determinism(product):
    n, d = 200_000, 0.8
    states = []
    for seed in (7, 7, 8):
        with product.sim("relu_grid", n, 100, 1.0) as sim:
            sim.seed_sphere(n, d, seed)
            assert sim.n() == n
            states.append(sim.get_state())
    assert np.array_equal(states[0], states[1])       # the seed is the tissue
    assert not np.array_equal(states[0], states[2])
    X = states[0].astype(np.float64)
    radius = workloads.ball_radius(n, d)              # inits.cuh:41-48
    r = np.linalg.norm(X, axis=1)
    assert r.max() <= radius * (1 + 1e-5)
    # uniform in the ball: r^3 uniform, directions isotropic
    counts, _ = np.histogram((r / radius) ** 3, bins=10, range=(0, 1))
    assert np.all(np.abs(counts - n / 10) < 5 * np.sqrt(n / 10))
    assert np.all(np.abs(X.mean(axis=0)) < 5 * radius / np.sqrt(3 * n))
    octant = (X > 0).astype(int) @ np.array([1, 2, 4])
    assert np.all(np.abs(np.bincount(octant, minlength=8) - n / 8)
                  < 5 * np.sqrt(n / 8))


def test_seeded_sphere_prefix_is_independent_of_n(product):
    # counter-based generator keyed by (seed, cell): cell i gets the same
    # uniforms however many cells are generated; only the radius scales
    out = []
    for n in (1000, 5000):
        with product.sim("relu_grid", 5000, 50, 1.0) as sim:
            sim.seed_sphere(n, 0.8, 3)
            out.append(sim.get_state())
    scale = (1000 / 5000) ** (1.0 / 3.0)
    assert np.allclose(out[0], out[1][:1000] * scale, rtol=1e-5, atol=1e-6)


def test_relaxed_seeded_sphere(product):
    from scipy.spatial import cKDTree
    n, d = 2000, 0.75
    with product.sim("relu_grid", n, 50, 1.0) as sim:
        sim.seed_sphere(n, d, 11, relax_steps=-1)  # the reference's 2000 steps
        X = sim.get_state().astype(np.float64)
        assert sim.n() == n
    assert np.all(np.isfinite(X))
    nearest = cKDTree(X).query(X, k=2)[0][:, 1]
    # relaxed: nobody much closer than the equilibrium distance, and the mean
    # nearest-neighbour distance close to it (relu_force, inits.cuh:78-93)
    assert nearest.min() > 0.6 * d
    assert abs(np.median(nearest) - d) < 0.12 * d
    with product.sim("relu_grid", n, 50, 1.0) as sim:
        sim.seed_sphere(n, d, 11, relax_steps=-1)
        assert np.array_equal(sim.get_state().astype(np.float64), X)


def test_seeded_sphere_needs_the_product_library(oracle):
    with oracle.sim("relu_grid", 100, 20, 1.0) as sim:
        with pytest.raises(Exception):
            sim.seed_sphere(100, 0.8, 1)


@pytest.fixture(scope="module")
def million():
    n = 1_000_000
    return n, workloads.grid_size_for(n, 0.8), workloads.lattice_ball(
        n, 0.8, np.random.default_rng(16))


def test_full_size_determinism_and_momentum(product, million):
    n, gs, X = million
    runs = []
    for _ in range(2):
        with product.sim("relu_grid", n, gs, 1.0) as sim:
            sim.set_state(X)
            sim.step(0.1, 5)
            runs.append(sim.get_state())
    assert np.array_equal(runs[0], runs[1])  # no atomics-order dependence
    out = runs[0].astype(np.float64)
    assert np.all(np.isfinite(out))
    drift = np.abs(out.mean(axis=0) - X.astype(np.float64).mean(axis=0))
    assert np.all(drift < 1e-5)  # centre of mass is fixed (solvers.cuh:241-255)
    assert 1e-3 < np.max(np.abs(out - X)) < 0.5  # it did move, sanely


def test_full_size_host_roundtrip_matches_device_resident(product, million):
    n, gs, X = million
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(0.1, 2)
        resident = sim.get_state()
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_state(X)  # zero velocities, like the run above
        out = np.zeros_like(X)
        count = sim.step_host(X, 0.1, 2, out)
    assert count == n
    assert np.array_equal(out, resident)


def test_full_size_sample_vs_oracle(product, oracle):
    # 200k cells is the largest size the oracle finishes in seconds
    n = 200_000
    X = workloads.lattice_ball(n, 0.8, np.random.default_rng(17))
    case = dict(model="relu_grid", X=X, dt=0.1, steps=2,
                grid_size=workloads.grid_size_for(n, 0.8), params={},
                types=None, links=None)
    got, want = run_case(product, case), run_case(oracle, case)
    assert_states_close(got["X_out"], want["X_out"], 2, "200k cells")


def test_reproducible_division_growth(product):
    """Growth with Cell_division (b200/division.cuh) instead of the example's
    proliferate kernel: same division rule, but Philox keyed by (seed, cell,
    step) and daughters appended in mother order -- two runs give the very same
    arrays (with atomicAdd(d_n, 1) the daughters' slots differ from run to run),
    and the growth rate matches the example's kernel statistically."""
    n, steps = 40000, 12
    rng = np.random.default_rng(5)
    X = workloads.polarized_ball(n, 0.75, rng, lattice=True, noise=0.0)
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    gs = workloads.grid_size_for(n, 0.75, growth=2.0)
    runs = {}
    for mode in (1, 1, 0):
        with product.sim("growth", 2 * n, gs, 1.0) as sim:
            for key, value in (("prolif_rate", 0.02), ("mean_dist", 0.75), ("seed", 3),
                               ("reproducible_division", mode)):
                sim.set_param(key, value)
            sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(0.1, steps)
            state = sim.get_state()
            runs.setdefault(mode, []).append((state, sim.get_ints("type")))
    (a, types_a), (b, types_b) = runs[1]
    assert len(a) == len(b) > n
    assert np.array_equal(a, b) and np.array_equal(types_a, types_b)
    grown, grown_ref = len(a) - n, len(runs[0][0][0]) - n
    assert abs(grown - grown_ref) < 0.1 * grown_ref + 50, (grown, grown_ref)
    assert np.all(np.isfinite(a))


# ---- the state-carrying grid build (place_cells + settle_cells) --------------------
def test_state_carrying_build_matches_oracle(oracle, tmp_path):
    """The build tail used for float3/float4 tissues with n_max >= 4 M is chosen
    once per process; force it in a fresh interpreter and compare a small
    tissue with the oracle (positions after 3 steps, Grid and Gabriel-free
    models with and without extra lanes)."""
    import subprocess
    import sys
    script = f"""
import sys
import numpy as np
sys.path.insert(0, {ROOT!r})
import yalla_b200 as yb
from yalla_b200 import workloads
rng = np.random.default_rng(21)
product = yb.product()
for model, lanes in (("relu_grid", 3), ("epithelium", 5)):
    n = 30000
    X = np.zeros((n, lanes), dtype=np.float32)
    X[:, :5 if lanes == 5 else 3] = (workloads.polarized_ball(n, 0.8, rng)
                                     if lanes == 5 else workloads.random_ball(n, 0.8, rng))
    with product.sim(model, n, 60, 1.0) as sim:
        sim.set_state(X)
        sim.step(0.05, 3)
        np.save(sys.argv[1] + "/" + model + ".npy", sim.get_state())
    np.save(sys.argv[1] + "/" + model + "_in.npy", X)
"""
    env = dict(os.environ, YALLA_B200_CARRY_STATE="1")
    result = subprocess.run([sys.executable, "-c", script, str(tmp_path)], env=env,
                            capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stderr[-2000:]
    for model in ("relu_grid", "epithelium"):
        X = np.load(tmp_path / f"{model}_in.npy")
        got = np.load(tmp_path / f"{model}.npy")
        with oracle.sim(model, len(X), 60, 1.0) as sim:
            sim.set_state(X)
            sim.step(0.05, 3)
            want = sim.get_state()
        assert_states_close(got, want, 3, f"carried build, {model}", 4.0)
