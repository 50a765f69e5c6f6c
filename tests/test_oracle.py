"""Pin the CPU oracle (oracle/yalla_oracle.cpp) before anything trusts it.

Three kinds of pins, all CPU-only:
  1. the known-answer vectors and integer-exact expectations the reference's
     own tests hold (tests/test_polarity.cu, tests/test_solvers.cu,
     tests/test_links.cu in /root/reference);
  2. the behavioural properties those tests assert (relaxation to the spring
     length, momentum conservation, Tile == Grid, fixed point stays put);
  3. golden outputs of the reference's own sm_100a build for seeded inputs
     (tests/golden/*.npz, made on a B200 by scripts/make_golden.py).
"""
import numpy as np
import pytest

from conftest import golden
from helpers import assert_states_close, model_cases, run_case
import make_golden


def isclose(a, b):
    """minunit.cuh:37 of the reference."""
    return abs(a - b) <= 1e-6 + 1e-2 * abs(b)


# ---- 1. known answers --------------------------------------------------------
def test_bending_force_known_answer(oracle):
    # tests/test_polarity.cu:78-94
    Xi = np.array([[0.935, 0.675, 0.649, 0.793, 0.073]], dtype=np.float32)
    Xj = np.array([[0.566, 0.809, 0.533, 0.297, 0.658]], dtype=np.float32)
    dF = oracle.bending_force(Xi, Xj)[0]
    for got, want in zip(dF, (0.214, -0.971, -1.802, -0.339, 0.453)):
        assert isclose(got, want), (dF, want)


def test_polarization_force_known_answer(oracle):
    # tests/test_polarity.cu:20-34
    Xi = np.array([[0.601, 0.305, 0.320, 0.209, 0.295]], dtype=np.float32)
    Xj = np.array([[0.762, 0.403, 0.121, 0.340, 0.431]], dtype=np.float32)
    dF = oracle.polarization_force(Xi, Xj)[0]
    for got, want in zip(dF, (0, 0, 0, 0.126, 0.215)):
        assert isclose(got, want), (dF, want)


def test_nhood_order(oracle):
    # solvers.cuh:472-484: x fastest (-1, 0, 1), then 0, -gs, +gs, then z
    gs = 50
    want = []
    for dz in (0, -1, 1):
        for dy in (0, -1, 1):
            for dx in (-1, 0, 1):
                want.append(dx + dy * gs + dz * gs * gs)
    assert oracle.nhood(gs).tolist() == want
    assert np.array_equal(oracle.nhood(gs), golden("nhood")["gs50"])


def grid_of(lib, X, gs, cs):
    n, lanes = X.shape
    cube_id = np.zeros(n, dtype=np.int32)
    point_id = np.zeros(n, dtype=np.int32)
    start = np.zeros(gs ** 3, dtype=np.int32)
    end = np.zeros(gs ** 3, dtype=np.int32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    lib.grid_build(X.ctypes.data, n, lanes, gs, cs, cube_id.ctypes.data,
                   point_id.ctypes.data, start.ctypes.data, end.ctypes.data)
    return cube_id, point_id, start, end


def test_grid_spacing_lattice(oracle):
    # tests/test_solvers.cu:247-315: 7^3 points at half-integers, grid 70
    X, gs, _ = make_golden.grid_cases()["lattice_cs1"]
    origin = gs ** 3 // 2 + gs ** 2 // 2 + gs // 2
    cube_id, point_id, start, end = grid_of(oracle, X, gs, 1.0)
    for i in range(len(X)):
        x, y, z = i % 7, (i // 7) % 7, i // 49
        expected = origin + x + gs * y + gs * gs * z
        assert start[expected] == end[expected]  # one point per cube
        assert cube_id[i] == expected            # hence no reordering
    cube_id, point_id, start, end = grid_of(oracle, X, gs, 2.0)
    for i in range(len(X)):
        x, y, z = i % 7, (i // 7) % 7, i // 49
        expected = origin + x // 2 + gs * (y // 2) + gs * gs * (z // 2)
        members = point_id[start[expected]:end[expected] + 1]
        assert i in members


@pytest.mark.parametrize("name", sorted(make_golden.grid_cases()))
def test_grid_matches_reference_build(oracle, name):
    X, gs, cs = make_golden.grid_cases()[name]
    want = golden("grid_" + name)
    assert np.array_equal(X, want["X_in"])
    got = grid_of(oracle, X, gs, cs)
    for array, key in zip(got, ("cube_id", "point_id", "cube_start", "cube_end")):
        assert np.array_equal(array, want[key]), key


# ---- 2. behavioural pins --------------------------------------------------------
def centre_of_mass(X):
    return X[:, :3].astype(np.float64).mean(axis=0)


@pytest.mark.parametrize("model", ["spring_tile", "spring_grid"])
def test_tetrahedron_relaxes(oracle, model):
    # tests/test_solvers.cu:55-98
    rng = np.random.default_rng(5)
    X = (rng.random((4, 3)).astype(np.float32) - 0.5) * 0.5
    with oracle.sim(model, 4) as sim:
        sim.set_state(X)
        sim.step(0.1, 500)
        out = sim.get_state()
    for i in range(1, 4):
        assert isclose(np.linalg.norm(out[0] - out[i]), 0.5)
    assert np.allclose(centre_of_mass(X), centre_of_mass(out), atol=1e-5)


def test_tile_and_grid_agree(oracle):
    # tests/test_solvers.cu:102-125
    from yalla_b200 import workloads
    X = workloads.random_ball(50, 0.733333, np.random.default_rng(6))
    results = []
    for model in ("spring_tile", "spring_grid"):
        with oracle.sim(model, 50) as sim:
            sim.set_state(X)
            sim.step(0.1, 2)
            results.append(sim.get_state())
    assert_states_close(results[1], results[0], 2, "grid vs tile")


def test_fixed_point_stays(oracle):
    # tests/test_solvers.cu:228-244
    from yalla_b200 import workloads
    X = workloads.random_ball(100, 0.733333, np.random.default_rng(7))
    X[13] = 0
    with oracle.sim("spring_tile", 100) as sim:
        sim.set_param("fix_point", 13)
        sim.set_state(X)
        sim.step(0.1, 1)
        out = sim.get_state()
    assert np.all(np.abs(out[13]) < 1e-6)


def test_square_of_four_links(oracle):
    # tests/test_links.cu:15-50: only links act (cells start 2 apart)
    X = np.array([[1, 1, 0], [1, -1, 0], [-1, -1, 0], [-1, 1, 0]],
                 dtype=np.float32)
    links = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.int32)
    with oracle.sim("protrusions", 4) as sim:
        sim.set_links(links)
        sim.set_state(X)
        sim.step(0.1, 100)
        out = sim.get_state()
    assert np.allclose(centre_of_mass(out), 0, atol=1e-5)
    assert np.linalg.norm(out[0] - out[1]) < 2.0 - 0.5  # pulled together
    assert isclose(np.linalg.norm(out[0] - out[1]), np.linalg.norm(out[1] - out[2]))


def test_link_force_value(oracle):
    # links.cuh:99-111: -/+ strength * r / |r|
    X = np.array([[0, 0, 0], [3, 4, 0]], dtype=np.float32)
    dX = np.zeros_like(X)
    links = np.array([[0, 1]], dtype=np.int32)
    oracle.link_forces(X.ctypes.data, dX.ctypes.data, 3, 2, links.ctypes.data, 1,
                       0.2)
    assert np.allclose(dX[0], [0.12, 0.16, 0], atol=1e-7)
    assert np.allclose(dX[1], [-0.12, -0.16, 0], atol=1e-7)


def test_growth_needs_curand(oracle):
    import yalla_b200 as yb
    with oracle.sim("growth", 16) as sim:
        sim.set_state(np.zeros((8, 5), dtype=np.float32))
        with pytest.raises(yb.YallaError):
            sim.step(0.1, 1)  # proliferation on: not available on the CPU


# ---- 3. golden vectors from the reference's own build ------------------------------
@pytest.mark.parametrize("name", sorted(model_cases()))
def test_model_matches_reference_build(oracle, name):
    case = model_cases()[name]
    want = golden("model_" + name)
    assert np.array_equal(case["X"], want["X_in"]), "generator drifted"
    got = run_case(oracle, case)
    # libm vs CUDA transcendentals and no FMA on the host: allow 4x the
    # per-step budget for the polarity models
    factor = 4.0 if case["X"].shape[1] > 3 else 1.0
    assert_states_close(got["X_out"], want["X_out"], case["steps"], name, factor)
    assert_states_close(got["v_out"], want["v_out"], case["steps"],
                        name + " velocities", 10 * factor)
    for key in ("mes_nbs", "epi_nbs"):
        if key in want:
            assert np.array_equal(got[key], want[key]), key  # integers: exact


def test_polarity_pairs_match_reference_build(oracle):
    want = golden("polarity_pairs")
    got = oracle.bending_force(want["Xi"], want["Xj"])
    scale = np.maximum(np.abs(want["bending"]), 1.0)
    assert np.max(np.abs(got - want["bending"]) / scale) < 2e-5
    got = oracle.polarization_force(want["Xi"], want["Xj"])
    scale = np.maximum(np.abs(want["polarization"]), 1.0)
    assert np.max(np.abs(got - want["polarization"]) / scale) < 2e-5


def test_link_forces_match_reference_build(oracle):
    want = golden("link_forces")
    X = np.ascontiguousarray(want["X_in"])
    links = np.ascontiguousarray(want["links"])
    dX = np.zeros_like(X)
    oracle.link_forces(X.ctypes.data, dX.ctypes.data, 3, len(X),
                       links.ctypes.data, len(links), float(want["strength"]))
    assert np.max(np.abs(dX - want["dX"])) < 1e-6
