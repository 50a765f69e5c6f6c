"""Parity at the BASELINE sizes: product vs the reference's own sm_100a build
(oracle/_ref/libyalla_ref.so) on a real B200, through the C ABI (pytest -m gpu).

Two kinds of comparison (scripts/parity_full.py holds the shared code):

* step by step ("resync"): before every one of the K steps both libraries are
  loaded with the reference's state of the step before -- positions,
  polarities, old velocities -- so each step is compared on identical inputs.
  This is the north_star's bar: 1e-5 relative PER STEP on the max-norm, for the
  positions and for the polarities separately (as unit vectors; as angles away
  from the coordinate poles), neighbour counters bit for bit.
* free running: K steps from the same start. The model forces are
  discontinuous at the cut-off (relu-type: F(1) = -0.2, inits.cuh:85-87), so
  once the last-bit differences of the two builds' drift sums (thrust::reduce
  vs a fixed-order sum) put a pair on different sides of the cut-off, that pair's
  cells move apart by O(0.2 dt) -- in ANY implementation that is not
  bit-identical to the reference. The free run therefore asserts the
  tolerance on all but a bounded fraction of the cells, and exactly where no
  such flip happens (10 M cells x 3-5 steps, protrusions).
"""
import numpy as np
import pytest

import parity_full
from parity_full import TOL_PER_STEP, full_cases, make_inputs

pytestmark = pytest.mark.gpu

# fraction of cells allowed outside K * 1e-5 after K free-running steps
FREE_OUTLIERS = 5e-3


@pytest.fixture(scope="module")
def inputs_of():
    cache = {}

    def get(name):
        if name not in cache:
            cache.clear()  # one 1 M / 10 M tissue at a time
            cache[name] = make_inputs(full_cases()[name])
        return cache[name]
    return get


@pytest.mark.parametrize("name", ["growth_1M", "epithelium_1M", "branching_1M",
                                  "protrusions_1M"])
def test_every_step_matches_reference(product, reference, inputs_of, name):
    spec = full_cases()[name]
    worst = parity_full.run_resync(product, reference, spec, inputs_of(name))
    n = spec["n"]
    # a pair can straddle the cut-off within ONE step too (the predictor
    # positions X1 of the two builds differ in the last bit): allow a handful
    assert worst["pos_cells_over_tol"] <= 4, worst
    assert worst["pos_err_median"] <= TOL_PER_STEP / 10, worst
    if "pol_dir_err" in worst:
        assert worst["pol_dir_cells_over_tol"] <= 4, worst
        assert worst["pol_angle_cells_over_tol"] <= 4, worst
    for key in ("mes_nbs_mismatches", "epi_nbs_mismatches"):
        if key in worst:
            assert worst[key] <= 4, worst  # counters of the flipped pairs
    if "conc_err" in worst:
        assert worst["conc_err"] <= 10 * TOL_PER_STEP, worst
    assert n == len(inputs_of(name)[0])


@pytest.mark.parametrize("name", ["growth_1M", "epithelium_1M", "branching_1M"])
def test_free_run_matches_reference_statistically(product, reference, inputs_of,
                                                  name):
    spec = full_cases()[name]
    report = parity_full.run_free(product, reference, spec, inputs_of(name))
    n = spec["n"]
    assert report["pos_err_median"] <= TOL_PER_STEP * spec["steps"] / 10, report
    assert report["pos_cells_over_tol"] <= FREE_OUTLIERS * n, report
    assert report["pol_dir_cells_over_tol"] <= FREE_OUTLIERS * n, report
    for key in ("mes_nbs_mismatches", "epi_nbs_mismatches"):
        if key in report:
            assert report[key] <= FREE_OUTLIERS * n, report


@pytest.mark.parametrize("name", ["branching_10M", "relu_10M", "protrusions_1M"])
def test_free_run_matches_reference_exactly(product, reference, inputs_of, name):
    # no cut-off flip within these runs: the plain north_star bar
    spec = full_cases()[name]
    report = parity_full.run_free(product, reference, spec, inputs_of(name))
    assert report["pos_cells_over_tol"] == 0, report
    assert report["pos_err"] <= TOL_PER_STEP * spec["steps"], report
    if "pol_dir_err" in report:
        assert report["pol_dir_cells_over_tol"] == 0, report
        assert report["pol_angle_cells_over_tol"] == 0, report
        assert report["conc_err"] <= TOL_PER_STEP * spec["steps"], report
    for key in ("mes_nbs_mismatches", "epi_nbs_mismatches"):
        if key in report:
            assert report[key] == 0, report


# ---- the public Grid at 1 M cells, bit for bit ------------------------------------
@pytest.mark.parametrize("gs", [112, 128, 256])
def test_grid_build_1M_matches_reference(product, reference, gs):
    import torch
    from yalla_b200 import workloads
    n = 1_000_000
    X = workloads.lattice_ball(n, 0.8, np.random.default_rng(gs))
    d_X = torch.from_numpy(X).cuda()
    out = []
    for lib in (product, reference):
        arrays = [torch.full((n,), -7, dtype=torch.int32, device="cuda"),
                  torch.full((n,), -7, dtype=torch.int32, device="cuda"),
                  torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda"),
                  torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda")]
        lib.grid_build(d_X.data_ptr(), n, 3, gs, 1.0,
                       *[a.data_ptr() for a in arrays])
        out.append(arrays)
    for got, want, key in zip(out[0], out[1], ("cube_id", "point_id",
                                               "cube_start", "cube_end")):
        assert torch.equal(got, want), key


def test_grid_build_many_more_cubes_than_cells(product, oracle):
    # 256^3 cubes (4096 scan tiles of look-back) for 5000 cells
    import torch
    from yalla_b200 import workloads
    n, gs = 5000, 256
    X = workloads.random_ball(n, 0.8, np.random.default_rng(5)) * 8.0
    X = X.astype(np.float32)
    d_X = torch.from_numpy(X).cuda()
    arrays = [torch.full((n,), -7, dtype=torch.int32, device="cuda"),
              torch.full((n,), -7, dtype=torch.int32, device="cuda"),
              torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda"),
              torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda")]
    for _ in range(3):  # the scan's epoch-tagged status words are reused
        product.grid_build(d_X.data_ptr(), n, 3, gs, 1.0,
                           *[a.data_ptr() for a in arrays])
    want = [np.zeros(n, np.int32), np.zeros(n, np.int32),
            np.zeros(gs ** 3, np.int32), np.zeros(gs ** 3, np.int32)]
    oracle.grid_build(X.ctypes.data, n, 3, gs, 1.0, *[a.ctypes.data for a in want])
    for got, ref in zip(arrays, want):
        assert np.array_equal(got.cpu().numpy(), ref)


# ---- with noise: ensemble statistics from the same curand seeds ---------------------
def test_growth_ensemble_statistics_match_reference(product, reference):
    """8 seeds, 30 steps of growth with division: the cell count over time, the
    radius of gyration and the histogram of mesenchymal-neighbour counts agree
    between the builds within the spread over seeds."""
    from yalla_b200 import workloads
    rng = np.random.default_rng(31)
    n0, n_max, gs = 20_000, 60_000, 60
    X = workloads.polarized_ball(n0, 0.75, rng, lattice=True, noise=0.0)
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    seeds, checkpoints = range(8), (10, 20, 30)
    stats = {}
    for name, lib in (("product", product), ("reference", reference)):
        counts, gyration, hist = [], [], []
        for seed in seeds:
            with lib.sim("growth", n_max, gs, 1.0) as sim:
                sim.set_param("prolif_rate", 0.01)
                sim.set_param("seed", seed)
                sim.set_ints("type", types)
                sim.set_state(X)
                series = []
                for _ in checkpoints:
                    sim.step(0.1, 10)
                    series.append(sim.n())
                state = sim.get_state().astype(np.float64)
                nbs = sim.get_ints("mes_nbs")
            assert np.all(np.isfinite(state))
            counts.append(series)
            centred = state[:, :3] - state[:, :3].mean(axis=0)
            gyration.append(np.sqrt((centred ** 2).sum(axis=1).mean()))
            hist.append(np.bincount(np.clip(nbs, 0, 39), minlength=40) / len(nbs))
        stats[name] = (np.array(counts, dtype=np.float64), np.array(gyration),
                       np.array(hist))
    (n_a, rg_a, h_a), (n_b, rg_b, h_b) = stats["product"], stats["reference"]
    assert np.all(n_a[:, -1] > n0)
    # n(t): ensemble means within 1 % at every checkpoint, and within three
    # standard errors of the seed-to-seed spread (floor: 0.2 %)
    spread = np.maximum(n_b.std(axis=0) / np.sqrt(len(seeds)), 0.002 * n_b.mean(axis=0))
    assert np.all(np.abs(n_a.mean(axis=0) - n_b.mean(axis=0)) < 0.01 * n_b.mean(axis=0))
    assert np.all(np.abs(n_a.mean(axis=0) - n_b.mean(axis=0)) < 3 * np.sqrt(2) * spread)
    assert abs(rg_a.mean() - rg_b.mean()) < 0.005 * rg_b.mean()
    # neighbour-count histogram: total-variation distance of the ensemble means
    assert 0.5 * np.abs(h_a.mean(axis=0) - h_b.mean(axis=0)).sum() < 0.01
