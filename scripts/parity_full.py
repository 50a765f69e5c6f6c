"""Product vs the reference's own sm_100a build at the BASELINE sizes.

    gpurun -- python scripts/parity_full.py [case ...] > gpurun_out/parity_full.log

Runs each case through libyalla_b200.so and oracle/_ref/libyalla_ref.so on the
same seeded inputs and prints one JSON line per case: max-norm errors of the
positions and of the polarities (each on its own scale), integer counters
compared bit for bit, and how many cells deviate by more than the tolerance.
tests/test_gpu_parity_full.py asserts on the same numbers (full_cases() and
compare() below are shared).
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import workloads  # noqa: E402

TOL_PER_STEP = 1e-5  # north_star: relative, per step, max-norm


def full_cases():
    """name -> dict(model, n, lanes, d, dt, steps, typed, links_per_cell, params)"""
    return {
        # configs[1] without the (noisy) division: 20 Heun steps at 1 M cells
        "growth_1M": dict(model="growth", n=1_000_000, d=0.75, dt=0.2, steps=20,
                          typed=True, params={"prolif_rate": 0.0}),
        # configs[2]
        "epithelium_1M": dict(model="epithelium", n=1_000_000, d=0.8, dt=0.05,
                              steps=20, typed=False, params={}),
        # configs[3]'s functor at 1 M and at its named size
        "branching_1M": dict(model="branching", n=1_000_000, d=0.75, dt=0.2,
                             steps=10, typed=True, params={}),
        "branching_10M": dict(model="branching", n=10_000_000, d=0.75, dt=0.2,
                              steps=3, typed=True, params={}),
        # float3 at 10 M: the state-carrying build tail (n_max >= 4 M)
        "relu_10M": dict(model="relu_grid", n=10_000_000, d=0.8, dt=0.1, steps=5,
                         typed=False, params={}),
        "protrusions_1M": dict(model="protrusions", n=1_000_000, d=0.8, dt=0.1,
                               steps=10, typed=False, links_per_cell=1,
                               params={"link_strength": 0.2}),
    }


def make_inputs(spec, seed=4242):
    rng = np.random.default_rng(seed)
    lanes = yb.MODEL_LANES[spec["model"]]
    n, d = spec["n"], spec["d"]
    if lanes == 3:
        X = workloads.lattice_ball(n, d, rng)
    else:
        X = np.zeros((n, lanes), dtype=np.float32)
        X[:, :5] = workloads.polarized_ball(
            n, d, rng, lattice=True, noise=0.0 if spec["typed"] else 0.5)
        if lanes > 5:
            X[:, 5:] = rng.random((n, lanes - 5)).astype(np.float32) * 0.2
    types = None
    if spec["typed"]:
        types = workloads.shell_types(X)
        X[types == 0, 3:5] = 0
    links = None
    if spec.get("links_per_cell"):
        links = workloads.random_links(X, spec["links_per_cell"] * n, 2.0,
                                       np.random.default_rng(77))
    gs = workloads.grid_size_for(n, d)
    return X, types, links, gs


def run(lib, spec, X, types, links, gs):
    with lib.sim(spec["model"], len(X), gs, 1.0) as sim:
        for key, value in spec["params"].items():
            sim.set_param(key, value)
        if types is not None:
            sim.set_ints("type", types)
        if links is not None:
            sim.set_links(links)
        sim.set_state(X)
        start = time.perf_counter()
        sim.step(spec["dt"], spec["steps"])
        sim.sync()
        seconds = time.perf_counter() - start
        out = {"X": sim.get_state(), "v": sim.get_velocities(),
               "ms_per_step": seconds * 1e3 / spec["steps"]}
        if types is not None:
            out["mes_nbs"] = sim.get_ints("mes_nbs")
            out["epi_nbs"] = sim.get_ints("epi_nbs")
        return out


def pol_vectors(X):
    theta, phi = X[:, 3].astype(np.float64), X[:, 4].astype(np.float64)
    return np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi),
                     np.cos(theta)], axis=1)


def compare(got, want, steps):
    """Errors of got (product) against want (reference build) after `steps`
    steps from the same state. Positions on the max-norm of the positions;
    polarities as unit vectors (scale 1) and, away from the coordinate poles
    where phi is ill-conditioned (d phi / dt ~ 1 / sin theta), as angles on the
    max-norm of the angles."""
    a, b = got["X"].astype(np.float64), want["X"].astype(np.float64)
    tol = TOL_PER_STEP * steps
    report = {}
    pos_scale = max(float(np.max(np.abs(b[:, :3]))), 1.0)
    pos_err = np.max(np.abs(a[:, :3] - b[:, :3]), axis=1) / pos_scale
    report["pos_err"] = float(pos_err.max())
    report["pos_err_median"] = float(np.median(pos_err))
    report["pos_cells_over_tol"] = int(np.sum(pos_err > tol))
    if b.shape[1] >= 5:
        dir_err = np.max(np.abs(pol_vectors(a) - pol_vectors(b)), axis=1)
        report["pol_dir_err"] = float(dir_err.max())
        report["pol_dir_cells_over_tol"] = int(np.sum(dir_err > tol))
        off_pole = np.abs(np.sin(b[:, 3])) > 0.1
        ang_scale = max(float(np.max(np.abs(b[:, 3:5]))), 1.0)
        ang_err = np.max(np.abs(a[:, 3:5] - b[:, 3:5]), axis=1) / ang_scale
        report["pol_angle_err_off_pole"] = float(ang_err[off_pole].max())
        report["pol_angle_cells_over_tol"] = int(np.sum(ang_err[off_pole] > tol))
    if b.shape[1] > 5:  # concentrations
        c_scale = max(float(np.max(np.abs(b[:, 5:]))), 1.0)
        report["conc_err"] = float(np.max(np.abs(a[:, 5:] - b[:, 5:])) / c_scale)
    for key in ("mes_nbs", "epi_nbs"):
        if key in want:
            report[key + "_mismatches"] = int(np.sum(got[key] != want[key]))
    return report


def run_free(product, reference, spec, inputs):
    """K steps in both libraries from the same start; errors at the end. Forces
    that are discontinuous at the cut-off (relu-type: F(1) = -0.2) amplify the
    last-bit differences of the two drift sums once a pair straddles the
    cut-off in one build only, so over many steps a few cells leave the
    tolerance; run_resync bounds the error of every single step instead."""
    got = run(product, spec, *inputs)
    want = run(reference, spec, *inputs)
    report = compare(got, want, spec["steps"])
    report["ms_per_step"] = {"product": got["ms_per_step"],
                             "reference": want["ms_per_step"]}
    return report


def run_resync(product, reference, spec, inputs, steps=None):
    """Step-by-step parity: before every step both libraries are loaded with the
    reference's state (positions, polarities, old velocities) of the step
    before, so each of the K steps is compared on identical inputs -- the
    north_star's "1e-5 per step on the max-norm". Returns the worst step."""
    X, types, links, gs = inputs
    steps = steps or spec["steps"]
    v = np.zeros((len(X), 3), dtype=np.float32)
    worst = {}
    with product.sim(spec["model"], len(X), gs, 1.0) as ours, \
            reference.sim(spec["model"], len(X), gs, 1.0) as theirs:
        for sim in (ours, theirs):
            for key, value in spec["params"].items():
                sim.set_param(key, value)
            if types is not None:
                sim.set_ints("type", types)
            if links is not None:
                sim.set_links(links)
        for _ in range(steps):
            out = []
            for sim in (ours, theirs):
                sim.set_state(X, reset_v=False)
                sim.set_velocities(v)
                sim.step(spec["dt"], 1)
                state = {"X": sim.get_state(), "v": sim.get_velocities()}
                if types is not None:
                    state["mes_nbs"] = sim.get_ints("mes_nbs")
                    state["epi_nbs"] = sim.get_ints("epi_nbs")
                out.append(state)
            for key, value in compare(out[0], out[1], 1).items():
                worst[key] = max(worst.get(key, 0), value)
            X, v = out[1]["X"], out[1]["v"]
    return worst


def main():
    import torch
    assert torch.cuda.is_available()
    product, reference = yb.product(), yb.reference()
    names = sys.argv[1:] or list(full_cases())
    for name in names:
        spec = full_cases()[name]
        inputs = make_inputs(spec)
        line = {"case": name, "cells": spec["n"], "steps": spec["steps"],
                "grid_size": inputs[3],
                "free": run_free(product, reference, spec, inputs)}
        if spec["n"] <= 2_000_000:
            line["resync_worst_step"] = run_resync(product, reference, spec, inputs)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
