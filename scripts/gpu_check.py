"""Development harness: A/B the product against the reference build on a GPU box.

    gpurun -- python scripts/gpu_check.py [quick|timing|all]

Prints one line per check; not part of the test-suite (tests/ holds the real
parity tests), just the fastest way to see what a fresh kernel does.
"""
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import workloads  # noqa: E402


def report(name, ok, detail=""):
    print(f"[{'ok' if ok else 'FAIL'}] {name} {detail}", flush=True)


def check_nhood(new, ref):
    for gs in (5, 50, 128):
        a, b = new.nhood(gs), ref.nhood(gs)
        report(f"nhood gs={gs}", np.array_equal(a, b), f"{a[:6]}...")


def grid_arrays(lib, X, gs, cs):
    n, lanes = X.shape
    d_X = torch.from_numpy(X).cuda()
    cube_id = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    point_id = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    start = torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda")
    end = torch.full((gs ** 3,), -7, dtype=torch.int32, device="cuda")
    lib.grid_build(d_X.data_ptr(), n, lanes, gs, cs, cube_id.data_ptr(),
                   point_id.data_ptr(), start.data_ptr(), end.data_ptr())
    return [t.cpu().numpy() for t in (cube_id, point_id, start, end)]


def check_grid(new, ref):
    rng = np.random.default_rng(1)
    for n, lanes, gs, cs in ((343, 3, 70, 1.0), (5000, 3, 50, 1.0),
                             (5000, 5, 50, 2.0), (100000, 7, 128, 1.0),
                             (20000, 4, 64, 0.7)):
        X = np.zeros((n, lanes), dtype=np.float32)
        X[:, :3] = workloads.random_ball(n, 0.8, rng)
        a = grid_arrays(new, X, gs, cs)
        b = grid_arrays(ref, X, gs, cs)
        same = [np.array_equal(x, y) for x, y in zip(a, b)]
        report(f"grid n={n} lanes={lanes} gs={gs} cs={cs}", all(same), str(same))


def run_model(lib, model, X, dt, steps, gs=50, setup=None):
    with lib.sim(model, len(X) if model != "growth" else 2 * len(X), gs, 1.0) as sim:
        if setup:
            setup(sim)
        sim.set_state(X)
        sim.step(dt, steps)
        out = sim.get_state()
        extra = {}
        if model in ("growth", "branching"):
            extra["mes_nbs"] = sim.get_ints("mes_nbs")
            extra["epi_nbs"] = sim.get_ints("epi_nbs")
        return out, extra


def check_models(new, ref):
    rng = np.random.default_rng(2)
    cases = []
    cases.append(("springs", workloads.random_ball(800, 0.5, rng), 0.001, 5, 50, None))
    cases.append(("spring_tile", workloads.random_ball(50, 0.7333, rng), 0.1, 5, 50, None))
    cases.append(("spring_grid", workloads.random_ball(50, 0.7333, rng), 0.1, 5, 50, None))
    cases.append(("relu_tile", workloads.random_ball(700, 0.8, rng), 0.1, 5, 50, None))
    cases.append(("relu_grid", workloads.random_ball(20000, 0.8, rng), 0.1, 5, 50, None))
    X5 = workloads.polarized_ball(5000, 0.8, rng)
    cases.append(("epithelium", X5, 0.05, 5, 50, None))
    types = (np.linalg.norm(X5[:, :3], axis=1) > 0.8 * np.max(
        np.linalg.norm(X5[:, :3], axis=1))).astype(np.int32)

    def growth_setup(sim):
        sim.set_param("prolif_rate", 0.0)
        sim.set_ints("type", types)
    cases.append(("growth", X5, 0.1, 3, 50, growth_setup))
    Xp = workloads.random_ball(3000, 0.8, rng)
    links = workloads.random_links(Xp, 3000, 2.0, rng)

    def prot_setup(sim):
        sim.set_links(links)
    cases.append(("protrusions", Xp, 0.1, 5, 50, prot_setup))
    X7 = np.zeros((5000, 7), dtype=np.float32)
    X7[:, :5] = X5
    X7[:, 5:] = rng.random((5000, 2)).astype(np.float32) * 0.2

    def branching_setup(sim):
        sim.set_ints("type", types)
    cases.append(("branching", X7, 0.1, 3, 50, branching_setup))

    for model, X, dt, steps, gs, setup in cases:
        try:
            a, ea = run_model(new, model, X, dt, steps, gs, setup)
            b, eb = run_model(ref, model, X, dt, steps, gs, setup)
            err = np.max(np.abs(a - b))
            scale = np.max(np.abs(b))
            moved = np.max(np.abs(b - X))
            ok = err <= 1e-5 * steps * max(scale, 1.0)
            detail = f"max|d|={err:.3e} scale={scale:.3g} moved={moved:.3g}"
            for key in ea:
                same = np.array_equal(ea[key], eb[key])
                detail += f" {key}:{'same' if same else 'DIFF'}"
                ok = ok and same
            report(f"model {model} n={len(X)} steps={steps}", ok, detail)
        except Exception:
            report(f"model {model}", False, traceback.format_exc())


def time_model(lib, model, X, dt, steps, gs, warmup=3, setup=None):
    n_max = len(X) if model != "growth" else 2 * len(X)
    with lib.sim(model, n_max, gs, 1.0) as sim:
        if setup:
            setup(sim)
        sim.set_state(X)
        sim.step(dt, warmup)
        sim.sync()
        ms, updates = sim.step_timed(dt, steps)
        return ms / steps, updates / (ms * 1e-3)


def check_timing(new, ref, sizes=(100_000, 1_000_000)):
    rng = np.random.default_rng(3)
    for n in sizes:
        gs = workloads.grid_size_for(n, 0.8)
        X3 = workloads.lattice_ball(n, 0.8, rng)
        X5 = workloads.polarized_ball(n, 0.8, rng, lattice=True)
        for model, X, dt in (("relu_grid", X3, 0.1), ("epithelium", X5, 0.05)):
            for name, lib in (("new", new), ("ref", ref)):
                try:
                    t0 = time.time()
                    ms, rate = time_model(lib, model, X, dt, 10, gs)
                    report(f"time {model} n={n} gs={gs} {name}", True,
                           f"{ms:.3f} ms/step {rate / 1e9:.3f} G updates/s "
                           f"(wall {time.time() - t0:.1f}s)")
                except Exception:
                    report(f"time {model} n={n} {name}", False,
                           traceback.format_exc())


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    new, ref = yb.product(), yb.reference()
    print(new.build_info, "|", ref.build_info, "|", torch.cuda.get_device_name(0))
    if what in ("quick", "all"):
        check_nhood(new, ref)
        check_grid(new, ref)
        check_models(new, ref)
    if what in ("timing", "all"):
        check_timing(new, ref)


if __name__ == "__main__":
    main()
