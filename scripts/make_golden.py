"""Generate tests/golden/*.npz from the UNMODIFIED reference build on a B200.

    gpurun -- python scripts/make_golden.py        (writes gpurun_out/golden/)
    cp gpurun_out/golden/*.npz tests/golden/

Every fixture holds seeded inputs and what oracle/_ref/libyalla_ref.so (the
reference's own headers compiled for sm_100a, see oracle/Makefile) produced for
them. tests/test_oracle.py pins the CPU oracle against these files, and
tests/test_gpu_parity.py the product. The inputs are regenerated from the seeds
in `cases()` by the tests as well, so a fixture cannot silently drift from the
generator.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import workloads  # noqa: E402


def cases():
    """name -> dict(model, X, dt, steps, grid_size, params, types, links)"""
    out = {}
    rng = np.random.default_rng(20261017)

    def add(name, model, X, dt, steps, gs=50, params=None, types=None,
            links=None):
        out[name] = dict(model=model, X=np.ascontiguousarray(X, np.float32),
                         dt=dt, steps=steps, grid_size=gs, params=params or {},
                         types=types, links=links)

    add("springs", "springs", workloads.random_ball(200, 0.5, rng), 0.001, 5)
    X = workloads.random_ball(300, 0.7333, rng)
    add("spring_tile", "spring_tile", X, 0.1, 5)
    add("spring_grid", "spring_grid", X, 0.1, 5)
    add("relu_tile", "relu_tile", workloads.lattice_ball(400, 0.8, rng), 0.1, 5)
    add("relu_grid", "relu_grid", workloads.lattice_ball(3000, 0.8, rng), 0.1, 5,
        gs=30)
    add("relu_grid_random", "relu_grid", workloads.random_ball(3000, 0.8, rng),
        0.05, 3, gs=30)
    add("relu_grid_fix_point", "relu_grid",
        workloads.lattice_ball(1000, 0.8, rng), 0.1, 4, gs=24,
        params={"fix_point": 13})
    add("relu_grid_fix_xy", "relu_grid", workloads.lattice_ball(1000, 0.8, rng),
        0.1, 4, gs=24, params={"fix_point_xy": 7})
    X5 = workloads.polarized_ball(2000, 0.8, rng, lattice=True)
    add("epithelium", "epithelium", X5, 0.05, 5, gs=30)
    types = workloads.shell_types(X5)
    add("growth_static", "growth", X5, 0.1, 3, gs=30,
        params={"prolif_rate": 0.0}, types=types)
    Xp = workloads.lattice_ball(1500, 0.8, rng)
    add("protrusions", "protrusions", Xp, 0.1, 5, gs=30,
        params={"link_strength": 0.2},
        links=workloads.random_links(Xp, 1500, 2.0, rng))
    X7 = np.zeros((2000, 7), dtype=np.float32)
    X7[:, :5] = X5
    X7[:, 5:] = rng.random((2000, 2)).astype(np.float32) * 0.2
    add("branching", "branching", X7, 0.1, 3, gs=30, types=types)
    return out


def run_case(lib, case):
    X = case["X"]
    n_max = len(X)
    with lib.sim(case["model"], n_max, case["grid_size"], 1.0) as sim:
        for key, value in case["params"].items():
            sim.set_param(key, value)
        if case["types"] is not None:
            sim.set_ints("type", case["types"])
        if case["links"] is not None:
            sim.set_links(case["links"])
        sim.set_state(X)
        sim.step(case["dt"], case["steps"])
        result = {"X_out": sim.get_state(), "v_out": sim.get_velocities()}
        if case["types"] is not None:
            result["mes_nbs"] = sim.get_ints("mes_nbs")
            result["epi_nbs"] = sim.get_ints("epi_nbs")
        return result


def grid_cases():
    rng = np.random.default_rng(7)
    out = {}
    # the 7^3 lattice of tests/test_solvers.cu:287-315, cube sizes 1 and 2
    k = np.arange(7, dtype=np.float32) + 0.5
    gz, gy, gx = np.meshgrid(k, k, k, indexing="ij")
    lattice = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    out["lattice_cs1"] = (lattice, 70, 1.0)
    out["lattice_cs2"] = (lattice, 70, 2.0)
    out["ball3"] = (workloads.random_ball(2000, 0.8, rng), 24, 1.0)
    X5 = np.zeros((1500, 5), dtype=np.float32)
    X5[:, :3] = workloads.random_ball(1500, 0.8, rng)
    out["ball5_cs07"] = (X5, 32, 0.7)
    return out


def polarity_pairs():
    rng = np.random.default_rng(11)
    Xi = rng.random((256, 5)).astype(np.float32)
    Xj = rng.random((256, 5)).astype(np.float32)
    Xi[:, 3] *= np.pi
    Xj[:, 3] *= np.pi
    Xi[:, 4] = (Xi[:, 4] * 2 - 1) * np.pi
    Xj[:, 4] = (Xj[:, 4] * 2 - 1) * np.pi
    return Xi, Xj


def main():
    import torch
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    os.makedirs(out_dir, exist_ok=True)
    ref = yb.reference()
    print(ref.build_info)

    for name, case in cases().items():
        result = run_case(ref, case)
        np.savez_compressed(os.path.join(out_dir, f"model_{name}.npz"),
                            X_in=case["X"], **result)
        print("model", name, result["X_out"].shape)

    for name, (X, gs, cs) in grid_cases().items():
        n, lanes = X.shape
        d_X = torch.from_numpy(X).cuda()
        arrays = [torch.zeros(n, dtype=torch.int32, device="cuda"),
                  torch.zeros(n, dtype=torch.int32, device="cuda"),
                  torch.zeros(gs ** 3, dtype=torch.int32, device="cuda"),
                  torch.zeros(gs ** 3, dtype=torch.int32, device="cuda")]
        ref.grid_build(d_X.data_ptr(), n, lanes, gs, cs,
                       *[a.data_ptr() for a in arrays])
        cube_id, point_id, start, end = [a.cpu().numpy() for a in arrays]
        np.savez_compressed(os.path.join(out_dir, f"grid_{name}.npz"), X_in=X,
                            grid_size=gs, cube_size=cs, cube_id=cube_id,
                            point_id=point_id, cube_start=start, cube_end=end)
        print("grid", name, n)

    np.savez_compressed(os.path.join(out_dir, "nhood.npz"),
                        **{f"gs{gs}": ref.nhood(gs) for gs in (5, 50, 128)})

    Xi, Xj = polarity_pairs()
    np.savez_compressed(os.path.join(out_dir, "polarity_pairs.npz"), Xi=Xi, Xj=Xj,
                        bending=ref.bending_force(Xi, Xj),
                        polarization=ref.polarization_force(Xi, Xj))

    rng = np.random.default_rng(13)
    X = workloads.lattice_ball(1000, 0.8, rng)
    links = workloads.random_links(X, 1500, 2.0, rng)
    d_X = torch.from_numpy(X).cuda()
    d_dX = torch.zeros_like(d_X)
    d_links = torch.from_numpy(links).cuda()
    ref.link_forces(d_X.data_ptr(), d_dX.data_ptr(), 3, len(X),
                    d_links.data_ptr(), len(links), 0.2)
    np.savez_compressed(os.path.join(out_dir, "link_forces.npz"), X_in=X,
                        links=links, strength=0.2, dX=d_dX.cpu().numpy())
    print("done ->", out_dir)


if __name__ == "__main__":
    main()
