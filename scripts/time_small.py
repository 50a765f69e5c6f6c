"""Time the small configurations (C1 springs n=800 Tile; small grids): ms/step, both builds."""
import sys
import numpy as np
sys.path.insert(0, ".")
import yalla_b200 as yb
from yalla_b200 import workloads

rng = np.random.default_rng(1)
cases = [("springs", workloads.random_ball(800, 0.5, rng), 0.001, 50),
         ("spring_tile", workloads.random_ball(5000, 0.8, rng), 0.05, 50),
         ("relu_grid", workloads.lattice_ball(5000, 0.8, rng), 0.1, 30),
         ("relu_grid", workloads.lattice_ball(100000, 0.8, rng), 0.1, 48),
         ("epithelium", workloads.polarized_ball(250, 0.8, rng), 0.05, 50)]
for name, lib in (("product", yb.product()), ("reference", yb.reference())):
    for model, X, dt, gs in cases:
        with lib.sim(model, len(X), gs, 1.0) as sim:
            sim.set_state(X)
            sim.step(dt, 20)
            sim.sync()
            best = 1e9
            for _ in range(3):
                ms, _ = sim.step_timed(dt, 200)
                best = min(best, ms / 200)
            print(f"{name:9s} {model:12s} n={len(X):6d}: {best * 1e3:8.1f} us/step", flush=True)
