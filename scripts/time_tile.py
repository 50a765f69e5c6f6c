"""Tile solver, one thread per cell vs split pairs (Tile_computer::split_pairs):
us per step, and the deviation between the two after 10 steps."""
import sys
import numpy as np
sys.path.insert(0, ".")
import yalla_b200 as yb
from yalla_b200 import workloads
rng = np.random.default_rng(1)
lib = yb.product()
for model, X, dt in (("springs", workloads.random_ball(800, 0.5, rng), 0.001),
                     ("spring_tile", workloads.random_ball(5000, 0.8, rng), 0.05)):
    ends = {}
    for split in (0, 1):
        with lib.sim(model, len(X), 50, 1.0) as sim:
            sim.set_param("split_pairs", split)
            sim.set_state(X)
            sim.step(dt, 10)
            ends[split] = sim.get_state()
            best = min(sim.step_timed(dt, 200)[0] / 200 for _ in range(3))
            print(f"{model} n={len(X)} split={split}: {best*1e3:.1f} us/step", flush=True)
    print("  max |split - plain| after 10 steps:", np.abs(ends[1] - ends[0]).max())
