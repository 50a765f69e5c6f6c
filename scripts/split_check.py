"""Split-pair Tile sweep against the one-thread-per-cell sweep (GPU)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import yalla_b200 as yb
from yalla_b200 import workloads
rng = np.random.default_rng(1)
lib = yb.product()
for model, n, d, dt in (("springs", 800, 0.5, 0.001), ("spring_tile", 5000, 0.8, 0.05),
                        ("spring_tile", 3000, 0.8, 0.05), ("relu_tile", 5000, 0.8, 0.05),
                        ("spring_tile", 2049, 0.8, 0.05), ("spring_tile", 16, 0.8, 0.05)):
    X = workloads.random_ball(n, d, rng)
    for steps in (1, 10, 100):
        ends = {}
        for split in (0, 1):
            with lib.sim(model, len(X), 50, 1.0) as sim:
                sim.set_param("split_pairs", split)
                sim.set_state(X); sim.step(dt, steps); sim.sync()
                ends[split] = sim.get_state()
        print(model, n, steps, "max|X|", np.abs(ends[0]).max(), np.abs(ends[1]).max(),
              "diff", np.abs(ends[1] - ends[0]).max(), flush=True)
