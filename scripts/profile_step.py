"""Run a few steps of one model, for ncu:  profile_step.py <model> <n> [steps] [impl]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import yalla_b200 as yb  # noqa: E402
from yalla_b200 import workloads  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "relu_grid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
impl = sys.argv[4] if len(sys.argv) > 4 else "product"
lib = yb.product() if impl == "product" else yb.reference()
rng = np.random.default_rng(3)
gs = workloads.grid_size_for(n, 0.8)
lanes = yb.MODEL_LANES[model]
if lanes == 3:
    X = workloads.lattice_ball(n, 0.8, rng)
else:
    X = np.zeros((n, lanes), dtype=np.float32)
    X[:, :5] = workloads.polarized_ball(n, 0.8, rng, lattice=True)
dt = 0.05 if lanes > 3 else 0.1
with lib.sim(model, n, gs, 1.0) as sim:
    sim.set_state(X)
    sim.step(dt, 2)
    sim.sync()
    ms, updates = sim.step_timed(dt, steps)
    print(f"{model} n={n} gs={gs} {impl}: {ms / steps:.3f} ms/step")
