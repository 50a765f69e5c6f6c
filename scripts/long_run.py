"""Stress run: growth_1M until the tissue is full (n = n_max) and beyond.
Prints the cell count, step time and sanity checks every 50 steps."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import yalla_b200 as yb

spec = bench.WORKLOADS["growth_1M"]
lib = yb.product()
X, types, gs = bench.make_state(spec, seed=1000)
with bench.new_sim(lib, spec, X, types, gs) as sim:
    for block in range(8):
        ms, updates = sim.step_timed(spec["dt"], 50)
        state = sim.get_state()
        r = np.linalg.norm(state[:, :3], axis=1)
        print(f"steps {50 * (block + 1):4d}: n = {len(state):8d}, {ms / 50:.3f} ms/step, "
              f"finite = {bool(np.all(np.isfinite(state)))}, r_max = {r.max():.1f} "
              f"(grid half width {gs / 2})", flush=True)
