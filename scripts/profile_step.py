"""Run a few steps of one bench workload, for ncu:
    profile_step.py <workload> [steps] [impl]"""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
import yalla_b200 as yb  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "growth_1M"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
impl = sys.argv[3] if len(sys.argv) > 3 else "product"
lib = yb.product() if impl == "product" else yb.reference()
spec = bench.WORKLOADS[workload]
X, types, gs = bench.make_state(spec, seed=1000)
with bench.new_sim(lib, spec, X, types, gs) as sim:
    sim.step(spec["dt"], 3)
    sim.sync()
    repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    for _ in range(repeats):
        ms, updates = sim.step_timed(spec["dt"], steps)
        print(f"{workload} gs={gs} {impl}: {ms / steps:.3f} ms/step, "
              f"{updates / ms / 1e6:.3f} G cell-updates/s", flush=True)
