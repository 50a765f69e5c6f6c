#!/bin/bash
# ms/step of every bench workload, product and the reference's own build
# (best of 3 x 10 steps each; the reference's timings scatter, see
# profiles/r01_sweep_tuning.md), plus the small configurations.
for w in growth_1M branching_growth_1M epithelium_1M relu_1M protrusions_1M branching_1M relu_10M branching_10M branching_growth_10M; do
    for impl in product reference; do
        python scripts/profile_step.py $w 10 $impl 3 2>&1 | sort -t: -k2 -n | head -1
    done
done
python scripts/time_small.py 2>&1 | tail -12
python scripts/time_tile.py 2>&1 | tail -6
