#!/bin/bash
# Step times of the heavy-point workloads for every library in
# yalla_b200/_lib/variants (built with -DYB_SWEEP_HEAVY_* overrides).
for lib in yalla_b200/_lib/libyalla_b200.so yalla_b200/_lib/variants/*.so; do
    for w in growth_1M epithelium_1M branching_1M; do
        echo -n "$(basename $lib) "
        YALLA_B200_LIB=$PWD/$lib python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
    done
done
