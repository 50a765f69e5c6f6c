#!/bin/bash
# Step times of the heavy-point workloads for every library in
# yalla_b200/_lib/variants (built with -D overrides, yalla_b200/build.py
# build_variant), and for the fused sweep (YALLA_B200_SPLIT_SWEEP=0).
WORKLOADS=${WORKLOADS:-"growth_1M epithelium_1M branching_1M"}
for w in $WORKLOADS; do
    echo -n "fused "
    YALLA_B200_SPLIT_SWEEP=0 python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
done
for lib in yalla_b200/_lib/libyalla_b200.so yalla_b200/_lib/variants/*.so; do
    for w in $WORKLOADS; do
        echo -n "$(basename $lib) "
        YALLA_B200_LIB=$PWD/$lib python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
    done
done
