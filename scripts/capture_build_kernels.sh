#!/bin/bash
# ncu --set full of the large-tissue build kernels and the slab compaction.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"place_cells|settle_cells" \
    -s 6 -c 2 -f -o gpurun_out/build_relu_10M python scripts/profile_step.py relu_10M 2 \
    > gpurun_out/ncu_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:slab_select \
    -s 9 -c 3 -f -o gpurun_out/slab_select python bench.py --workload sphere_dd --steps 2 \
    --warmup 3 > gpurun_out/ncu_select.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
