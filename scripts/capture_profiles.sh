#!/bin/bash
# Everything profiles/r02_* is regenerated from (run under gpurun, one GPU):
#   launch list of the bench command, full captures of the sweep kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_bench_growth_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-decomposed \
    > gpurun_out/bench_under_ncu.log 2>&1
for k in interact_lists list_cubes; do
    ncu --set full --clock-control none --import-source on -k regex:$k \
        -s 8 -c 1 -f -o gpurun_out/r02_${k}_growth_1M python scripts/profile_step.py growth_1M 2 \
        > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:sweep_cubes \
    -s 8 -c 1 -f -o gpurun_out/r02_sweep_cubes_relu_1M python scripts/profile_step.py relu_1M 2 \
    > gpurun_out/ncu_relu.log 2>&1
for k in interact_lists list_cubes; do
    ncu --set full --clock-control none --import-source on -k regex:$k \
        -s 8 -c 1 -f -o gpurun_out/r02_${k}_epithelium_1M python scripts/profile_step.py epithelium_1M 2 \
        > gpurun_out/ncu_epi_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
