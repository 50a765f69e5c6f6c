#!/bin/bash
# Everything profiles/ is regenerated from (run under gpurun, one GPU):
#   launch list of the bench command, full captures of the sweep kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_growth_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for w in growth_1M epithelium_1M relu_1M; do
    ncu --set full --clock-control none --import-source on -k regex:sweep_cubes \
        -s 8 -c 1 -f -o gpurun_out/sweep_$w python scripts/profile_step.py $w 2 \
        > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out/
