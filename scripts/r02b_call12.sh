python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gputest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_growth_1M.json 2> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_bench_growth_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-decomposed \
    > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_gputest.log; tail -c 400 gpurun_out/r02_bench_growth_1M.json
