#!/bin/bash
# Final evidence of round 2 from one B200 (run under gpurun, one GPU):
#   gpurun --timeout 1500 -- bash scripts/collect_round2_final.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gputest.log
python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/r02_bench_growth_1M_reference.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_growth_1M.json 2> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_bench_growth_1M.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-decomposed \
    > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python -m pytest tests/test_dd_bricks.py -m gpu -q -k "registered or decomposed_step" \
    > gpurun_out/r02_sanitizer_dd.log 2>&1
echo "sanitizer exit code $?" >> gpurun_out/r02_sanitizer_dd.log
tail -3 gpurun_out/r02_gputest.log; tail -5 gpurun_out/r02_sanitizer_dd.log
tail -c 600 gpurun_out/r02_bench_growth_1M_reference.json; tail -c 2500 gpurun_out/r02_bench_growth_1M.json
