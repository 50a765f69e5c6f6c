#!/bin/bash
# Everything profiles/r02_* holds from one B200 (run under gpurun, one GPU):
#   gpurun --timeout 2400 -- bash scripts/collect_round2.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gputest.log
python scripts/parity_full.py > gpurun_out/r02_parity_full.log 2>&1
python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/r02_bench_growth_1M_reference.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_growth_1M.json 2> gpurun_out/r02_bench.err
bash scripts/final_numbers.sh > gpurun_out/r02_final_numbers.log 2>&1
bash scripts/tune_variants.sh > gpurun_out/r02_tune_final.log 2>&1
tests/_bin/grid_ab_product > gpurun_out/r02_grid_build_ab.json
tests/_bin/grid_ab_reference >> gpurun_out/r02_grid_build_ab.json
bash scripts/capture_profiles.sh > gpurun_out/r02_capture.log 2>&1
tail -3 gpurun_out/r02_gputest.log
cat gpurun_out/r02_final_numbers.log
