set -x
timeout 400 python -m pytest tests/test_dd_bricks.py tests/test_gpu_parity.py tests/test_upstream_gpu.py -m gpu -q -k "misuse or gabriel or upstream" 2>&1 | tail -8 > gpurun_out/r02b_t_gabriel2.log
cat gpurun_out/r02b_t_gabriel2.log
python scripts/profile_step.py gabriel_1M 10 product 3 > gpurun_out/r02b_gabriel_ab2.log 2>&1
YALLA_B200_GABRIEL_LISTS=0 python scripts/profile_step.py gabriel_1M 10 product 3 >> gpurun_out/r02b_gabriel_ab2.log 2>&1
cat gpurun_out/r02b_gabriel_ab2.log
