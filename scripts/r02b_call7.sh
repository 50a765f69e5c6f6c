set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_upstream_gpu.py -m gpu -q -k "gabriel or upstream" 2>&1 | tail -8 > gpurun_out/r02b_t_gabriel.log
cat gpurun_out/r02b_t_gabriel.log
python scripts/profile_step.py gabriel_1M 10 product 3 > gpurun_out/r02b_gabriel_ab.log 2>&1
YALLA_B200_GABRIEL_LISTS=0 python scripts/profile_step.py gabriel_1M 10 product 3 >> gpurun_out/r02b_gabriel_ab.log 2>&1
python scripts/profile_step.py gabriel_1M 10 reference 3 >> gpurun_out/r02b_gabriel_ab.log 2>&1
cat gpurun_out/r02b_gabriel_ab.log
