for w in growth_1M epithelium_1M relu_1M branching_1M protrusions_1M; do
  python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
done > gpurun_out/r02b_divisor.log
cat gpurun_out/r02b_divisor.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_upstream_gpu.py tests/test_gpu_parity_full.py -m gpu -q 2>&1 | tail -12
