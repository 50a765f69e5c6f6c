for w in growth_1M epithelium_1M relu_1M branching_1M protrusions_1M gabriel_1M; do
  python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
done > gpurun_out/r02b_selfpair.log
cat gpurun_out/r02b_selfpair.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
