#!/bin/bash
# Step times of the three 1 M-cell tuning workloads (best of 3 x 10 steps) and a
# quick parity pass over the sweep tests.
for w in relu_1M epithelium_1M growth_1M; do
    python scripts/profile_step.py $w 10 product 3 2>&1 | sort -t: -k2 -n | head -1
done
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or oracle or ragged or crowded or boundary or reference" 2>&1 | tail -3
