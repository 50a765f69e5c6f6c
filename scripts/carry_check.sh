for c in 0 1; do
  for w in relu_1M growth_1M relu_10M branching_10M; do
    echo -n "carry=$c "; YALLA_B200_CARRY_STATE=$c python scripts/profile_step.py $w 10 product 3 | sort -t: -k2 -n | head -1
  done
done
YALLA_B200_CARRY_STATE=1 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --workload sphere_dd --steps 6 --warmup 3 2>&1 | tail -1 | cut -c1-230
