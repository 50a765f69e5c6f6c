set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_dd_bricks.py -m gpu -q -k branching 2>&1 | tail -15 > gpurun_out/r02b_t_branch.log
timeout 200 $TR --master-port 29517 scripts/dd_bricks_check.py 400000 5 > gpurun_out/r02b_dd_check_n2.log 2> gpurun_out/r02b_dd_check_n2.err
timeout 300 $TR --master-port 29519 scripts/dd_growth_check.py 200000 5 10 > gpurun_out/r02b_dd_growth_n2.log 2> gpurun_out/r02b_dd_growth_n2.err
timeout 300 $TR --master-port 29521 scripts/dd_branching_check.py 200000 10 10 > gpurun_out/r02b_dd_branching_n2.log 2> gpurun_out/r02b_dd_branching_n2.err
timeout 300 $TR --master-port 29523 bench.py --gpus 2 --workload sphere_dd --steps 10 --warmup 3 > gpurun_out/r02b_sdd_n2_overlap.json 2> gpurun_out/r02b_sdd_n2_overlap.err
YALLA_B200_DD_OVERLAP=0 timeout 300 $TR --master-port 29525 bench.py --gpus 2 --workload sphere_dd --steps 10 --warmup 3 > gpurun_out/r02b_sdd_n2_serial.json 2> gpurun_out/r02b_sdd_n2_serial.err
tail -3 gpurun_out/r02b_t_branch.log; cat gpurun_out/r02b_dd_check_n2.log gpurun_out/r02b_dd_growth_n2.log gpurun_out/r02b_dd_branching_n2.log; tail -3 gpurun_out/*n2*.err | tail -30
