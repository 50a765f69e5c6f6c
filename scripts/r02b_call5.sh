set -x
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 scripts/dd_branching_check.py 400000 10 10 > gpurun_out/r02b_dd_branching_n$N.log 2> gpurun_out/r02b_dd_branching_n$N.err
timeout 200 $TR --master-port 29543 scripts/dd_growth_check.py 400000 5 10 > gpurun_out/r02b_dd_growth_n$N.log 2> gpurun_out/r02b_dd_growth_n$N.err
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err
grep '^{' gpurun_out/r02b_dd_branching_n$N.log gpurun_out/r02b_dd_growth_n$N.log | cut -c1-1800
tail -c 1500 gpurun_out/r02b_bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02b_*_n$N.err | tail -20
