set -x
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29547 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err
tail -c 1200 gpurun_out/r02b_bench_n$N.json
