set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
tail -c 3000 gpurun_out/r02b_bench_n2.json; tail -5 gpurun_out/r02b_bench_n2.err
