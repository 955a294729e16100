N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r4d_bench_n$N.log 2> gpurun_out/r4d_bench_n$N.err; tail -c 300 gpurun_out/r4d_bench_n$N.err
