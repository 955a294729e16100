N=${1:-8}
python -m pytest tests/test_multirank.py -x -q > gpurun_out/r3p_multirank$N.log 2>&1; tail -3 gpurun_out/r3p_multirank$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r3p_bench_n$N.log 2> gpurun_out/r3p_bench_n$N.err; tail -c 400 gpurun_out/r3p_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r3p_ref_n$N.log 2> gpurun_out/r3p_ref_n$N.err; tail -c 300 gpurun_out/r3p_ref_n$N.log
