set -x
python -m pytest tests/test_gpu_fc.py -x -q > gpurun_out/r3j_pytest_fc.log 2>&1; tail -3 gpurun_out/r3j_pytest_fc.log
python -m pytest tests/test_gpu_hmc.py -x -q > gpurun_out/r3j_pytest_hmc.log 2>&1; tail -3 gpurun_out/r3j_pytest_hmc.log
python -m pytest tests/test_multirank.py -x -q > gpurun_out/r3j_multirank2.log 2>&1; tail -3 gpurun_out/r3j_multirank2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --skip co,cpu > gpurun_out/r3j_bench_n2.log 2> gpurun_out/r3j_bench_n2.err; tail -c 600 gpurun_out/r3j_bench_n2.err
