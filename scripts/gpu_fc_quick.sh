python -m pytest tests/test_gpu_fc.py tests/test_gpu_cnn.py -x -q 2>&1 | tail -2
P=64 python scripts/bench_fc.py 2>&1 | cut -c1-420
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fc_gemm3 -s 3 -c 3 env P=16 REPS=1 python scripts/bench_fc.py 2>&1 | grep -E "fc_gemm3_kernel|gpu__time_duration|tensor_cycles"
