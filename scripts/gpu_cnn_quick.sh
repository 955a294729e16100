python -m pytest tests/test_gpu_cnn.py -x -q 2>&1 | tail -2
P=16 python scripts/bench_cnn.py 2>&1 | cut -c1-260
ncu --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:cnn_conv -s 1 -c 1 env P=8 REPS=1 python scripts/bench_cnn.py 2>&1 | grep -E "cnn_conv_kernel|gpu__time|fma_cycles|wavefronts|conflicts"
