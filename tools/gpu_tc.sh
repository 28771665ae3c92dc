#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_tc.log 2>&1
tail -n 25 gpurun_out/pytest_tc.log | cut -c1-300
I2P_MLP_TC=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1
tail -n 1 gpurun_out/bench_tc.log | cut -c1-300
