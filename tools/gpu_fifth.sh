#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_mlp_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q --no-header -p no:cacheprovider -s -k "mlp or model or select" > gpurun_out/pytest_gpu_all.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py step > gpurun_out/ncu_step.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pw_linear_fwd_kernel<64>|pw_linear_bwd" -c 60 \
    -f -o gpurun_out/prof_mlp python tools/profile_step.py step > gpurun_out/ncu_mlp.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu_all.log | tail -n 30; tail -n 2 gpurun_out/bench_graph.log
