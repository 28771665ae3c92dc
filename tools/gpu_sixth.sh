#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --channels-last > gpurun_out/bench_graph_cl.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"select_k_kernel" -c 1 \
    -f -o gpurun_out/prof_select python tools/profile_step.py fwd > gpurun_out/ncu_select.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pw_linear_fwd_kernel" -s 12 -c 1 \
    -f -o gpurun_out/prof_fwd python tools/profile_step.py fwd > gpurun_out/ncu_fwd.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu_all.log | tail -n 10; tail -n 1 gpurun_out/bench_graph.log gpurun_out/bench_graph_cl.log | cut -c1-400
