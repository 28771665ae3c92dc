#!/bin/bash
# First GPU pass of a round: record the reference kernels' golden vectors, run the GPU parity
# suite, the smoke, a bench in both modes, and the ncu launch list of one eager step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python tests/golden/make_ref_gpu_golden.py > gpurun_out/ref_golden.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_graph.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py step > gpurun_out/ncu_step.log 2>&1
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_eager.log gpurun_out/bench_graph.log
