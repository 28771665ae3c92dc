#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench.log 2>&1
timeout 200 python tools/bench_infer.py > $O/bench_infer.log 2>&1
timeout 300 python tools/profile_step.py kineto $O/kineto_step.md > $O/kineto.log 2>&1
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | tail -n 12
tail -n 1 $O/bench.log | cut -c1-2600
cat $O/bench_infer.log | tail -3
head -24 $O/kineto_step.md | cut -c1-150
