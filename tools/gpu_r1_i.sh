#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_mlp_gpu.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_sw.log 2>&1
echo "pytest exit $?" >> $O/pytest_sw.log
timeout 200 python tools/bench_mlp.py > $O/bench_mlp.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_m7.log 2>&1
I2P_MLP_TC=23 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_m23.log 2>&1
grep -E "passed|failed|FAILED|exit" $O/pytest_sw.log | tail -n 25
cut -c1-370 $O/bench_mlp.log | head -12
for f in bench_m7 bench_m23; do tail -n 1 $O/$f.log | cut -c1-180; done
