#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
I2P_TC_BDB=1 timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_mlp_gpu.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_bdb.log 2>&1
echo "pytest exit $?" >> $O/pytest_bdb.log
timeout 200 python tools/bench_mlp.py > $O/bench_mlp_bdb0.log 2>&1
I2P_TC_BDB=1 timeout 200 python tools/bench_mlp.py > $O/bench_mlp_bdb1.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_bdb0.log 2>&1
I2P_TC_BDB=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_bdb1.log 2>&1
grep -E "passed|failed|FAILED|exit" $O/pytest_bdb.log | tail -n 6
echo BDB0; cut -c1-330 $O/bench_mlp_bdb0.log | head -6; echo BDB1; cut -c1-330 $O/bench_mlp_bdb1.log | head -6
for f in bench_bdb0 bench_bdb1; do tail -n 1 $O/$f.log | cut -c1-180; done
