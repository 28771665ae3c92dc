#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_rgb_gpu.py tests/test_mlp_tc_gpu.py tests/test_mlp_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_sel.log 2>&1
echo "pytest exit $?" >> $O/pytest_sel.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_pf2.log 2>&1
I2P_TC_PREFETCH=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_pf1.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --cudnn-benchmark > $O/bench_cudnnbench.log 2>&1
timeout 300 python tools/bench_mlp.py > $O/bench_mlp_pf2.log 2>&1
I2P_TC_PREFETCH=1 timeout 300 python tools/bench_mlp.py > $O/bench_mlp_pf1.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_marked.csv python tools/profile_step.py marked $O/markers.json > $O/ncu_marked.log 2>&1
grep -E "passed|failed|FAILED|exit" $O/pytest_sel.log | tail -n 12
for f in bench_pf2 bench_pf1 bench_cudnnbench; do echo $f; tail -n 1 $O/$f.log | cut -c1-200; done
echo PF2; cut -c1-330 $O/bench_mlp_pf2.log | head -5; echo PF1; cut -c1-330 $O/bench_mlp_pf1.log | head -5
tail -3 $O/ncu_marked.log
du -sh $O
