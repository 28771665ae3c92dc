#!/bin/bash
# full GPU suite, bench, launch list and full captures of the hot kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.log 2>&1
timeout 300 python tools/bench_mlp.py > $O/bench_mlp.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_step.csv python tools/profile_step.py step > $O/ncu_step.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"fwd_kernel|dx_kernel|dw_kernel|select_k" \
    -f -o $O/full_mlp python tools/profile_step.py mlp > $O/ncu_full_mlp.log 2>&1
ncu -i $O/full_mlp.ncu-rep --page raw --csv > $O/full_mlp.raw.csv 2>/dev/null
ls -la $O/full_mlp.ncu-rep
[ $(stat -c %s $O/full_mlp.ncu-rep) -gt 40000000 ] && rm -f $O/full_mlp.ncu-rep
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | tail -n 15
tail -n 1 $O/bench.log | cut -c1-2500
cat $O/bench_mlp.log | cut -c1-330
du -sh $O
