#!/bin/bash
# round-1 ground truth: GPU parity tests, bench (both MLP paths), ncu launch list + full captures
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.log 2>&1
I2P_MLP_TC=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_fma.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_step.csv python tools/profile_step.py step > $O/ncu_step.log 2>&1
for k in pw_linear_fwd_tc_kernel pw_linear_bwd_dx_kernel pw_linear_bwd_dw_kernel select_k_kernel; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$k" -c 40 \
      -f -o $O/full_$k python tools/profile_step.py step > $O/ncu_full_$k.log 2>&1
  ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/full_$k.raw.csv 2>/dev/null
done
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | tail -n 12
tail -n 1 $O/bench.log | cut -c1-1500; tail -n 1 $O/bench_fma.log | cut -c1-300; tail -n 1 $O/bench_ref.log | cut -c1-400
ls -la $O
