#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py step > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pw_linear_fwd_kernel<64>" -s 12 -c 2 \
    -f -o gpurun_out/prof_fwd python tools/profile_step.py step > gpurun_out/ncu_fwd.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pw_linear_bwd_dw_kernel<64, 64" -c 3 \
    -f -o gpurun_out/prof_dw python tools/profile_step.py step > gpurun_out/ncu_dw.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"select_k_kernel" -c 1 \
    -f -o gpurun_out/prof_select python tools/profile_step.py fwd > gpurun_out/ncu_select.log 2>&1
ls -la gpurun_out
