#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cv_gpu.py tests/test_quat_gpu.py tests/test_rgb_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_new.log 2>&1
echo "pytest exit $?" >> $O/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_cv_gpu.py --deselect tests/test_quat_gpu.py --deselect tests/test_rgb_gpu.py --deselect tests/test_model_gpu.py > $O/pytest_rest.log 2>&1
echo "pytest exit $?" >> $O/pytest_rest.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_marked.csv python tools/profile_step.py marked $O/markers.json > $O/ncu_marked.log 2>&1
grep -E "passed|failed|FAILED|exit|Error" $O/pytest_new.log | tail -n 12
grep -E "passed|failed|FAILED|exit" $O/pytest_rest.log | tail -n 8
tail -n 1 $O/bench.log | cut -c1-260
du -sh $O
