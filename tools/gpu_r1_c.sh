#!/bin/bash
# validate the RGB block tail + tensor-core MLP family, then time them
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_mlp_tc_gpu.py tests/test_rgb_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > $O/pytest_new.log 2>&1
echo "pytest exit $?" >> $O/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_mlp_tc_gpu.py --deselect tests/test_rgb_gpu.py > $O/pytest_rest.log 2>&1
echo "pytest exit $?" >> $O/pytest_rest.log
timeout 300 python tools/bench_mlp.py > $O/bench_mlp.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_tc7.log 2>&1
I2P_MLP_TC=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_tc0.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_mlp_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "r300 or pack" > $O/sanitizer.log 2>&1
grep -E "passed|failed|FAILED|exit|Error" $O/pytest_new.log | tail -n 30
grep -E "passed|failed|FAILED|exit" $O/pytest_rest.log | tail -n 12
cat $O/bench_mlp.log | cut -c1-400
tail -n 1 $O/bench_tc7.log | cut -c1-250; tail -n 1 $O/bench_tc0.log | cut -c1-250
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/sanitizer.log | head
du -sh $O
