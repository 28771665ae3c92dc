#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/gpus.txt
timeout 300 python tools/profile_step.py kineto $O/kineto_step.md > $O/kineto.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.log 2>&1
echo "n2 exit $?" >> $O/bench_n2.log
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref.log 2>&1
head -40 $O/kineto_step.md | cut -c1-170
tail -n 3 $O/bench_n2.log | cut -c1-600
tail -n 1 $O/bench_ref.log | cut -c1-300
