"""One launch of every PointNet++ operator kernel at a sweep size, inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off --set full` (BASELINE configs[4]; SURVEY.md section 8 d2: per-kernel counters).

    ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_ops_full python tools/profile_ops.py [N]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.sweep_ops import Ours, ops_for  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(n)
ops = ops_for(Ours, dev, n, n // 4, gen)
for _, fn, _, _ in ops:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _, fn, _, _ in ops:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
