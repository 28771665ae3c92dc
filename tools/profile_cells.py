"""Launch the cell-list neighbour queries once per size (for an ncu launch list: the per-kernel split of the build and the
query).  python tools/profile_cells.py [n ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi  # noqa: E402

B = 8
dev = torch.device("cuda:0")
for n in [int(a) for a in sys.argv[1:]] or [8192, 131072]:
    m = n // 4
    g = torch.Generator(device=dev).manual_seed(n)
    xyz = (torch.rand(B, n, 3, device=dev, generator=g) * torch.tensor([80., 80., 4.], device=dev) - torch.tensor([40., 40., 3.], device=dev)).contiguous()
    q = xyz[:, torch.randperm(n, device=dev, generator=g)[:m]].contiguous()
    bidx = torch.zeros(B, m, 32, dtype=torch.int32, device=dev)
    for _ in range(2):
        _cabi.ball_query(B, n, m, 0.5, 32, q, xyz, bidx)
        d3, i3 = torch.empty(B, n, 3, device=dev), torch.empty(B, n, 3, dtype=torch.int32, device=dev)
        _cabi.three_nn(B, n, m, xyz, q, d3, i3)
        kd, ki = torch.empty(B, m, 16, device=dev), torch.empty(B, m, 16, dtype=torch.int32, device=dev)
        _cabi.knn(B, m, n, 16, q, xyz, kd, ki)
    torch.cuda.synchronize()
print("ok")
