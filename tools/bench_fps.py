"""Furthest point sampling alone: us per launch and ns per sampling step for a list of (B, N, M), at the cluster size the
launcher picks or a forced one (I2P_FPS_CLUSTER, read once per process -- run one process per setting).

    for c in 0 1 2 4 8 16; do I2P_FPS_CLUSTER=$c python tools/bench_fps.py; done
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi  # noqa: E402

SHAPES = [(1, 8192, 2048), (8, 8192, 2048), (8, 2048, 1024), (8, 1024, 256), (8, 4096, 1024), (8, 16384, 4096), (8, 20480, 5120),
          (32, 8192, 2048), (8, 32768, 8192), (8, 65536, 16384)]


def main():
    dev = torch.device("cuda:0")
    _cabi.lib()
    c = os.environ.get("I2P_FPS_CLUSTER", "0")
    for b, n, m in SHAPES:
        g = torch.Generator(device=dev).manual_seed(n + b)
        xyz = torch.rand(b, n, 3, device=dev, generator=g) * 80 - 40
        temp = torch.empty(b, n, device=dev)
        idx = torch.zeros(b, m, dtype=torch.int32, device=dev)
        ts = []
        try:
            for i in range(6):
                temp.fill_(1e10)
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _cabi.furthest_point_sampling(b, n, m, xyz, temp, idx)
                e.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(e))
        except Exception as exc:   # a forced cluster size that cannot hold N
            print("cluster %s B %d N %d M %d: %s" % (c, b, n, m, str(exc)[:80]), flush=True)
            continue
        ms = statistics.median(ts[1:])
        print("cluster %s B %d N %d M %d: %.1f us, %.0f ns / step, checksum %d" % (c, b, n, m, ms * 1e3, ms * 1e6 / m, int(idx.long().sum())),
              flush=True)


if __name__ == "__main__":
    main()
