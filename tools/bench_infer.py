"""Inference latency of the iterative-refinement model (src/modellearn_proj_center_iter.py: coarse pose + six
level-3 refinements), the shape of the reference's evaluation loop (evaluation_proj.py:239-264): batch 1 and 8,
eval-free (training-mode statistics as in the parity tests), forward only, CUDA-graph replay, CUDA events."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg  # noqa: E402
from i2pnet_b200.modellearn_proj_center_iter import RegNet_v2  # noqa: E402
from i2pnet_b200.synthetic import make_pairs  # noqa: E402


def small_range(dev):
    """BASELINE.json configs[0] shapes on the GPU: the small-range model (src/modellearn.py; 160x512 image + 8192 points
    -> 2048 / 1024 / 256 / 64), forward only."""
    from i2pnet_b200.config_lidarcenter import I2PNetConfig as scfg
    from i2pnet_b200.modellearn import RegNet_v2 as SmallNet
    from i2pnet_b200.synthetic import make_pairs_small
    torch.manual_seed(0)
    model = SmallNet(cfg=scfg).to(dev).train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    for batch in (1, 8):
        d = {k: v.to(dev) for k, v in make_pairs_small(batch, seed=3).items()}
        run = lambda d=d: model(d["rgb"], d["lidar"], None, d["intrinsic"], None, None, None, None, cfg=scfg, lidar_img_raw=d["raw_point_xyz"])
        yield "small-range model (8192 points)", batch, run


def iterative(dev):
    torch.manual_seed(0)
    model = RegNet_v2(cfg=cfg).to(dev).train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    for batch in (1, 8):
        d = {k: v.to(dev) for k, v in make_pairs(batch, seed=3).items()}
        run = lambda d=d: model(d["rgb"], d["lidar"], d["raw_point_xyz"], None, d["intrinsic"], None, None, None, d["lidar_feats"], cfg)
        yield "iter model (6 refinements)", batch, run


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name, batch, run in list(iterative(dev)) + list(small_range(dev)):
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    run()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run()
            ts = []
            for _ in range(30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); graph.replay(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        ms = statistics.median(ts[5:])
        print("%s batch %d: %.2f ms per forward = %.1f pairs/s; out_3[0] = %s" % (
            name, batch, ms, batch / ms * 1e3, [round(float(v), 4) for v in out[0][0]]), flush=True)


if __name__ == "__main__":
    main()
