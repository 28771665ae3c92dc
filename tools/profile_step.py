"""One profiled training step for ncu (use with --profile-from-start off): a warm eager step
outside the capture range, then one eager step (or the select kernel alone) inside it.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py step
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from i2pnet_b200.engine import TrainStep  # noqa: E402
from i2pnet_b200.synthetic import make_pairs  # noqa: E402


def mlp_micro(dev):
    """The cost-volume-1 shared MLP (rows = 8 x 228 x 80, 262 -> 128 -> 64 -> 64) forward + backward and
    the SA1 select, alone: a small memory footprint keeps ncu's save/restore between replays cheap."""
    from i2pnet_b200.projectPN import PPBackbone_center as P
    from i2pnet_b200.projectPN.utils import FLAG_COPY, FLAG_SHIFT, StrideGrid, project_seq, select_flat
    torch.manual_seed(0)
    mods, c = [], 262
    for co in (128, 64, 64):
        mods.append(P.Conv2d(c, co, [1, 1], bn=True).to(dev))
        c = co
    x = torch.randn(8, 228, 80, 262, device=dev, requires_grad=True)
    d = make_pairs(8, seed=7)
    _, (cam,) = project_seq(d["raw_point_xyz"].to(dev), [d["lidar"].to(dev)], 64, 1800, False)
    grid = StrideGrid(8, 16, 225, 4, 8, dev)

    def body():
        out = P.run_mlp(mods, x)
        out.backward(torch.ones_like(out))
        select_flat(cam, cam, grid, [9, 15], 32, FLAG_SHIFT | FLAG_COPY, 0.75)
    body()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    body()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "step"
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if what == "mlp":
        return mlp_micro(dev)
    eng = TrainStep(8, device=dev, use_graph=False)
    eng.load({k: v.to(dev) for k, v in make_pairs(8, seed=0).items()})
    for _ in range(2):
        eng._step_body()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if what == "step":
        eng._step_body()
    else:  # forward only
        with torch.no_grad():
            x = eng.inputs
            eng.model(x["rgb"], x["lidar"], x["raw_point_xyz"], None, x["intrinsic"], None, None, None,
                      x["lidar_feats"], eng.cfg)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
