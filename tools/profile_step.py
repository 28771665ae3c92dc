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


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "step"
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
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
