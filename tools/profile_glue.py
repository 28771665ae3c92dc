"""Which Python lines launch the library (ATen) kernels that remain in the training step: an eager step under the torch
profiler with Python stacks; device time and launch counts of the non-own kernels grouped by the innermost frame inside
i2pnet_b200/.   python tools/profile_glue.py [out.md]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from i2pnet_b200.engine import TrainStep  # noqa: E402
from i2pnet_b200.synthetic import make_pairs  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda:0")
    eng = TrainStep(8, device=dev, use_graph=False)
    eng.load({k: v.to(dev) for k, v in make_pairs(8, seed=0).items()})
    for _ in range(2):
        eng._step_body()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        eng._step_body()
        torch.cuda.synchronize()
    rows = collections.defaultdict(lambda: [0.0, 0, collections.Counter()])
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels or any(c.kernels for c in e.cpu_children):
            continue                              # innermost events that own kernels (parents repeat their children's)
        kern = [k for k in e.kernels if "i2p::" not in k.name]
        if not kern:
            continue
        op, stack, p = None, None, e
        while p is not None:                      # nearest operator above the launch, and the nearest Python stack
            if op is None and (p.name.startswith("aten::") or "Backward" in p.name):
                op = p.name
            if stack is None and p.stack:
                stack = p.stack
            p = p.cpu_parent
        frame = "?"
        for f in stack or []:
            if "i2pnet_b200/" in f and "engine.py" not in f:
                frame = f.split("i2pnet_b200/")[-1]
                break
        r = rows[frame]
        r[0] += sum(k.duration for k in kern)
        r[1] += len(kern)
        r[2][op or e.name] += len(kern)
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/glue.md"
    total_t, total_n = sum(r[0] for r in rows.values()), sum(r[1] for r in rows.values())
    with open(out, "w") as fh:
        fh.write("# library (ATen) kernels of one eager training step by the Python line that launches them\n\n")
        fh.write("%d launches, %.0f us of device time (backward kernels are attributed to the autograd node, frame '?')\n\n" % (total_n, total_t))
        fh.write("| us | launches | where | operators |\n|---:|---:|---|---|\n")
        for frame, r in sorted(rows.items(), key=lambda kv: -kv[1][0]):
            fh.write("| %.1f | %d | `%s` | %s |\n" % (r[0], r[1], frame[:90], ", ".join("%s x%d" % kv for kv in r[2].most_common(6))))
    print("wrote", out, total_n, "launches", total_t, "us")


if __name__ == "__main__":
    main()
