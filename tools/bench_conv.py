"""Isolated timings of the RGB pyramid's convolution kernels (csrc/conv.cu) at the batch-8 KITTI shapes: forward,
data gradient, weight gradient per distinct layer shape, L2 flushed, CUDA events; algorithmic bytes / flops beside.

    python tools/bench_conv.py [batch]
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
PROFILE = len(sys.argv) > 1 and sys.argv[1] == "profile"     # one launch per kernel and shape, for ncu --set full
B = int(sys.argv[1]) if len(sys.argv) > 1 and not PROFILE else 8
L = _cabi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# (cin, cout, H, W, how many layers of the pyramid have this shape)
SHAPES = [(3, 16, 160, 512, 1), (16, 16, 80, 256, 3), (16, 32, 80, 256, 1), (32, 32, 40, 128, 1), (32, 32, 20, 64, 3),
          (32, 64, 20, 64, 1), (64, 64, 10, 32, 4), (64, 128, 10, 32, 1)]


def timed(fn, reps=20, skip=4):
    if PROFILE:
        flush.fill_(1)
        fn()
        torch.cuda.synchronize()
        return 1.0
    evs = []
    for i in range(reps):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs[skip:]) * 1e3


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("| cin -> cout @ HxW (x layers) | fwd us (GB/s) | dgrad us | wgrad us (TFLOP/s f32) |")
print("|---|---:|---:|---:|")
for cin, cout, H, W, n in (SHAPES[:3] + SHAPES[6:7] if PROFILE else SHAPES):
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.1
    bias = torch.randn(cout, device=dev)
    y = torch.empty(B, cout, H, W, device=dev)
    dy = torch.randn(B, cout, H, W, device=dev)
    dx = torch.empty_like(x)
    dw = torch.zeros_like(w)
    tiles = torch.empty(cout, L.i2p_conv3x3_stat_slots(B, cout, H, W), 3, device=dev)
    pf = torch.empty(L.i2p_conv3x3_pack_floats(cin, cout, 0), device=dev)
    pd = torch.empty(L.i2p_conv3x3_pack_floats(cin, cout, 1), device=dev)
    _cabi.call("i2p_conv3x3_pack", dev, cin, cout, 0, w.data_ptr(), pf.data_ptr())
    _cabi.call("i2p_conv3x3_pack", dev, cin, cout, 1, w.data_ptr(), pd.data_ptr())
    f = timed(lambda: _cabi.call("i2p_conv3x3_tc", dev, B, cin, cout, H, W, x.data_ptr(), pf.data_ptr(), bias.data_ptr(),
                                 y.data_ptr(), tiles.data_ptr()))
    d = timed(lambda: _cabi.call("i2p_conv3x3_tc", dev, B, cout, cin, H, W, dy.data_ptr(), pd.data_ptr(), None,
                                 dx.data_ptr(), None))
    g = timed(lambda: _cabi.call("i2p_conv3x3_wgrad", dev, B, cin, cout, H, W, x.data_ptr(), dy.data_ptr(), dw.data_ptr()))
    bytes_f = 4 * B * H * W * (cin + cout)
    flops = 2.0 * B * H * W * cin * cout * 9
    print("| %d -> %d @ %dx%d (x%d) | %.1f (%.0f) | %.1f | %.1f (%.1f) |" % (cin, cout, H, W, n, f, bytes_f / f / 1e3, d, g,
                                                                        flops / g / 1e6))
    tot["fwd"] += n * f
    tot["dgrad"] += n * d if cin > 3 else 0.0
    tot["wgrad"] += n * g
print("\nper step (15 layers, batch %d): forward %.0f us, data gradient %.0f us, weight gradient %.0f us" % (
    B, tot["fwd"], tot["dgrad"], tot["wgrad"]))
