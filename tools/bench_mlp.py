"""Time the shared-MLP kernels at the step's main layer shapes, f32 FMA vs tcgen05 (CUDA events, L2 flushed)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi
from i2pnet_b200._cabi import call

dev = torch.device("cuda:0")
L = _cabi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [  # rows, cin, cout
    (145920, 262, 128), (145920, 128, 64), (145920, 64, 64), (145920, 128, 128), (58368, 134, 128),
    (115712, 35, 32), (29184, 67, 64), (29184, 64, 128), (14848, 131, 128), (14848, 128, 256), (921600, 16, 32),
]


def timeit(fn, n=12):
    ts = []
    for i in range(n):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[2:])


for rows, cin, cout in SHAPES:
    x = torch.randn(rows, cin, device=dev)
    sc, sh = torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev)
    w, b = torch.randn(cout, cin, device=dev) * 0.1, torch.randn(cout, device=dev)
    y = torch.empty(rows, cout, device=dev)
    tiles = torch.empty(L.i2p_pw_num_tiles(rows), cout, 2, device=dev)
    res = {}
    for tc in (0, 1):
        L.i2p_set_mlp_tensor_cores(tc)
        f = lambda: call("i2p_pw_linear_fwd", dev, rows, cin, cout, x.data_ptr(), sc.data_ptr(), sh.data_ptr(), 0.1,
                         w.data_ptr(), b.data_ptr(), y.data_ptr(), tiles.data_ptr())
        res[tc] = timeit(f)
        if tc == 0:
            y0 = y.clone()
    err = float((y - y0).abs().max() / y0.abs().max())
    t_mm = timeit(lambda: torch.addmm(b, x, w.t(), out=y))
    gf = 2.0 * rows * cin * cout / 1e9
    mb = 4.0 * rows * (cin + cout) / 1e6
    print("fwd rows=%7d %3d->%3d  fma %7.1f us (%5.1f TF/s)  tcgen05 %7.1f us (%5.1f TF/s, %4.0f GB/s)  cublas-addmm %7.1f us  maxrel %.1e"
          % (rows, cin, cout, res[0], gf / res[0] * 1e3, res[1], gf / res[1] * 1e3, mb / res[1] * 1e3, t_mm, err))
