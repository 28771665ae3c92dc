import os, sys, statistics, subprocess
if len(sys.argv) == 1:
    for d in (0, 1, 2, 4, 8, 6, 14, 15):
        out = subprocess.run([sys.executable, __file__, str(d)], capture_output=True, text=True, env=dict(os.environ, I2P_TC_DBG=str(d), I2P_MLP_TC="1"))
        print("dbg=%2d" % d, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
    sys.exit(0)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi
from i2pnet_b200._cabi import call
dev = torch.device("cuda:0"); L = _cabi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = []
for rows, cin, cout in [(145920, 128, 64), (145920, 262, 128)]:
    x = torch.randn(rows, cin, device=dev); sc = torch.rand(cin, device=dev) + 0.5; sh = torch.randn(cin, device=dev)
    w = torch.randn(cout, cin, device=dev) * 0.1; b = torch.randn(cout, device=dev); y = torch.empty(rows, cout, device=dev)
    tiles = torch.empty(L.i2p_pw_num_tiles(rows), cout, 2, device=dev)
    ts = []
    for i in range(10):
        flush.fill_(i)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        call("i2p_pw_linear_fwd", dev, rows, cin, cout, x.data_ptr(), sc.data_ptr(), sh.data_ptr(), 0.1, w.data_ptr(), b.data_ptr(), y.data_ptr(), tiles.data_ptr())
        e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e) * 1e3)
    res.append("%d->%d %.1f us" % (cin, cout, statistics.median(ts[2:])))
print("  ".join(res))
