"""Time the shared-MLP GEMM kernels at the step's main layer shapes: f32 FMA vs tcgen05 (forward, dX, dW),
CUDA events on the launching stream, L2 flushed between launches.  Prints one line per shape and kernel with
the achieved algorithmic HBM rate (bytes of SURVEY.md section 8d: 4*rows*(cin+cout) forward,
4*rows*(2*cout+cin) backward) next to the measured copy peak."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi  # noqa: E402
from i2pnet_b200._cabi import call  # noqa: E402

dev = torch.device("cuda:0")
L = _cabi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
SHAPES = [  # rows, cin, cout
    (145920, 262, 128), (145920, 128, 64), (145920, 64, 64), (145920, 128, 128), (58368, 134, 128), (58368, 128, 64),
    (29184, 67, 64), (29184, 64, 128), (14848, 131, 128), (14848, 128, 256), (145920, 6, 64),
    # the narrow layers of SA1 (8 x 3600 x 32 rows) and SA2 (8 x 904 x 16 rows): FMA kernels only (run once with
    # I2P_DW_SKINNY=0 and once with the default to compare the two dW kernels for cin <= 16)
    (921600, 10, 16), (921600, 16, 16), (921600, 16, 32), (115712, 35, 32), (115712, 32, 32), (115712, 32, 64),
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


def main():
    for rows, cin, cout in SHAPES:
        x = torch.randn(rows, cin, device=dev)
        sc, sh = torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev)
        has_tf = cin % 16 == 0
        w, b = torch.randn(cout, cin, device=dev) * 0.1, torch.randn(cout, device=dev)
        y = torch.empty(rows, cout, device=dev)
        g = torch.randn(rows, cout, device=dev)
        tiles = torch.empty(L.i2p_pw_num_tiles(rows), cout, 2, device=dev)
        st = torch.empty(4, cout, device=dev)
        tc_fwd = bool(L.i2p_pw_tc_supported(0, rows, cin, cout))
        pack = torch.empty(max(L.i2p_pw_pack_floats(cin, cout), 4) if tc_fwd else 4, device=dev)
        s12 = torch.zeros(2, cout, dtype=torch.float64, device=dev)
        ps12 = torch.zeros(2, cin, dtype=torch.float64, device=dev)
        dx = torch.empty(rows, cin, device=dev)
        dw = torch.zeros(cout, cin, device=dev)
        psc, psh = (sc.data_ptr(), sh.data_ptr()) if has_tf else (None, None)
        pst = torch.stack([sh, sc, sc, sh]).contiguous()
        prev = (x.data_ptr(), pst[0].data_ptr(), pst[1].data_ptr(), pst[2].data_ptr(), pst[3].data_ptr(), 0.1) if has_tf \
            else (None, None, None, None, None, 1.0)
        pp = ps12.data_ptr() if has_tf else None
        t_pack = timeit(lambda: call("i2p_pw_pack_weights", dev, cin, cout, w.data_ptr(), pack.data_ptr())) if tc_fwd else 0.0

        def fwd_mask(mask):
            L.i2p_set_mlp_tensor_cores(mask)
            if mask & 1:
                return lambda: call("i2p_pw_linear_fwd_tc", dev, rows, cin, cout, x.data_ptr(), psc, psh, 0.1, pack.data_ptr(),
                                    b.data_ptr(), y.data_ptr(), tiles.data_ptr())
            return lambda: call("i2p_pw_linear_fwd", dev, rows, cin, cout, x.data_ptr(), psc, psh, 0.1, w.data_ptr(),
                                b.data_ptr(), y.data_ptr(), tiles.data_ptr())
        res = {}
        for name, mask in (("fma", 0), ("tc", 1), ("tc_sw", 17)):
            if mask and not tc_fwd:
                continue
            if mask & 16:   # the pack layout follows the mask
                L.i2p_set_mlp_tensor_cores(mask)
                call("i2p_pw_pack_weights", dev, cin, cout, w.data_ptr(), pack.data_ptr())
            res["fwd_" + name] = timeit(fwd_mask(mask))
            if name == "fma":
                y0 = y.clone()
        err = float((y - y0).abs().max() / y0.abs().max())
        L.i2p_set_mlp_tensor_cores(7)
        if tc_fwd:
            call("i2p_pw_pack_weights", dev, cin, cout, w.data_ptr(), pack.data_ptr())
        call("i2p_bn_finalize", dev, rows, cout, tiles.data_ptr(), sc[:1].expand(cout).contiguous().data_ptr(), b.data_ptr(), 1e-5,
             st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr())
        bn = (y.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), 0.1)
        res["dx_fma"] = timeit(lambda: call("i2p_pw_linear_bwd_dx", dev, rows, cin, cout, g.data_ptr(), None, None, 1, *bn,
                                            s12.data_ptr(), w.data_ptr(), dx.data_ptr(), *prev, pp))
        res["dw_fma"] = timeit(lambda: call("i2p_pw_linear_bwd_dw", dev, rows, cin, cout, g.data_ptr(), None, None, 1, *bn,
                                            s12.data_ptr(), x.data_ptr(), psc, psh, 0.1 if has_tf else 1.0, dw.data_ptr()))
        if tc_fwd and L.i2p_pw_tc_supported(1, rows, cin, cout):
            res["dx_tc"] = timeit(lambda: call("i2p_pw_linear_bwd_dx_tc", dev, rows, cin, cout, g.data_ptr(), None, None, 1, *bn, s12.data_ptr(),
                                               pack.data_ptr(), dx.data_ptr(), *prev, pp))
            L.i2p_set_mlp_tensor_cores(23)
            call("i2p_pw_pack_weights", dev, cin, cout, w.data_ptr(), pack.data_ptr())
            res["dx_tc_sw"] = timeit(lambda: call("i2p_pw_linear_bwd_dx_tc", dev, rows, cin, cout, g.data_ptr(), None, None, 1, *bn, s12.data_ptr(),
                                                  pack.data_ptr(), dx.data_ptr(), *prev, pp))
            L.i2p_set_mlp_tensor_cores(7)
            res["dw_tc"] = timeit(lambda: call("i2p_pw_linear_bwd_dw_tc", dev, rows, cin, cout, g.data_ptr(), None, None, 1, *bn, s12.data_ptr(),
                                               x.data_ptr(), psc, psh, 0.1 if has_tf else 1.0, dw.data_ptr()))
        t_mm = timeit(lambda: torch.addmm(b, x, w.t(), out=y))
        fb, bb = 4.0 * rows * (cin + cout), 4.0 * rows * (2 * cout + cin)
        line = "rows=%7d %3d->%3d pack %5.1f us | fwd" % (rows, cin, cout, t_pack)
        for k in ("fma", "tc", "tc_sw"):
            if "fwd_" + k not in res:
                continue
            t = res["fwd_" + k]
            line += "  %s %6.1f us (%4.0f GB/s %.2f)" % (k, t, fb / t / 1e3, fb / t / 1e3 / PEAK)
        line += "  cublas %6.1f us  maxrel %.1e |" % (t_mm, err)
        for k in ("dx_fma", "dx_tc", "dx_tc_sw", "dw_fma", "dw_tc"):
            if k in res:
                line += "  %s %6.1f us (%4.0f GB/s %.2f)" % (k, res[k], bb / res[k] / 1e3, bb / res[k] / 1e3 / PEAK)
        print(line, flush=True)


if __name__ == "__main__":
    main()
