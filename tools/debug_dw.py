"""Diagnose the tensor-core dW kernel: dY is made equal to g (scale 1, shift large, zero sums), so dw = g^T x."""
import os, sys, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run():
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for rows, cin, cout in ((256, 128, 128), (256, 64, 64), (1000, 262, 128)):
        x = torch.randn(rows, cin, device=dev)
        g = torch.randn(rows, cout, device=dev)
        y = torch.zeros(rows, cout, device=dev)
        st = torch.stack([torch.zeros(cout), torch.ones(cout), torch.ones(cout), torch.full((cout,), 10.0)]).to(dev).contiguous()
        s12 = torch.zeros(2, cout, dtype=torch.float64, device=dev)
        dw = torch.zeros(cout, cin, device=dev)
        call("i2p_pw_linear_bwd_dw_tc", dev, rows, cin, cout, g.data_ptr(), None, None, 1, y.data_ptr(), st[0].data_ptr(), st[1].data_ptr(),
             st[2].data_ptr(), st[3].data_ptr(), 0.1, s12.data_ptr(), x.data_ptr(), None, None, 1.0, dw.data_ptr())
        torch.cuda.synchronize()
        truth = (g.double().t() @ x.double())
        d = dw.double()
        err = float((d - truth).abs().max() / truth.abs().max())
        corr = float((d * truth).sum() / (d.norm() * truth.norm()).clamp_min(1e-30))
        print("DBG=%s rows %d %d->%d: max|dw| %.3e max|truth| %.3e rel err %.3e cosine %.4f frac zero %.3f nan %d" % (
            os.environ.get("I2P_DW_DBG", "0"), rows, cin, cout, float(d.abs().max()), float(truth.abs().max()), err, corr,
            float((d == 0).double().mean()), int(torch.isnan(d).sum())))
        if err > 1e-3:
            # does the result match the truth under a transpose / a permutation of small blocks?
            for name, alt in (("2x", 2 * truth), ("0.5x", 0.5 * truth)):
                print("    vs %s: %.3e" % (name, float((d - alt).abs().max() / truth.abs().max())))
            # one-hot probe: x = e_i for a single row r*, g = e_o -> dw should have a single 1 at (o, i)
            for (r_, i_, o_) in ((0, 0, 0), (1, 0, 0), (8, 0, 0), (0, 1, 0), (0, 4, 0), (0, 0, 1), (0, 0, 4), (9, 5, 6), (33, 17, 70 % cout)):
                x.zero_(); g.zero_(); dw.zero_()
                x[r_, i_] = 1.0; g[r_, o_] = 1.0
                call("i2p_pw_linear_bwd_dw_tc", dev, rows, cin, cout, g.data_ptr(), None, None, 1, y.data_ptr(), st[0].data_ptr(), st[1].data_ptr(),
                     st[2].data_ptr(), st[3].data_ptr(), 0.1, s12.data_ptr(), x.data_ptr(), None, None, 1.0, dw.data_ptr())
                torch.cuda.synchronize()
                nz = torch.nonzero(dw.abs() > 1e-3)
                print("    one-hot row %d: x[:, %d] g[:, %d] -> nonzero dw at (o, i) = %s values %s" % (
                    r_, i_, o_, nz[:6].tolist(), [round(float(dw[a, b]), 3) for a, b in nz[:6].tolist()]))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for d in (0, 2, 1, 4):
            out = subprocess.run([sys.executable, __file__, "x"], capture_output=True, text=True, env=dict(os.environ, I2P_DW_DBG=str(d)))
            print(out.stdout[-6000:], out.stderr[-1500:] if out.returncode else "", flush=True)
