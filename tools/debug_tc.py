import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200 import _cabi
from i2pnet_b200.projectPN import PPBackbone_center as P
from i2pnet_b200.projectPN.fused_mlp import FusedMLPFunction
L = _cabi.lib(); dev = torch.device("cuda:0")
for cin, chans, B, n, k, leaky in [(262, [128, 64, 64], 2, 228, 80, True), (131, [128, 128, 256], 2, 116, 16, False)]:
    torch.manual_seed(1)
    mods, c = [], cin
    for co in chans:
        mods.append(P.Conv2d(c, co, [1, 1], bn=True, leaky_relu=leaky).to(dev)); c = co
    x0 = torch.randn(B, n, k, cin, device=dev) * 3 + 1.5
    gout = None
    ref = None
    for rep in range(6):
        tc = rep % 2
        L.i2p_set_mlp_tensor_cores(tc)
        x = x0.clone().requires_grad_(True)
        for m in mods: m.zero_grad()
        out = P.run_mlp(mods, x, reduce_k=False)
        if gout is None: gout = torch.randn_like(out)
        saved = [t.detach().clone() for t in out.grad_fn.next_functions[0][0].saved_tensors if t is not None]
        out.backward(gout)
        cur = dict(out=out.detach().clone(), dx=x.grad.clone(), saved=saved, gw=[m.conv.weight.grad.clone() for m in mods])
        if ref is None: ref = cur
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        print("cin%d rep%d tc=%d out %.1e dx %.1e gw %s" % (cin, rep, tc, rel(cur["out"], ref["out"]), rel(cur["dx"], ref["dx"]),
              ["%.1e" % rel(a, b) for a, b in zip(cur["gw"], ref["gw"])]))
        if rep == 1:
            for i, (a, b) in enumerate(zip(cur["saved"], ref["saved"])):
                if not a.dtype.is_floating_point or rel(a, b) == 0: continue
                if a.dim() == 2 and a.shape[0] == 4:
                    print("    saved[%d] stats rows rel:" % i, ["%.1e" % rel(a[j], b[j]) for j in range(4)])
                else:
                    d = (a - b).abs()
                    print("    saved[%d] %s rel %.1e  n(|d|>1e-4*max)=%d  argmax %s" % (i, tuple(a.shape), rel(a, b), int((d > 1e-4 * b.abs().max()).sum()), tuple(int(v) for v in torch.nonzero(d == d.max())[0])))
