"""Debug: the golden batch-2 model with own vs library convolutions: RF3, coarse pose, kNN sets, cost_volume2, out3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.test_host_logic_cpu import build_model, load_golden_model
from i2pnet_b200.modules import basicConv
from i2pnet_b200.projectPN import PPBackbone_center as P
from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g, state = load_golden_model()
dev = "cuda:0"
model = build_model(state, dev)
t = lambda k: torch.from_numpy(g[k]).to(dev)
rgb = torch.from_numpy(g["rgb_u8"]).float().to(dev)
res = {}
orig_knn = P.knn_point
for own in (True, False, True):
    basicConv.USE_OWN_CONV = own
    cap = {}
    def knn(k, a, b, _cap=cap):
        out = orig_knn(k, a, b)
        _cap["knn"] = out.clone()
        _cap["warped"] = b.clone()
        return out
    P.knn_point = knn
    hooks = []
    for name in ("RGB_net3", "cost_volume1", "cost_volume2", "l4_head"):
        def hook(mod, args, out, name=name, _cap=cap):
            _cap[name] = (out[0] if isinstance(out, tuple) else out).detach().clone()
        hooks.append(getattr(model, name).register_forward_hook(hook))
    with torch.no_grad():
        out = model(rgb, t("lidar"), t("raw_point_xyz"), None, t("intrinsic"), None, None, None, t("lidar_feats"), cfg)
    torch.cuda.synchronize()
    cap["out3"], cap["out4"] = out[0].clone(), out[1].clone()
    for h in hooks:
        h.remove()
    res.setdefault(own, []).append(cap)
P.knn_point = orig_knn
basicConv.USE_OWN_CONV = True
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
a, b, a2 = res[True][0], res[False][0], res[True][1]
for k in ("RGB_net3", "cost_volume1", "l4_head", "out4", "warped", "cost_volume2", "out3"):
    print("%-14s own vs library %.2e   own vs own (rerun) %.2e" % (k, rel(a[k], b[k]), rel(a[k], a2[k])))
same = (torch.sort(a["knn"], -1)[0] == torch.sort(b["knn"], -1)[0]).all(-1)
print("kNN sets differing: %d of %d queries" % (int((~same).sum()), same.numel()))
print("golden: out3 own %.2e library %.2e ; out4 own %.2e library %.2e" % (
    rel(a["out3"].cpu(), torch.from_numpy(g["out3"])), rel(b["out3"].cpu(), torch.from_numpy(g["out3"])),
    rel(a["out4"].cpu(), torch.from_numpy(g["out4"])), rel(b["out4"].cpu(), torch.from_numpy(g["out4"]))))
