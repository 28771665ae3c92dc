"""Debug: one pyramid (own kernels) against the ATen / library formulation in f64, block by block, repeated to expose
run-to-run differences."""
import copy
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from i2pnet_b200.modules import basicConv
from i2pnet_b200.modules.basicConv import createCNNs

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, cin, chans, strides, H, W = 2, 3, [16, 16, 32], [2, 1, 2], 160, 512
torch.manual_seed(sum(chans) + H)
net = createCNNs(cin, chans, strides).to(dev)
with torch.no_grad():
    for m in net:
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.uniform_(0.5, 1.5)
            m.bias.uniform_(-0.5, 0.5)
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.5, 2.0)
net.train()
x = torch.rand(B, cin, H, W, device=dev) * 4 - 1


def run(dtype, own):
    n = copy.deepcopy(net).to(dtype)
    xx = x.detach().clone().to(dtype).requires_grad_(True)
    basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = own
    outs = []
    h = xx
    mods = list(n)
    for i in range(0, len(mods), 4):
        sub = basicConv._Pyramid()
        for j, m in enumerate(mods[i:i + 4]):
            sub.add_module(str(j), m)
        h = sub(h)
        outs.append(h)
    torch.manual_seed(1)
    g = torch.randn(h.shape, device=dev, dtype=torch.float64).to(dtype)
    h.backward(g)
    basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = True
    return [o.detach().double() for o in outs], xx.grad.double(), {k: p.grad.double() for k, p in n.named_parameters()}


truth = run(torch.float64, False)
aten = run(torch.float32, False)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
l2 = lambda a, b: float((a - b).norm() / b.norm())
print("aten  fwd per block", [rel(a, t) for a, t in zip(aten[0], truth[0])], "dx l2", l2(aten[1], truth[1]))
for rep in range(4):
    mine = run(torch.float32, True)
    torch.cuda.synchronize()
    diff = (mine[1] - truth[1]).abs()
    srt = diff.flatten().sort(descending=True)[0]
    print("   |dx diff| / max|dx|: top", [float(v / truth[1].abs().max()) for v in srt[[0, 10, 100, 1000, 10000, 100000]]],
          "per (b, c) l2:", [["%.1e" % l2(mine[1][b, c], truth[1][b, c]) for c in range(3)] for b in range(2)])
    nd = diff > 1e-3 * truth[1].abs().max()
    print("own   fwd per block", [rel(a, t) for a, t in zip(mine[0], truth[0])], "dx l2", l2(mine[1], truth[1]),
          "dx elements off by > 1e-3 max:", int(nd.sum()), "grads", {k: "%.1e" % l2(v, truth[2][k]) for k, v in mine[2].items() if "bias" not in k or k.startswith("1") or k.startswith("5")})
