"""GPU: the fused shared-MLP kernels (csrc/mlp.cu) against the layer-by-layer ATen formulation of
the same Conv2d blocks, and both against an f64 evaluation of the reference's formulation
(NCHW 1x1 conv + BatchNorm2d + activation, PPBackbone_center.py:35-46)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (cin, channels, batch, n, k, reduce over k, leaky, input needs grad)
CONFIGS = [
    (10, [16, 16, 32], 2, 3600, 32, True, False, False),     # SA1
    (35, [32, 32, 64], 2, 904, 16, True, False, True),       # SA2
    (131, [128, 128, 256], 2, 116, 16, True, False, True),   # SA4
    (67, [128, 64], 2, 228, 8, True, True, True),            # up-conv mlp_conv
    (262, [128, 64, 64], 2, 228, 80, False, True, True),     # cost volume 1 mlp1
    (6, [64], 2, 228, 80, False, True, False),               # pi_encoding (single layer)
    (320, [128, 64], 3, 116, 1, False, True, True),          # flow predictor
    (192, [64], 1, 100, 3, False, True, True),               # odd row count (300 rows: partial tile)
]


def _reference_f64(x, mods, reduce_k):
    y = x.double()
    for m in mods:
        w = m.conv.weight.double().view(m.out_channels, m.in_channels)
        y = y @ w.t() + m.conv.bias.double()
        flat = y.reshape(-1, y.shape[-1])
        mean, var = flat.mean(0), flat.var(0, unbiased=False)
        y = (y - mean) / torch.sqrt(var + m.bn_linear.eps) * m.bn_linear.weight.double() + m.bn_linear.bias.double()
        y = F.leaky_relu(y, 0.1) if m.leaky_relu else F.relu(y)
    return y.max(dim=2)[0] if reduce_k else y


@pytest.fixture(params=[0, 1], ids=["fma", "tcgen05"])
def tensor_cores(request):
    """Both kernel families: f32 FMA and tcgen05 tf32 with the 3-term hi/lo split."""
    from i2pnet_b200 import _cabi
    L = _cabi.lib()
    before = L.i2p_get_mlp_tensor_cores()
    L.i2p_set_mlp_tensor_cores(request.param)
    yield request.param
    L.i2p_set_mlp_tensor_cores(before)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "cin%d_%s_k%d" % (c[0], "x".join(map(str, c[1])), c[4]))
def test_fused_mlp_matches_layerwise_and_f64(cfg, tensor_cores):
    from i2pnet_b200.projectPN import PPBackbone_center as P
    cin, chans, B, n, k, reduce_k, leaky, need_grad = cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(sum(chans) + cin)
    mods, c = [], cin
    for co in chans:
        m = P.Conv2d(c, co, [1, 1], bn=True, leaky_relu=leaky).to(dev)
        with torch.no_grad():
            m.bn_linear.weight.uniform_(0.5, 1.5)
            m.bn_linear.bias.uniform_(-0.5, 0.5)
        mods.append(m)
        c = co
    x0 = (torch.randn(B, n, k, cin, device=dev) * 3 + 1.5)
    gout = None
    res = {}
    for mode in ("fused", "layerwise", "f64"):
        x = x0.clone().requires_grad_(need_grad)
        for m in mods:
            m.zero_grad()
        if mode == "f64":
            out = _reference_f64(x, mods, reduce_k)
        else:
            P.USE_FUSED_MLP = mode == "fused"
            out = P.run_mlp(mods, x, reduce_k=reduce_k)
        P.USE_FUSED_MLP = True
        if gout is None:
            gout = torch.randn(out.shape, device=dev)
        out.backward(gout.to(out.dtype))
        res[mode] = dict(out=out.detach().double(), dx=x.grad.double() if need_grad else None,
                         grads=[(n_, p.grad.double().clone()) for m in mods for n_, p in m.named_parameters()])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    l2 = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    truth = res["f64"]
    for mode in ("fused", "layerwise"):
        r = res[mode]
        assert rel(r["out"], truth["out"]) < 1e-5, mode
    # Gradients.  act'(z) is discontinuous at z = 0: a forward rounding difference of 1e-6 flips the
    # slope (1 <-> 0.1 or 0) of a few dozen of the ~5e6 pre-activations, and each flip changes one row
    # of dx by tens of percent.  The max-norm therefore measures luck, not accuracy (the f32-FMA path
    # differs from f64 by 1e-7 per layer and usually flips nothing; the tcgen05 3xTF32 path differs
    # by 2e-6 and flips ~50).  Gradients are compared in relative L2, plus the fraction of elements
    # off by more than 1e-4 of the maximum.
    f, lw = res["fused"], res["layerwise"]

    def check(name, g, t):
        assert l2(g, t) < 2e-3, (name, l2(g, t))
        outliers = float(((g - t).abs() > 1e-4 * t.abs().max()).float().mean())
        assert outliers < 2e-3, (name, outliers)

    if need_grad:
        check("dx", f["dx"], truth["dx"])
    for (name, gf), (_, gt) in zip(f["grads"], truth["grads"]):
        if name == "conv.bias":
            assert float(gf.abs().max()) == 0.0          # exactly zero under batch-norm
            continue
        check(name, gf, gt)
    if not tensor_cores:   # the f32 FMA kernels are as close to f64 as the ATen formulation
        if need_grad:
            assert l2(f["dx"], truth["dx"]) < max(1e-5, 3 * l2(lw["dx"], truth["dx"]))
        for (name, gf), (_, gl), (_, gt) in zip(f["grads"], lw["grads"], truth["grads"]):
            if name != "conv.bias":
                assert l2(gf, gt) < max(1e-5, 3 * l2(gl, gt)), (name, l2(gf, gt), l2(gl, gt))
