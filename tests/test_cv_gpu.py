"""GPU: the fused cost-volume glue (csrc/cv.cu: operand build + its backward, softmax-weighted sum + its
backward) inside CostVolume against the reference's broadcast / mask / max / concatenate / softmax formulation
(src/projectPN/PPBackbone_center.py:354-490) run through ATen on the same device, and against f64."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(kind, dev):
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg
    from i2pnet_b200.projectPN.PPBackbone_center import CostVolume
    i = 0 if kind == "all_pixels" else 1
    torch.manual_seed(3 + i)
    cv = CostVolume(H=4, W=57, kernel_size=cfg.cost_volume_kernel_size[i], distance=cfg.cost_volume_dis[i],
                    nsample=cfg.cost_volume_nsamples[0], nsample_q=cfg.cost_volume_nsamples[1][i], rgb_in_channels=128,
                    lidar_in_channels=128, mlp1=cfg.cost_volume_mlps[0], mlp2=cfg.cost_volume_mlps[1],
                    backward_validation=cfg.backward_validation[i], use_trans=cfg.use_trans).to(dev)
    return cv, cfg


def _inputs(dev, B=2, H=4, W=57, n2=80, seed=0):
    from i2pnet_b200.projectPN.utils import StrideGrid
    g = torch.Generator(device=dev).manual_seed(seed)
    N = H * W
    xyz = torch.randn(B, N, 3, device=dev, generator=g) * 5
    xyz[..., 2] = xyz[..., 2].abs() + 2.0
    xyz[:, ::7] = 0                                              # empty range-image cells: invalid points
    z = xyz[:, :, 2:]
    uv = xyz / (z + 1e-10)
    feats = torch.randn(B, N, 128, device=dev, generator=g)
    pix_xyz = torch.cat([torch.rand(B, n2, 2, device=dev, generator=g) * 2 - 1, torch.ones(B, n2, 1, device=dev)], -1)
    pix_feat = torch.randn(B, n2, 128, device=dev, generator=g)
    return xyz.view(B, H, W, 3), uv, feats, StrideGrid(B, H, W, 1, 1, dev), pix_xyz, pix_feat, z


@pytest.mark.parametrize("kind", ["all_pixels", "knn_pixels"])
def test_cost_volume_fused_glue_matches_reference_formulation(kind):
    from i2pnet_b200 import _cabi
    from i2pnet_b200.projectPN import PPBackbone_center as P
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    cv, cfg = _make(kind, dev)
    raw, uv, feats, grid, pxyz, pfeat, z = _inputs(dev)
    res = {}
    gout = None
    for mode in ("fused", "aten"):
        P.USE_FUSED_CV = mode == "fused"
        try:
            m = copy.deepcopy(cv)
            a = [t.clone().requires_grad_(True) for t in (uv, feats, pxyz, pfeat)]
            before = _cabi.launch_count()
            out = m(raw, a[0], a[1], grid, a[2], a[3], z, cfg=cfg)
            if gout is None:
                gout = torch.randn_like(out)
            out.backward(gout)
            launches = _cabi.launch_count() - before
        finally:
            P.USE_FUSED_CV = True
        res[mode] = dict(out=out.detach(), grads=[t.grad for t in a], params=[p.grad for p in m.parameters()], launches=launches)
    assert res["fused"]["launches"] > res["aten"]["launches"]                 # the four extra kernels ran
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
    l2 = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-30))
    assert rel(res["fused"]["out"], res["aten"]["out"]) < 1e-5
    names = ["d warped_xyz", "d warped_points", "d f2_xyz", "d f2_points"]
    for n, gf, ga in zip(names, res["fused"]["grads"], res["aten"]["grads"]):
        assert l2(gf, ga) < 2e-3, (n, l2(gf, ga))            # LeakyReLU slope flips: relative L2, see test_mlp_gpu.py
    for gf, ga in zip(res["fused"]["params"], res["aten"]["params"]):
        if ga.abs().max() > 1e-6:
            assert l2(gf, ga) < 2e-3


def test_softmax_wsum_matches_torch():
    from i2pnet_b200.projectPN.PPBackbone_center import _softmax_wsum
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for (B, N, K, C, masked) in [(2, 228, 80, 64, False), (2, 228, 4, 64, True), (1, 7, 3, 5, True)]:
        l0, v0 = torch.randn(B, N, K, C, device=dev) * 3, torch.randn(B, N, K, C, device=dev)
        mask = (torch.rand(B, N, K, 1, device=dev) > 0.3).float() if masked else None
        if masked:
            mask[:, 0] = 0                                        # a point with no valid neighbour at all
        outs = []
        for dt, fused in ((torch.float32, True), (torch.float64, False)):
            l, v = l0.detach().to(dt).clone().requires_grad_(True), v0.detach().to(dt).clone().requires_grad_(True)
            m = mask.to(dt) if masked else None
            if fused:
                o = _softmax_wsum(l, v, m)
            else:
                ll = l * m + -1e10 * (1 - m) if masked else l
                o = torch.sum(F.softmax(ll, dim=2) * v, dim=2)
            o.backward(torch.ones_like(o) * 0.5)
            outs.append((o.detach().double(), l.grad.double(), v.grad.double()))
        for a, b in zip(*outs):
            assert float((a - b).abs().max()) < 2e-5 * max(1.0, float(b.abs().max()))
