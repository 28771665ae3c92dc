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


def _prep_reference(uv, z, pf, qf, has_max):
    """The ATen spelling of the operand preparation (PPBackbone_center.py:379-397 as _first_layer_operand_fused had it)."""
    from i2pnet_b200.projectPN.PPBackbone_center import _standardise
    from i2pnet_b200.projectPN.utils import check_valid
    xyz = uv.mul(z)
    pi_n, qi_n = _standardise(pf), _standardise(qf)
    maxc = None
    if has_max:
        valid = check_valid(xyz) > 0
        any_valid = valid.any(dim=1, keepdim=True)
        hi = torch.where(valid, pi_n, float("-inf")).max(dim=1, keepdim=True)[0]
        lo = torch.where(valid, pi_n, float("inf")).min(dim=1, keepdim=True)[0]
        hi, lo = torch.where(any_valid, hi, 0.0), torch.where(any_valid, lo, 0.0)
        maxc = torch.where(qi_n > 0, qi_n * hi, qi_n * lo)
        maxc = torch.where(any_valid, maxc, -1e10)
    return xyz, pi_n, qi_n, maxc


@pytest.mark.parametrize("shape", [(2, 228, 80, 128), (3, 57, 33, 64), (1, 5, 3, 7), (2, 40, 20, 256)],
                         ids=lambda s: "b%d_n%d_n2%d_c%d" % s)
@pytest.mark.parametrize("has_max", [True, False], ids=["max", "nomax"])
def test_cv_prep_matches_aten_and_f64(shape, has_max):
    """csrc/cv.cu cv_prep (one kernel per direction) against the element-wise formulation through ATen in f32 and f64:
    outputs and the four input gradients; one cloud without any valid point, constant rows (clipped denominator)."""
    from i2pnet_b200 import _cabi
    from i2pnet_b200.projectPN.PPBackbone_center import _CvPrep
    B, N, N2, C = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(N + C)
    uv = torch.randn(B, N, 3, device=dev, generator=g)
    z = torch.rand(B, N, 1, device=dev, generator=g) * 30 + 1
    z[:, ::5] = 0                                                   # invalid points
    if B > 1:
        z[-1] = 0                                                   # a cloud without a valid point
    pf = torch.randn(B, N, C, device=dev, generator=g) * 2 + 0.3
    qf = torch.randn(B, N2, C, device=dev, generator=g)
    pf[0, 1] = 0.75                                                 # constant rows: std = 0, denominator clipped
    qf[0, 0] = 0.0
    gouts = None
    res = {}
    for mode, dt in (("own", torch.float32), ("aten", torch.float32), ("f64", torch.float64)):
        a = [t.detach().clone().to(dt).requires_grad_(True) for t in (uv, z, pf, qf)]
        before = _cabi.launch_count()
        outs = _CvPrep.apply(*a, has_max) if mode == "own" else _prep_reference(*a, has_max)
        if mode == "own":
            assert _cabi.launch_count() - before == (2 if has_max else 1)       # + the element-wise maxc kernel
        outs = [o for o in outs if o is not None]
        if gouts is None:
            gouts = [torch.randn(o.shape, device=dev, generator=g) for o in outs]
        torch.autograd.backward(outs, [go.to(dt) for go in gouts])
        res[mode] = dict(outs=[o.detach().double() for o in outs], grads=[t.grad.double() for t in a])
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
    names = ["xyz", "pi", "qi", "maxc"]
    for i, (o, a_, t) in enumerate(zip(res["own"]["outs"], res["aten"]["outs"], res["f64"]["outs"])):
        assert rel(o, t) < max(2e-6, 2 * rel(a_, t)), (names[i], rel(o, t), rel(a_, t))
    for i, (o, a_, t) in enumerate(zip(res["own"]["grads"], res["aten"]["grads"], res["f64"]["grads"])):
        # the constant rows' gradient is (g - mean g) / 1e-12: compare relative to the largest entry
        assert rel(o, t) < max(5e-6, 2 * rel(a_, t)), (["d uv", "d z", "d pf", "d qf"][i], rel(o, t), rel(a_, t))


def test_pixel_rays_match_the_reference_formulation():
    """csrc/cv.cu pixel_rays against change_intrinsic -> inverse -> set_id_grid -> batched product
    (src/modellearn_proj_center.py:275-287) in f64."""
    from i2pnet_b200.modellearn_proj_center import change_intrinsic, set_id_grid
    from i2pnet_b200.projectPN.utils import pixel_rays
    dev = torch.device("cuda:0")
    B, h, w, H, W = 3, 5, 16, 160, 512
    K = torch.tensor([[718.856, 0.0, 607.19], [0.0, 718.856, 185.2], [0.0, 0.0, 1.0]], device=dev).repeat(B, 1, 1)
    K[1, 0, 0] *= 0.5
    K[2, 1, 2] += 3.0
    rays = pixel_rays(K, h, w, H, W)
    rf = torch.empty(B, 8, h, w, device=dev, dtype=torch.float64)
    K3 = change_intrinsic(K.double(), rf, torch.empty(B, 3, H, W))
    ref = torch.bmm(torch.linalg.inv(K3), set_id_grid(rf.permute(0, 2, 3, 1)).permute(0, 2, 1)).permute(0, 2, 1)
    assert rays.shape == (B, h * w, 3)
    assert float((rays.double() - ref).abs().max()) < 1e-6 * float(ref.abs().max())
    assert torch.allclose(pixel_rays(K.cpu(), h, w, H, W), rays.cpu(), rtol=1e-5, atol=1e-6)      # the host spelling
