"""GPU: the fused Hamilton product (csrc/quat.cu) behind warp_utils.mul_q / warp_quat_xyz against the
reference's component-wise formulation (src/modules/warp_utils.py:25-94), values and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shapes", [((3, 1, 4), (3, 228, 4)), ((3, 228, 4), (3, 1, 4)), ((2, 57, 4), (2, 57, 4)),
                                    ((4, 4), (4, 4))], ids=["bcast_a", "bcast_b", "full", "pose"])
def test_mul_q_matches_reference_formulation(shapes):
    from i2pnet_b200 import _cabi
    from i2pnet_b200.modules import warp_utils as W
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    res = {}
    a0, b0 = torch.randn(*shapes[0], device=dev), torch.randn(*shapes[1], device=dev)
    g = None
    for fused in (True, False):
        W.USE_FUSED_QUAT = fused
        try:
            a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            before = _cabi.launch_count()
            c = W.mul_q(a, b)
            if g is None:
                g = torch.randn_like(c)
            c.backward(g)
            launches = _cabi.launch_count() - before
        finally:
            W.USE_FUSED_QUAT = True
        res[fused] = (c.detach(), a.grad, b.grad, launches)
    assert res[True][3] == 3 and res[False][3] == 0
    assert torch.equal(res[True][0], res[False][0])                       # same operation order: bit-exact forward
    for i in (1, 2):
        assert torch.allclose(res[True][i], res[False][i], rtol=1e-5, atol=1e-5)


def test_warp_quat_xyz_matches_reference_formulation():
    from i2pnet_b200.modules import warp_utils as W
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    out = {}
    p0, q0, t0 = torch.randn(2, 228, 3, device=dev) * 10, torch.randn(2, 4, device=dev), torch.randn(2, 4, device=dev)
    for fused in (True, False):
        W.USE_FUSED_QUAT = fused
        try:
            p, q, t = (v.clone().requires_grad_(True) for v in (p0, q0, t0))
            w = W.warp_quat_xyz(p, q / q.norm(dim=-1, keepdim=True), t)
            w.square().sum().backward()
        finally:
            W.USE_FUSED_QUAT = True
        out[fused] = (w.detach(), p.grad, q.grad, t.grad)
    for x, y in zip(out[True], out[False]):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-4 * float(y.abs().max()))


@pytest.mark.parametrize("shape", [(2, 1824), (3, 1), (1, 333), (8, 7296)], ids=lambda s: "b%d_n%d" % s)
@pytest.mark.parametrize("mask", [False, True], ids=["plain", "masked"])
def test_rigid_warp_one_kernel_per_direction(shape, mask):
    """csrc/quat.cu quat_warp (warp_quat_xyz and the pose composition t = R(q3) t_in + t3 as one launch per direction)
    against the reference's mul_q / inv_q composition through ATen in f32 and f64: output and input gradients; non-unit quaternions (q^-1 = conj(q) / (|q|^2 + 1e-10) is not the
    conjugate), all-zero points under the mask."""
    from i2pnet_b200 import _cabi
    from i2pnet_b200.modules import warp_utils as W
    B, N = shape
    dev = torch.device("cuda:0")
    g_ = torch.Generator(device=dev).manual_seed(B * 7 + N)
    p0 = torch.randn(B, N, 3, device=dev, generator=g_) * 10
    p0[:, ::5] = 0
    q0 = torch.randn(B, 4, device=dev, generator=g_) * 1.3
    t0 = torch.randn(B, 3, device=dev, generator=g_) * 2
    gout = torch.randn(B, N, 3, device=dev, generator=g_)
    res = {}
    for mode, dt in (("own", torch.float32), ("aten", torch.float32), ("f64", torch.float64)):
        W.USE_FUSED_QUAT = mode == "own"
        try:
            p, q, t = (v.clone().to(dt).requires_grad_(True) for v in (p0, q0, t0))
            before = _cabi.launch_count()
            out = W.rigid_warp(p, q, t, mask_invalid=mask)
            out.backward(gout.to(dt))
            launches = _cabi.launch_count() - before
        finally:
            W.USE_FUSED_QUAT = True
        res[mode] = (out.detach().double(), p.grad.double(), q.grad.double(), t.grad.double(), launches)
    assert res["own"][4] == 2 and res["aten"][4] == 0
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
    # (the products follow the operation order of mul_q, but ATen's sum of the four squares in |q|^2 changes its order
    # with the batch size, so the forward is held to the f64 composition like the gradients, not to bit equality)
    for i, name in ((0, "out"), (1, "d points"), (2, "d quaternion"), (3, "d translation")):
        assert rel(res["own"][i], res["f64"][i]) < max(2e-6, 2 * rel(res["aten"][i], res["f64"][i])), (name, rel(res["own"][i], res["f64"][i]))
