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
