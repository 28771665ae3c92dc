"""GPU: the fused pose head and pose loss (csrc/head.cu) against the reference formulation -- PoseHead
(src/projectPN/PPBackbone_center.py:503-560) and Get_loss (compute_loss.py:102-133) through ATen in f64 -- values and
every gradient, with and without dropout (same multipliers), for both head shapes of the model."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_head(pred, mask, w1, b1, wq, bq, wt, bt, drop):
    mask_p = F.softmax(mask, dim=1)
    pooled = torch.sum(pred * mask_p, dim=1)
    hidden = pooled @ w1[:, :, 0].t() + b1
    if drop is not None:
        hidden = hidden * drop
    q = hidden @ wq[:, :, 0].t() + bq
    t = hidden @ wt[:, :, 0].t() + bt
    q = q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10)
    return q, t, mask_p


@pytest.mark.parametrize("shape", [(8, 116, 64, 256), (2, 228, 64, 256), (3, 50, 128, 200), (1, 7, 32, 16)])
@pytest.mark.parametrize("dropout", [False, True])
def test_pose_head_matches_reference_formulation(shape, dropout):
    from i2pnet_b200.projectPN.PPBackbone_center import _PoseHeadFn
    B, N, C, Hd = shape
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    pred, mask = r(B, N, C), r(B, N, C) * 3
    params = [r(Hd, C, 1) * 0.2, r(Hd) * 0.1, r(4, Hd, 1) * 0.1, r(4) * 0.1, r(3, Hd, 1) * 0.1, r(3) * 0.1]
    drop = (torch.rand(B, Hd, device=DEV, generator=g) > 0.5).float() * 2.0 if dropout else None
    gq, gt = r(B, 4), r(B, 3)
    leaves = [t.clone().requires_grad_(True) for t in [pred, mask] + params]
    q, t, mp = _PoseHeadFn.apply(*leaves, drop)
    (q * gq).sum().add((t * gt).sum()).backward()
    ref_leaves = [t.double().clone().requires_grad_(True) for t in [pred, mask] + params]
    qr, tr, mpr = _ref_head(*ref_leaves, drop.double() if dropout else None)
    (qr * gq.double()).sum().add((tr * gt.double()).sum()).backward()
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(q, qr) < 2e-6 and rel(t, tr) < 2e-6 and rel(mp, mpr) < 2e-6
    for name, a, b in zip(["pred", "mask", "w1", "b1", "wq", "bq", "wt", "bt"], leaves, ref_leaves):
        assert rel(a.grad, b.grad) < 2e-5, (name, rel(a.grad, b.grad))


@pytest.mark.parametrize("l1", [True, False])
@pytest.mark.parametrize("B", [1, 8, 200])
def test_pose_loss_matches_reference_formulation(B, l1):
    from i2pnet_b200 import compute_loss

    class Cfg:
        l1_trans_loss = l1
    g = torch.Generator(device=DEV).manual_seed(B)
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    out3, out4, q_gt, t_gt = r(B, 7), r(B, 7), F.normalize(r(B, 4), dim=-1), r(B, 3) * 3
    sx, sq = torch.tensor([0.3], device=DEV), torch.tensor([-2.5], device=DEV)
    leaves = [t.clone().requires_grad_(True) for t in (out3, out4, sx, sq)]
    total, real, dual = compute_loss.Get_loss(leaves[0], leaves[1], q_gt, t_gt, leaves[2], leaves[3], Cfg)
    assert total.shape == (1,)
    total.backward()
    ref_leaves = [t.double().clone().requires_grad_(True) for t in (out3, out4, sx, sq)]
    # the same function on f64 CPU tensors takes the reference formulation
    cpu = [t.detach().cpu().requires_grad_(True) for t in ref_leaves]
    rt, rr, rd = compute_loss.Get_loss(cpu[0], cpu[1], q_gt.double().cpu(), t_gt.double().cpu(), cpu[2], cpu[3], Cfg)
    rt.backward()
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(total, rt.detach()) < 2e-6 and rel(real, rr.detach()) < 2e-6 and rel(dual, rd.detach()) < 2e-6
    for name, a, b in zip(["out3", "out4", "sx", "sq"], leaves, cpu):
        assert rel(a.grad, b.grad) < 1e-5, (name, rel(a.grad, b.grad))
