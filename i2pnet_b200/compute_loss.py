"""Pose loss with learnable balance (mirror of the reference's compute_loss.py:102-133, Get_loss)."""
import torch
import torch.nn.functional as F


def _pose_terms(pred, q_gt, t_gt, l1_trans):
    q, t = pred[:, :4], pred[:, 4:]
    loss_q = torch.mean(torch.sqrt(torch.sum((q_gt - q) * (q_gt - q), dim=-1, keepdim=True) + 1e-10))
    if l1_trans:
        loss_x = F.l1_loss(t, t_gt)
    else:
        loss_x = torch.mean(torch.sqrt(torch.sum((t - t_gt) * (t - t_gt), dim=-1, keepdim=True) + 1e-10))
    return loss_q, loss_x


def Get_loss(out3, out4, qq_gt, t_gt, w_x, w_q, cfg):
    """out3 refined (B,7), out4 coarse (B,7) -> (total, rotation part, translation part).
    total = 1.6*L(out4) + 0.8*L(out3), L = lx*exp(-w_x) + w_x + lq*exp(-w_q) + w_q."""
    q3, x3 = _pose_terms(out3, qq_gt, t_gt, cfg.l1_trans_loss)
    q4, x4 = _pose_terms(out4, qq_gt, t_gt, cfg.l1_trans_loss)
    l3 = x3 * torch.exp(-w_x) + w_x + q3 * torch.exp(-w_q) + w_q
    l4 = x4 * torch.exp(-w_x) + w_x + q4 * torch.exp(-w_q) + w_q
    return 1.6 * l4 + 0.8 * l3, 1.6 * q4 + 0.8 * q3, 1.6 * x4 + 0.8 * x3
