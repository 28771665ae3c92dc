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


class _PoseLoss(torch.autograd.Function):
    """Get_loss as one kernel forward and one backward (csrc/head.cu); the reference formulation is ~47 element-wise /
    reduction launches forward and as many backward on 7-float rows."""

    @staticmethod
    def forward(ctx, out3, out4, q_gt, t_gt, w_x, w_q, l1):
        from . import _cabi
        f32, dev = torch.float32, out3.device
        out3, out4, q_gt, t_gt = out3.contiguous(), out4.contiguous(), q_gt.contiguous(), t_gt.contiguous()
        B = out3.shape[0]
        loss3 = torch.empty(3, dtype=f32, device=dev)
        _cabi.call("i2p_pose_loss_fwd", dev, B, int(l1), _cabi._ptr(out3, f32, "out3", dev), _cabi._ptr(out4, f32, "out4", dev),
                   _cabi._ptr(q_gt, f32, "q_gt", dev), _cabi._ptr(t_gt, f32, "t_gt", dev), _cabi._ptr(w_x, f32, "w_x", dev),
                   _cabi._ptr(w_q, f32, "w_q", dev), loss3.data_ptr())
        ctx.save_for_backward(out3, out4, q_gt, t_gt, w_x, w_q)
        ctx.l1 = int(l1)
        ctx.params = (w_x, w_q)
        ctx.set_materialize_grads(False)
        return loss3[0:1], loss3[1:2], loss3[2:3]

    @staticmethod
    def backward(ctx, dtotal, dreal, ddual):
        from . import _cabi
        from .engine import grad_sink
        if dreal is not None or ddual is not None:
            raise _cabi.I2PError("Get_loss: only the total loss is differentiable here (the two parts are logged metrics)")
        out3, out4, q_gt, t_gt, w_x, w_q = ctx.saved_tensors
        f32, dev = torch.float32, out3.device
        if dtotal is None:
            return (None,) * 7
        d3, d4 = torch.empty_like(out3), torch.empty_like(out4)
        sinks = [grad_sink(p) for p in ctx.params]
        gs = [s if s is not None else torch.empty(1, dtype=f32, device=dev) for s in sinks]
        if sinks[0] is not None or sinks[1] is not None:     # the kernel WRITES d w_x / d w_q: go through temporaries, then add
            tmp = torch.empty(2, dtype=f32, device=dev)
            _cabi.call("i2p_pose_loss_bwd", dev, out3.shape[0], ctx.l1, out3.data_ptr(), out4.data_ptr(), q_gt.data_ptr(),
                       t_gt.data_ptr(), w_x.data_ptr(), w_q.data_ptr(), _cabi._ptr(dtotal.contiguous(), f32, "grad", dev),
                       d3.data_ptr(), d4.data_ptr(), tmp[0:1].data_ptr(), tmp[1:2].data_ptr())
            out = []
            for i, s in enumerate(sinks):
                if s is not None:
                    s.add_(tmp[i:i + 1])
                    out.append(None)
                else:
                    out.append(tmp[i:i + 1])
            return d3, d4, None, None, out[0], out[1], None
        _cabi.call("i2p_pose_loss_bwd", dev, out3.shape[0], ctx.l1, out3.data_ptr(), out4.data_ptr(), q_gt.data_ptr(),
                   t_gt.data_ptr(), w_x.data_ptr(), w_q.data_ptr(), _cabi._ptr(dtotal.contiguous(), f32, "grad", dev),
                   d3.data_ptr(), d4.data_ptr(), gs[0].data_ptr(), gs[1].data_ptr())
        return d3, d4, None, None, gs[0], gs[1], None


def Get_loss(out3, out4, qq_gt, t_gt, w_x, w_q, cfg):
    """out3 refined (B,7), out4 coarse (B,7) -> (total, rotation part, translation part).
    total = 1.6*L(out4) + 0.8*L(out3), L = lx*exp(-w_x) + w_x + lq*exp(-w_q) + w_q."""
    if out3.is_cuda and out3.dtype == torch.float32:
        return _PoseLoss.apply(out3, out4, qq_gt, t_gt, w_x, w_q, bool(cfg.l1_trans_loss))
    q3, x3 = _pose_terms(out3, qq_gt, t_gt, cfg.l1_trans_loss)
    q4, x4 = _pose_terms(out4, qq_gt, t_gt, cfg.l1_trans_loss)
    l3 = x3 * torch.exp(-w_x) + w_x + q3 * torch.exp(-w_q) + w_q
    l4 = x4 * torch.exp(-w_x) + w_x + q4 * torch.exp(-w_q) + w_q
    return 1.6 * l4 + 0.8 * l3, 1.6 * q4 + 0.8 * q3, 1.6 * x4 + 0.8 * x3
