"""Quaternion helpers (mirror of src/modules/warp_utils.py: inv_q :10, mul_q :25, warp_quat_xyz
:78).  Device-agnostic: the reference hard-codes `.cuda()` on its constant tensors (:5,:18-19),
which pins everything to the current device."""
import torch
from torch.autograd import Function

USE_FUSED_QUAT = True   # False: the reference's component-wise formulation through ATen


class _QuatMul(Function):
    """c = a (x) b on (B,1|N,4) f32 CUDA tensors: one launch of csrc/quat.cu; backward is the same kernel twice."""

    @staticmethod
    def _launch(a, b, conj_a, conj_b):
        from .. import _cabi
        B, N = a.shape[0], max(a.shape[1], b.shape[1])
        a = a if a.data_ptr() % 16 == 0 else a.clone()      # the kernel moves quaternions as 128-bit words
        b = b if b.data_ptr() % 16 == 0 else b.clone()
        out = torch.empty(B, N, 4, dtype=torch.float32, device=a.device)
        _cabi.call("i2p_quat_mul", a.device, B, N, a.shape[1], b.shape[1], int(conj_a), int(conj_b),
                   _cabi._ptr(a, torch.float32, "q_a"), _cabi._ptr(b, torch.float32, "q_b", a.device), out.data_ptr())
        return out

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return _QuatMul._launch(a, b, False, False)

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        dc = dc.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _QuatMul._launch(dc, b, False, True)           # dc (x) conj(b)
            if a.shape[1] == 1 and da.shape[1] != 1:
                da = da.sum(dim=1, keepdim=True)
        if ctx.needs_input_grad[1]:
            db = _QuatMul._launch(a, dc, True, False)           # conj(a) (x) dc
            if b.shape[1] == 1 and db.shape[1] != 1:
                db = db.sum(dim=1, keepdim=True)
        return da, db


class _QuatWarp(Function):
    """out = (q (x) [0, p] (x) q^-1)[1:4] + t [, zero where p is all-zero]: csrc/quat.cu quat_warp, one launch per direction.
    p (B,N,3), q (B,4), t (B,3), f32 CUDA."""

    @staticmethod
    def forward(ctx, p, q, t, mask_invalid):
        from .. import _cabi
        f32 = torch.float32
        p, q, t = p.contiguous(), q.contiguous(), t.contiguous()
        if q.data_ptr() % 16:
            q = q.clone()
        B, N = p.shape[0], p.shape[1]
        out = torch.empty(B, N, 3, dtype=f32, device=p.device)
        _cabi.call("i2p_quat_warp_fwd", p.device, B, N, int(mask_invalid), _cabi._ptr(p, f32, "points"), _cabi._ptr(q, f32, "quaternion", p.device),
                   _cabi._ptr(t, f32, "translation", p.device), out.data_ptr())
        ctx.save_for_backward(p, q)
        ctx.mask_invalid = bool(mask_invalid)
        return out

    @staticmethod
    def backward(ctx, g):
        from .. import _cabi
        p, q = ctx.saved_tensors
        f32, dev = torch.float32, p.device
        B, N = p.shape[0], p.shape[1]
        g = g.contiguous()
        need_p, need_q, need_t = ctx.needs_input_grad[:3]
        dp = torch.empty_like(p) if need_p else None
        dq = torch.empty(B, 4, dtype=f32, device=dev) if need_q else None
        dt = torch.empty(B, 3, dtype=f32, device=dev) if need_t else None
        _cabi.call("i2p_quat_warp_bwd", dev, B, N, int(ctx.mask_invalid), p.data_ptr(), q.data_ptr(), _cabi._ptr(g, f32, "gradient", dev),
                   dp.data_ptr() if need_p else None, dq.data_ptr() if need_q else None, dt.data_ptr() if need_t else None)
        return dp, dq, dt, None


def rigid_warp(xyz, quat, trans, mask_invalid=False):
    """xyz (B,N,3), quat (B,4), trans (B,3) -> q [0,p] q^-1 + t, (B,N,3); mask_invalid: all-zero points (empty range-image
    cells) stay zero, the `* check_valid(xyz)` the reference applies after the warp (src/modellearn_proj_center.py:334).
    One kernel on the device, the reference's mul_q / inv_q composition elsewhere."""
    B, N, _ = xyz.shape
    quat, trans = quat.reshape(B, 4), trans.reshape(B, 3)
    if USE_FUSED_QUAT and xyz.is_cuda and xyz.dtype == torch.float32 and quat.dtype == torch.float32 and trans.dtype == torch.float32:
        return _QuatWarp.apply(xyz, quat, trans, mask_invalid)
    out = warp_quat_xyz(xyz, quat, torch.cat([trans.new_zeros(B, 1), trans], -1))
    return out * torch.any(torch.ne(xyz, 0), dim=-1, keepdim=True).float() if mask_invalid else out


def inv_q(q):
    """q (B,1,4) or (B,4) -> conj(q) / (|q|^2 + 1e-10), (B,4)"""
    q = q.reshape(q.shape[0], 4)
    q_2 = torch.sum(q * q, dim=-1, keepdim=True) + 1e-10
    return torch.cat([q[:, :1], -q[:, 1:]], dim=-1) / q_2


def mul_q(q_a, q_b):
    """Hamilton product with broadcasting over the point axis: (B,1|N,4) x (B,1|N,4) -> (B,N,4)"""
    if q_a.ndim == 2:
        q_a = q_a.unsqueeze(1)
    if q_b.ndim == 2:
        q_b = q_b.unsqueeze(1)
    if (USE_FUSED_QUAT and q_a.is_cuda and q_a.dtype == torch.float32 and q_b.dtype == torch.float32
            and q_a.ndim == 3 and q_b.ndim == 3 and q_a.shape[0] == q_b.shape[0]
            and (q_a.shape[1] == q_b.shape[1] or 1 in (q_a.shape[1], q_b.shape[1]))):
        return _QuatMul.apply(q_a, q_b)
    a0, a1, a2, a3 = q_a.unbind(-1)
    b0, b1, b2, b3 = q_b.unbind(-1)
    return torch.stack([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0], dim=-1)


def warp_quat_xyz(lidar_xyz, Hi_quat, H_trans):
    """p' = q [0,p] q^-1 + [0,t] for lidar_xyz (B,N,3), Hi_quat (B,4), H_trans (B,4) -> (B,N,3)"""
    B, N, _ = lidar_xyz.shape
    homo = torch.cat([lidar_xyz.new_zeros(B, N, 1), lidar_xyz], -1)
    homo = mul_q(mul_q(Hi_quat, homo), inv_q(Hi_quat)) + H_trans.reshape(B, 1, 4)
    return homo[:, :, 1:4]


def warp_quat(lidar_xyz, Hi_quat, H_trans, cam_intrinsic, img_shape, LF):
    """Small-range model (warp_utils.py:59-76): lidar_xyz (B,3,N) -> p' = q [0,p] q^-1 + [0,t], then the
    normalised-plane position p' / (z' + 1e-10) (B,N,3), the depth z' (B,N,1) and LF unchanged."""
    homo = warp_quat_xyz(lidar_xyz.permute(0, 2, 1), Hi_quat, H_trans)
    lidar_z = homo[:, :, 2:]
    return homo / (lidar_z + 1e-10), lidar_z, LF


def projection_initial(lidar_xyz, Hi, cam_intrinsic, img_shape, LF):
    """lidar_xyz (B,3,N) -> normalised-plane position p / z (B,N,3), depth (B,N,1), LF (warp_utils.py:146-155)"""
    lidar_xyz = lidar_xyz.permute(0, 2, 1)
    lidar_z = lidar_xyz[:, :, 2:3]
    return lidar_xyz / lidar_z, lidar_z, LF
