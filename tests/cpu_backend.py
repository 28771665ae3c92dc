"""Test-only stand-ins for i2pnet_b200._cabi entry points, backed by the C oracle (CPU tensors).
Used by the `oracle_backend` fixture to exercise the host-side module / autograd logic where no
GPU exists.  Never imported by the product."""
import numpy as np
import torch

from oracle import oracle as orc


def _np(t):
    return t.detach().cpu().numpy()


def _put(dst, arr):
    dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view_as(dst))


def select_k_flat(xyz1, xyz2, idx_n2, grid, kernel, K, flag, distance, stride, flat_idx, mask):
    B = xyz1.shape[0]
    if idx_n2 is None:
        out_h, out_w, sch, scw = grid
        hh, ww = np.meshgrid(np.arange(out_h) * sch, np.arange(out_w) * scw, indexing="ij")
        idx = np.broadcast_to(np.stack([hh, ww], -1).reshape(1, -1, 2), (B, out_h * out_w, 2)).astype(np.int32)
    else:
        idx = _np(idx_n2)
    total = kernel[0] * kernel[1]
    _, h, w, m = orc.fused_conv_select_k(_np(xyz1), _np(xyz2), idx, np.arange(total, dtype=np.int32), kernel[0],
                                         kernel[1], K, flag, distance, stride[0], stride[1])
    _put(flat_idx, (h * xyz2.shape[2] + w).astype(np.int32))
    _put(mask, m)


def fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kH, kW, K, flag, distance, stride_h, stride_w,
                        sel_b, sel_h, sel_w, sel_mask, small_h, small_w):
    b, h, w, m = orc.fused_conv_select_k(_np(xyz1), _np(xyz2), _np(idx_n2), _np(random_hw), kH, kW, K, flag,
                                         distance, stride_h, stride_w)
    _put(sel_b, b), _put(sel_h, h), _put(sel_w, w), _put(sel_mask, m)


def gather_rows(b, hw, c, m, feature, flat_idx, out):
    _put(out, orc.gather_rows(_np(feature), _np(flat_idx)))


def gather_rows_grad(b, hw, c, m, grad_out, flat_idx, grad_feature):
    grad_feature.add_(torch.from_numpy(orc.gather_rows_grad(_np(grad_out), _np(flat_idx), hw)))


def knn_point(b, n, s, nsample, xyz, new_xyz, group_idx, dist_out=None):
    _put(group_idx, orc.knn(nsample, _np(xyz), _np(new_xyz)))


def project_seq(xyz, feats, H, W, fup, fdown, xyz_proj, feat_projs, owner):
    xp, fps = orc.project_seq(_np(xyz), [_np(f) for f in feats], H, W, fup, fdown)
    _put(xyz_proj, xp)
    for dst, src in zip(feat_projs, fps):
        _put(dst, src)


def group_points(b, c, n, npoints, nsample, points, idx, out):
    _put(out, orc.group_points(_np(points), _np(idx)))


def group_points_grad(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    grad_points.add_(torch.from_numpy(orc.group_points_grad(_np(grad_out), _np(idx), n)))


def furthest_point_sampling(b, n, m, points, temp, idx):
    _put(idx, orc.furthest_point_sample(_np(points), m))


def patch(monkeypatch):
    from i2pnet_b200 import _cabi
    for name in ("select_k_flat", "fused_conv_select_k", "gather_rows", "gather_rows_grad", "knn_point",
                 "project_seq", "group_points", "group_points_grad", "furthest_point_sampling"):
        monkeypatch.setattr(_cabi, name, globals()[name])
    from i2pnet_b200.projectPN import PPBackbone_center as P
    monkeypatch.setattr(P, "_batch_norm_rows", P._batch_norm_rows_stable)   # see its docstring
