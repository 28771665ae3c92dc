"""CPU stand-ins that let the reference's UNCHANGED Python run in a container without a GPU.

TEST INFRASTRUCTURE ONLY (see oracle/i2p_oracle.c).  Used by tests/golden/make_golden.py to
import the real reference from /root/reference and record golden input/output vectors, and by
nothing in i2pnet_b200/.

`install()` registers, before the reference is imported:
  * `fused_conv_select_k_cuda` and `pointnet2.pointnet2_cuda`: the reference's two pybind
    modules (fused_conv_g.cpp:69-73, pointnet2_api.cpp:10-24), here operating on CPU tensors
    through the C oracle;
  * a stub `matplotlib.pyplot` (imported but unused by src/projectPN/utils.py:6);
  * no-op `torch.cuda.synchronize` (src/util/tracker.py:30-32 calls it at config import),
    identity `Tensor.cuda()` and CPU `torch.cuda.FloatTensor/IntTensor/LongTensor`
    constructors (src/modules/warp_utils.py:5,18-19, pointnet2_utils.py:28-29...).
"""
import sys
import types

import numpy as np
import torch

from . import oracle as orc


def _np(t):
    return t.detach().cpu().numpy()


def _fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kH, kW, K, flag, distance, stride_h,
                         stride_w, sel_b, sel_h, sel_w, valid_idx, valid_in_dis_idx, sel_mask, small_h, small_w):
    b, h, w, m = orc.fused_conv_select_k(_np(xyz1), _np(xyz2), _np(idx_n2), _np(random_hw), kH, kW, K, flag,
                                         distance, stride_h, stride_w)
    # the oracle starts from zeros exactly like the caller's pre-zeroed buffers (utils.py:86-94)
    sel_b.copy_(torch.from_numpy(b).view_as(sel_b))
    sel_h.copy_(torch.from_numpy(h).view_as(sel_h))
    sel_w.copy_(torch.from_numpy(w).view_as(sel_w))
    sel_mask.copy_(torch.from_numpy(m).view_as(sel_mask))


def _group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    out.copy_(torch.from_numpy(orc.group_points(_np(points), _np(idx))))
    return 1


def _group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    grad_points.add_(torch.from_numpy(orc.group_points_grad(_np(grad_out), _np(idx), n)))
    return 1


def _gather_points_wrapper(b, c, n, npoints, points, idx, out):
    out.copy_(torch.from_numpy(orc.gather_points(_np(points), _np(idx))))
    return 1


def _gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    grad_points.add_(torch.from_numpy(orc.gather_points_grad(_np(grad_out), _np(idx), n)))
    return 1


def _fps_wrapper(b, n, m, points, temp, idx):
    idx.copy_(torch.from_numpy(orc.furthest_point_sample(_np(points), m)))
    return 1


def _ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    idx.copy_(torch.from_numpy(orc.ball_query(radius, nsample, _np(xyz), _np(new_xyz))))
    return 1


def _three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    d, i = orc.three_nn(_np(unknown), _np(known))
    dist2.copy_(torch.from_numpy(d))
    idx.copy_(torch.from_numpy(i))


def _three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    out.copy_(torch.from_numpy(orc.three_interpolate(_np(points), _np(idx), _np(weight))))


def _three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    grad_points.add_(torch.from_numpy(orc.three_interpolate_grad(_np(grad_out), _np(idx), _np(weight), m)))


def install(reference_root="/root/reference"):
    fused = types.ModuleType("fused_conv_select_k_cuda")
    fused.fused_conv_select_k = _fused_conv_select_k
    sys.modules["fused_conv_select_k_cuda"] = fused

    pn2 = types.ModuleType("pointnet2.pointnet2_cuda")
    for name, fn in dict(group_points_wrapper=_group_points_wrapper,
                         group_points_grad_wrapper=_group_points_grad_wrapper,
                         gather_points_wrapper=_gather_points_wrapper,
                         gather_points_grad_wrapper=_gather_points_grad_wrapper,
                         furthest_point_sampling_wrapper=_fps_wrapper, ball_query_wrapper=_ball_query_wrapper,
                         three_nn_wrapper=_three_nn_wrapper, three_interpolate_wrapper=_three_interpolate_wrapper,
                         three_interpolate_grad_wrapper=_three_interpolate_grad_wrapper).items():
        setattr(pn2, name, fn)
    sys.modules["pointnet2.pointnet2_cuda"] = pn2

    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)

    torch.cuda.synchronize = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = lambda *shape: torch.empty(*shape, dtype=torch.float32)
    torch.cuda.IntTensor = lambda *shape: torch.empty(*shape, dtype=torch.int32)
    torch.cuda.LongTensor = lambda *shape: torch.empty(*shape, dtype=torch.int64)

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import pointnet2  # namespace package of the reference checkout
    pointnet2.pointnet2_cuda = pn2
