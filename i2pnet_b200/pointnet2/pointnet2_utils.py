"""PointNet++ operators with the reference's autograd.Function surface.

Same names, argument order, dtypes and return values as the reference's
pointnet2/pointnet2_utils.py (furthest_point_sample :67, gather_operation :104, three_nn :136,
three_interpolate :184, grouping_operation :228, ball_query :259, QueryAndGroup :262,
GroupAll :298, knn :38) so its callers and tests read the same; underneath, every op is one
launch of a hand-written sm_100a kernel through the C ABI (include/i2p_b200.h).  Outputs are
allocated on the inputs' device (the reference hard-codes the current CUDA device through
torch.cuda.FloatTensor).
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _cabi


def _new(like, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=like.device)


class KNN(Function):
    """k nearest `known` points of every `unknown` point -> (distance, idx int32), ascending.
    The reference declares this op (pointnet2_utils.py:11-37) but never binds its kernel."""

    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor):
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, k), torch.float32)
        idx = _new(unknown, (B, N, k), torch.int32)
        _cabi.knn(B, N, m, k, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B,N,3) -> (B,npoint) int32 indices, starting at index 0 (pointnet2_utils.py:42-60)."""
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = _new(xyz, (B, npoint), torch.int32)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        _cabi.furthest_point_sampling(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint) (pointnet2_utils.py:72-91)."""
        assert features.is_contiguous() and idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, npoint), torch.float32)
        _cabi.gather_points(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _cabi.gather_points_grad(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B,N,3), known (B,M,3) -> (dist (B,N,3) L2, idx (B,N,3) int32) (:110-130)."""
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, 3), torch.float32)
        idx = _new(unknown, (B, N, 3), torch.int32)
        _cabi.three_nn(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B,C,M), idx/weight (B,n,3) -> (B,C,n) (pointnet2_utils.py:142-162)."""
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = _new(features, (B, c, n), torch.float32)
        _cabi.three_interpolate(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
        _cabi.three_interpolate_grad(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample) (:190-209)."""
        assert features.is_contiguous() and idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, nfeatures, nsample), torch.float32)
        _cabi.group_points(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _cabi.group_points_grad(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) int32 (pointnet2_utils.py:233-252)."""
        assert new_xyz.is_contiguous() and xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        _cabi.ball_query(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping (pointnet2_utils.py:262-295): -> (B, 3 + C, npoint, nsample)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class GroupAll(nn.Module):
    """Single group holding every point (pointnet2_utils.py:298-321): -> (B, 3 + C, 1, N)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
