"""PointNet++ set abstraction of the small-range model (mirror of the reference's top-level pointnet_util.py:
knn_point :14, square_distance :36, index_points :60, sample_and_group :172, sample_and_group_all :234,
PointNetSetAbstraction :253).

Same constructor, forward signature ((B,C,N) in, (B,C',S) out), return tuple and state_dict keys
(`mlp_convs.i.weight` (C',C,1,1), `mlp_bns.i.*`).  Underneath: furthest point sampling is the cluster kernel
(csrc/fps.cu, start index 0 like the reference's CUDA op), neighbours come from the warp-per-query kNN kernel,
grouping is the channels-last row gather, and the shared MLP + max over the neighbours is the fused kernel chain of
projectPN/fused_mlp.py (tcgen05 for the wide layers) on a (B,S,K,C) tensor -- no (B,S,N) distance matrix, no
NCHW permutes, no materialised activations.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .pointnet2.pointnet2_utils import furthest_point_sample
from .projectPN import PPBackbone_center as _P
from .projectPN.utils import gather_rows, knn_point, square_distance  # noqa: F401  (reference names)


def index_points(points, idx):
    """points (B,N,C), idx (B,S) or (B,S,K) integer -> (B,S,C) / (B,S,K,C)"""
    return gather_rows(points.contiguous(), idx.to(torch.int32))


def farthest_point_sample(xyz, npoint):
    """xyz (B,N,3) -> (B,npoint) int64.  Starts at index 0, like the CUDA op the reference's sample_and_group
    calls (:189); the reference's naive form of this name starts at a random index (:94)."""
    return furthest_point_sample(xyz.contiguous(), npoint).long()


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, sample_idx=None, raw_feat_point=False,
                     raw_xyz=None, feat_mode=None):
    """xyz (B,N,3), points (B,N,D) or None -> new_xyz (B,S,3), new_points (B,S,K,C) [, grouped_xyz (B,S,K,3),
    fps_idx (B,S) int64, new_raw_xyz (B,S,3) or None].  Neighbourhood = the nsample nearest points (`radius` is
    unused, as in the reference :192); feat_mode "dim10feat": [offset, centre, neighbour, |offset|],
    "dist": |offset| alone, None: [offset, neighbour features]."""
    B, N, C = xyz.shape
    S = npoint
    xyz = xyz.contiguous()
    fps_idx = sample_idx if sample_idx is not None else farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = knn_point(nsample, xyz, new_xyz)
    if raw_feat_point:
        new_raw_xyz = index_points(raw_xyz, fps_idx)
        grouped_xyz = index_points(raw_xyz, idx)
        centre = new_raw_xyz.view(B, S, 1, C)
    else:
        new_raw_xyz = None
        grouped_xyz = index_points(xyz, idx)
        centre = new_xyz.view(B, S, 1, C)
    grouped_xyz_norm = grouped_xyz - centre
    if feat_mode == "dim10feat":
        dist = torch.norm(grouped_xyz_norm, p=2, dim=3, keepdim=True)
        new_points = torch.cat([grouped_xyz_norm, centre.expand(-1, -1, nsample, -1), grouped_xyz, dist], -1)
    elif feat_mode == "dist":
        new_points = torch.norm(grouped_xyz_norm, p=2, dim=3, keepdim=True)
    elif points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], -1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx, new_raw_xyz
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """One group holding every point: new_xyz (B,1,3) zeros, new_points (B,1,N,3+D)"""
    B, N, C = xyz.shape
    new_xyz = xyz.new_zeros(B, 1, C)
    grouped_xyz = xyz.view(B, 1, N, C)
    new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1) if points is not None else grouped_xyz
    return new_xyz, new_points


class _Layer:
    """One (nn.Conv2d 1x1, nn.BatchNorm2d, ReLU) triple under the attribute names the fused shared-MLP path reads."""
    bn, activation_fn, leaky_relu = True, True, False

    def __init__(self, conv, norm):
        self.conv, self.bn_linear = conv, norm
        self.in_channels, self.out_channels = conv.in_channels, conv.out_channels

    def _forward_layer(self, x):
        lead = x.shape[:-1]
        y = F.linear(x.reshape(-1, self.in_channels), self.conv.weight.view(self.out_channels, self.in_channels),
                     self.conv.bias)
        return F.relu(_P._batch_norm_rows(y, self.bn_linear)).view(*lead, self.out_channels)


class PointNetSetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel

    def forward(self, xyz, points, sample_idx=None, feat_mode=None, raw_feat_point=False, raw_xyz=None):
        """xyz (B,3,N), points (B,D,N) or None [, raw_xyz (B,N,3)] -> new_xyz (B,3,S), new_points (B,D',S),
        grouped_xyz (B,S,K,3), fps_idx (B,S), new_raw_xyz (B,S,3) or None"""
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        grouped_xyz, fps_idx, new_raw_xyz = [], [], None
        if self.group_all:
            new_xyz, new_points = sample_and_group_all(xyz, points)
        else:
            new_xyz, new_points, grouped_xyz, fps_idx, new_raw_xyz = sample_and_group(
                self.npoint, self.radius, self.nsample, xyz, points, returnfps=True, sample_idx=sample_idx,
                raw_feat_point=raw_feat_point, raw_xyz=raw_xyz, feat_mode=feat_mode)
        layers = [_Layer(c, n) for c, n in zip(self.mlp_convs, self.mlp_bns)]
        new_points = _P.run_mlp(layers, new_points, reduce_k=True)              # (B,S,D'): max over the neighbours
        return new_xyz.permute(0, 2, 1), new_points.permute(0, 2, 1), grouped_xyz, fps_idx, new_raw_xyz if raw_feat_point else None
