"""Building blocks of the small-range model (mirror of src/modules/MainModules.py: FlowPredictor :10, CostVolume :51,
PoseHead :245, ProjectMask :395, DelayWeight :433).

Constructor arguments, forward signatures, return values and state_dict keys are the reference's.  The cost
volume shares its kernels with the large-range one (projectPN/PPBackbone_center.py): kNN kernel instead of a
distance matrix + topk, row gathers, the first-layer operand built in one kernel from the per-point / per-pixel
standardised features (never the `repeat`ed (B,N,K,C) operands), fused shared-MLP chains, softmax-weighted sums as
one kernel each way.
"""
from enum import Enum

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..projectPN import PPBackbone_center as _P
from ..projectPN.PPBackbone_center import Conv2d as _Conv2d
from ..projectPN.PPBackbone_center import run_mlp
from ..projectPN.utils import gather_rows, knn_point
from .basicConv import Conv1d
from .point_utils import grouping


def Conv2d(in_channels, out_channels, kernel_size, stride=None, bn=False, activation_fn=True, leaky_relu=True):
    """src/modules/basicConv.py:23-59: the norm is a plain nn.BatchNorm2d (running statistics tracked)."""
    return _Conv2d(in_channels, out_channels, kernel_size, stride, bn=bn, activation_fn=activation_fn,
                   leaky_relu=leaky_relu, use_bn_input=False)


class FlowPredictor(nn.Module):
    """Per-point MLP over [points_f1, cost_volume, upsampled_feat] (B,n,c1+c3+c2) -> (B,n,mlp[-1])"""

    def __init__(self, in_channels, mlp, is_training, bn_decay, bn=True):
        super().__init__()
        self.in_channels, self.mlp, self.is_training, self.bn_decay, self.bn = in_channels, mlp, is_training, bn_decay, bn
        self.mlp_conv = nn.ModuleList()
        for c in mlp:
            self.mlp_conv.append(Conv2d(self.in_channels, c, [1, 1], stride=[1, 1], bn=bn))
            self.in_channels = c

    def forward(self, points_f1, upsampled_feat, cost_volume):
        parts = [points_f1, cost_volume] + ([upsampled_feat] if upsampled_feat is not None else [])
        return run_mlp(self.mlp_conv, torch.cat(parts, -1).unsqueeze(2)).squeeze(2)


class CostVolume(nn.Module):
    """Every LiDAR point attends over image pixels (all of them: nsample_q <= 0, or its nsample_q nearest on the
    normalised plane), then over its nsample nearest 3-D neighbours."""

    class CorrFunc(Enum):
        ELEMENTWISE_PRODUCT = 1
        CONCAT = 2
        COSINE_DISTANCE = 3

    def __init__(self, radius, nsample, nsample_q, rgb_in_channels, lidar_in_channels, mlp1, mlp2, is_training, bn_decay,
                 bn=True, pooling='max', knn=True, corr_func=CorrFunc.ELEMENTWISE_PRODUCT, backward_validation=False,
                 max_cost=False, backward_fc=False):
        super().__init__()
        self.radius, self.nsample, self.nsample_q, self.mlp1, self.mlp2 = radius, nsample, nsample_q, mlp1, mlp2
        self.is_training, self.bn_decay, self.bn, self.pooling, self.knn = is_training, bn_decay, bn, pooling, knn
        self.corr_func, self.backward_validation, self.max_cost, self.backward_fc = corr_func, backward_validation, max_cost, backward_fc
        if corr_func == CostVolume.CorrFunc.CONCAT:
            corr_channel = rgb_in_channels + lidar_in_channels
        elif corr_func in (CostVolume.CorrFunc.ELEMENTWISE_PRODUCT, CostVolume.CorrFunc.COSINE_DISTANCE):
            corr_channel = rgb_in_channels
        else:
            raise NotImplementedError
        if backward_validation:
            corr_channel += lidar_in_channels
        if backward_fc and backward_validation:
            self.inverse_fc = Conv2d(lidar_in_channels, lidar_in_channels, [1, 1], [1, 1], bn=True)
        self.in_channels = corr_channel + 6
        self.mlp1_convs = nn.ModuleList()
        if not max_cost:
            self.mlp2_convs = nn.ModuleList()
        self.mlp2_convs_2 = nn.ModuleList()
        for c in mlp1:
            self.mlp1_convs.append(Conv2d(self.in_channels, c, [1, 1], stride=[1, 1], bn=True))
            self.in_channels = c
        self.pi_encoding = Conv2d(6, mlp1[-1], [1, 1], stride=[1, 1], bn=True)
        if not max_cost:
            self.in_channels = 2 * mlp1[-1]
            for c in mlp2:
                self.mlp2_convs.append(Conv2d(self.in_channels, c, [1, 1], stride=[1, 1], bn=True))
                self.in_channels = c
        self.pc_encoding = Conv2d(10, mlp1[-1], [1, 1], stride=[1, 1], bn=True)
        self.in_channels = 2 * mlp1[-1] + lidar_in_channels
        for c in mlp2:
            self.mlp2_convs_2.append(Conv2d(self.in_channels, c, [1, 1], stride=[1, 1], bn=True))
            self.in_channels = c

    def _fusable_first_layer(self, x):
        return (_P.USE_FUSED_CV and x.is_cuda and self.corr_func == CostVolume.CorrFunc.ELEMENTWISE_PRODUCT
                and not (self.backward_validation and (self.backward_fc or self.nsample_q > 0)))

    def _first_layer_operand(self, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
        """The reference's formulation (:133-185) on broadcast views.
        -> operand of mlp1 (B,N,K,6+c), coordinate pairs (B,N,K,6), depth-restored xyz (B,N,3)"""
        N = warped_points.shape[1]
        if self.nsample_q > 0:
            idx = knn_point(self.nsample_q, f2_xyz.contiguous(), warped_xyz.contiguous()).to(torch.int32)
            qi_xyz, qi_points = gather_rows(f2_xyz.contiguous(), idx), gather_rows(f2_points.contiguous(), idx)
        else:
            qi_xyz = f2_xyz.unsqueeze(1).expand(-1, N, -1, -1)
            qi_points = f2_points.unsqueeze(1).expand(-1, N, -1, -1)
        K = qi_xyz.shape[2]
        warped_xyz = warped_xyz.mul(lidar_z)                               # restore depth (:146)
        xyz6 = torch.cat([warped_xyz[:, :, None, :].expand(-1, -1, K, -1), qi_xyz], dim=3)
        pi = warped_points[:, :, None, :].expand(-1, -1, K, -1)
        if self.corr_func == CostVolume.CorrFunc.ELEMENTWISE_PRODUCT:
            pi, qi_points = _P._standardise(pi), _P._standardise(qi_points)
            corr = pi * qi_points
        elif self.corr_func == CostVolume.CorrFunc.CONCAT:
            corr = torch.cat([pi, qi_points], dim=-1)
        else:
            pi, qi_points = F.normalize(pi, p=2, dim=-1, eps=1e-12), F.normalize(qi_points, p=2, dim=-1, eps=1e-12)
            corr = pi * qi_points
        parts = [xyz6, corr]
        if self.backward_validation:                                       # strongest response of every pixel over the points
            respond = torch.max(qi_points * pi, 1, keepdim=True)[0].expand(-1, N, -1, -1)
            if self.backward_fc:
                respond = self.inverse_fc(respond.contiguous())
            parts.append(respond)
        return torch.cat(parts, dim=3), xyz6, warped_xyz

    def _first_layer_operand_fused(self, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
        """Same values from the build kernel (csrc/cv.cu): points and pixels are standardised once each; the maximum
        over the points of pi[n,c] * qi[k,c] is qi[k,c] times the largest (qi > 0) or smallest (qi < 0) pi[:,c]."""
        idx = None
        if self.nsample_q > 0:
            idx = knn_point(self.nsample_q, f2_xyz.contiguous(), warped_xyz.contiguous()).to(torch.int32)
        warped_xyz = warped_xyz.mul(lidar_z)
        pi_n, qi_n = _P._standardise(warped_points), _P._standardise(f2_points)
        maxc = None
        if self.backward_validation:
            hi, lo = pi_n.max(dim=1, keepdim=True)[0], pi_n.min(dim=1, keepdim=True)[0]
            maxc = torch.where(qi_n > 0, qi_n * hi, qi_n * lo)
        X, xyz6 = _P._CvBuild.apply(warped_xyz, f2_xyz, pi_n, qi_n, maxc, idx)
        return X, xyz6, warped_xyz

    def forward(self, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
        """warped_xyz (B,N,3) on the normalised plane, warped_points (B,N,c), f2_xyz (B,N2,3), f2_points (B,N2,c),
        lidar_z (B,N,1) -> (B,N,mlp2[-1])"""
        build = self._first_layer_operand_fused if self._fusable_first_layer(warped_points) else self._first_layer_operand
        pi_feat1_new, xyz6, warped_xyz = build(warped_xyz, warped_points, f2_xyz, f2_points, lidar_z)
        pi_feat1_new = run_mlp(self.mlp1_convs, pi_feat1_new)
        if not self.max_cost:
            pi_concat = run_mlp(self.mlp2_convs, torch.cat([self.pi_encoding(xyz6), pi_feat1_new], dim=3))
            pi_feat1_new = _P._softmax_wsum(pi_concat, pi_feat1_new)      # B,N,mlp1[-1]
        else:
            pi_feat1_new = torch.max(pi_feat1_new, dim=2)[0]

        # second stage: re-weight over the nsample nearest 3-D neighbours
        pc_xyz_grouped, _, pc_points_grouped, _, _ = grouping(pi_feat1_new, self.nsample, warped_xyz, warped_xyz)
        pc_xyz_new = warped_xyz[:, :, None, :].expand(-1, -1, self.nsample, -1)
        pc_points_new = warped_points[:, :, None, :].expand(-1, -1, self.nsample, -1)
        pc_xyz_diff = pc_xyz_grouped - pc_xyz_new
        pc_euc_diff = torch.sqrt(torch.sum(pc_xyz_diff * pc_xyz_diff, dim=3, keepdim=True) + 1e-20)
        pc_xyz_encoding = self.pc_encoding(torch.cat([pc_xyz_new, pc_xyz_grouped, pc_xyz_diff, pc_euc_diff], dim=3))
        pc_concat = run_mlp(self.mlp2_convs_2, torch.cat([pc_xyz_encoding, pc_points_new, pc_points_grouped], dim=-1))
        return _P._softmax_wsum(pc_concat, pc_points_grouped)


class PoseHead(nn.Module):
    """Mask-weighted pooling over the points [-> global-attention refinement] -> hidden -> (unit quaternion, translation)"""

    class CorrFunc(Enum):
        DIFF = 1
        CONCAT = 2
        NORMALIZED_DIFF = 3

    def __init__(self, in_channels, mlp1, mlp2, hidden, q_dim, t_dim, dropout_rate=0.5, split_dp=False,
                 corr_func=CorrFunc.CONCAT, pos_embed=False, sigmoid=False, maxhead=False):
        super().__init__()
        self.corr_func, self.sigmoid, self.maxhead, self.pos_embed = corr_func, sigmoid, maxhead, pos_embed
        in_channel, l_feature_channel = in_channels
        if pos_embed:
            self.pos_encoder = Conv1d(3 + 3, in_channel, bn=True)
        self.mlps = nn.ModuleList()
        if corr_func == PoseHead.CorrFunc.CONCAT:
            last_dim = 2 * in_channel
        elif corr_func in (PoseHead.CorrFunc.DIFF, PoseHead.CorrFunc.NORMALIZED_DIFF):
            last_dim = in_channel
        else:
            raise NotImplementedError
        if pos_embed:
            last_dim += in_channel
        for c in mlp1:
            self.mlps.append(Conv1d(last_dim, c, bn=True))
            last_dim = c
        if len(mlp1) > 0:
            self.mlp2s = nn.ModuleList()
            last_dim = in_channel + mlp1[-1] + l_feature_channel
            for c in mlp2:
                self.mlp2s.append(Conv1d(last_dim, c, bn=True))
                last_dim = c
        self.DP1 = nn.Identity() if split_dp else nn.Dropout(dropout_rate)
        self.DP2 = nn.Dropout(dropout_rate) if split_dp else nn.Identity()
        self.hidden_layer = Conv1d(in_channel, hidden, use_activation=False)
        self.quat_head = Conv1d(hidden, q_dim, use_activation=False)
        self.trans_head = Conv1d(hidden, t_dim, use_activation=False)

    def forward(self, prediction, mask, xyz, feature, projection_mask):
        """prediction, mask (B,N,C), xyz (B,N,3), feature (B,N,C') -> q (B,4), t (B,3), mask_p (B,N,C)"""
        B, N, _ = prediction.shape
        if not self.sigmoid:
            if projection_mask is not None:
                projection_mask = torch.argmax(projection_mask.detach(), dim=-1, keepdim=True).float()
                mask = mask * projection_mask + -1e10 * (1. - projection_mask)
        else:
            prediction = prediction * projection_mask
        if self.maxhead:
            mask = torch.max(mask, dim=-1, keepdim=True)[0]
        mask_p = F.softmax(mask, dim=1)
        pooled = torch.sum(prediction * mask_p, dim=1, keepdim=True)       # B,1,C
        if len(self.mlps) > 0:
            if self.corr_func == PoseHead.CorrFunc.CONCAT:
                glob = torch.cat([prediction, pooled.expand(-1, N, -1)], dim=-1)
            elif self.corr_func == PoseHead.CorrFunc.DIFF:
                glob = prediction - pooled
            else:
                def norm(v):
                    return (v - v.mean(dim=-1, keepdim=True)) / (v.std(dim=-1, keepdim=True) + 1e-10)
                glob = norm(prediction) * norm(pooled)
            if self.pos_embed:
                pos_info = torch.cat([xyz, xyz - torch.mean(xyz, dim=1, keepdim=True)], dim=-1)
                glob = torch.cat([glob, self.pos_encoder(pos_info)], -1)
            for m in self.mlps:
                glob = m(glob)
            if len(self.mlp2s) > 0:
                glob = torch.cat([glob, mask, feature], dim=-1)
                for m in self.mlp2s:
                    glob = m(glob)
            result = torch.sum(torch.softmax(glob, dim=1) * prediction, dim=1, keepdim=True)
        else:
            result = pooled
        hidden = self.DP1(self.hidden_layer(result))
        q = self.quat_head(self.DP2(hidden)).squeeze(1)
        t = self.trans_head(self.DP2(hidden)).squeeze(1)
        q = q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10)
        return q, t, mask_p


class ProjectMask(nn.Module):
    """[feature, prediction] (B,N,C) -> in-image logits (B,N,2), or a probability (B,N,1) with sigmoid"""

    def __init__(self, in_channel, mlp, sigmoid=False, drop=0):
        super().__init__()
        self.mlps = nn.ModuleList()
        self.drop = nn.Dropout(p=drop) if drop > 0 else nn.Identity()
        last_dim = in_channel
        for c in mlp:
            self.mlps.append(Conv1d(last_dim, c, bn=True))
            last_dim = c
        self.out = Conv1d(mlp[-1], 1 if sigmoid else 2, use_activation=False)
        self.out_act = nn.Sigmoid() if sigmoid else nn.Identity()

    def forward(self, feature, prediction):
        x = torch.cat([feature, prediction], dim=-1) if feature is not None else prediction
        for m in self.mlps:
            x = self.drop(m(x))
        return self.out_act(self.out(x))


class DelayWeight(nn.Module):
    """Blend of the ground-truth and the predicted projection mask that moves to the prediction over `delay_step`
    training steps (`now_step` counts down in a non-trainable parameter, so it is checkpointed)."""

    def __init__(self, delay_step, delay, ab_delay):
        super().__init__()
        self.delay, self.ab_delay, self.delay_step = delay, ab_delay, delay_step
        self.now_step = nn.Parameter(torch.Tensor([delay_step]), requires_grad=False)

    def forward(self, gt, pred):
        if gt is None:
            return pred
        if pred is None:
            return gt
        if self.ab_delay:
            if torch.eq(self.now_step, 0).item():
                return pred
            if self.training:
                self.now_step.add_(-1).clip_(0)
            return gt
        weight = self.now_step.item() / (self.delay_step + 1e-10)
        mixed = gt * weight + F.softmax(pred, dim=-1) * (1 - weight)
        if self.training and self.delay:
            self.now_step.add_(-1).clip_(0)
        return mixed
