"""Small-range image-to-point-cloud registration network (mirror of the reference's src/modellearn.py:
RegNet_v2 :24-395, set_id_grid :398, change_intrinsic :416).  SURVEY.md section 8 row f3.

Same constructor, forward signature, return tuple and state_dict keys.  The point branch is four furthest-point /
kNN set abstractions on the raw cloud (8192 -> 2048 -> 1024 -> 256 -> 64 points), the cost volumes attend from
the 256 level-3 points over the level-3 image pixels.  Underneath every operator is a kernel of libi2p_b200.so
(FPS cluster kernel, kNN, row gathers, fused shared MLPs, cost-volume build, softmax-weighted sums, quaternion
products); the forward makes no host round-trip (the reference inverts the intrinsic on the CPU, :228), so the
step can be captured into a CUDA graph.
"""
import numpy as np
import torch
import torch.nn as nn

from .config_lidarcenter import I2PNetConfig as cfg_default
from .modellearn_proj_center import change_intrinsic, set_id_grid  # noqa: F401  (same helpers, same names)
from .modules import warp_utils
from .modules.basicConv import createCNNs
from .modules.MainModules import CostVolume, DelayWeight, FlowPredictor, PoseHead, ProjectMask
from .modules.pointnet2_module import SetUpconvModule
from .pointnet_util import PointNetSetAbstraction, farthest_point_sample, index_points
from .projectPN.utils import inverse3x3, pixel_rays  # noqa: F401
from .streams import Fork


class RegNet_v2(nn.Module):
    def __init__(self, bn_decay=None, eval_info=False, cfg=cfg_default):
        super().__init__()
        self.eval_info = eval_info
        npts = [cfg.lidar_in_points // s for s in np.cumprod(cfg.lidar_downsample_rate)]
        enc = cfg.lidar_encoder_mlps
        in_ch = [cfg.lidar_feature_size + 3] + [m[-1] + 3 for m in enc[:3]]
        radii = [0.5, 0.5, 1.0, 2.0]
        for lv in range(4):
            setattr(self, "LiDAR_lv%d" % (lv + 1), PointNetSetAbstraction(
                npoint=npts[lv], radius=radii[lv], nsample=cfg.lidar_group_samples[lv], in_channel=in_ch[lv], mlp=enc[lv],
                group_all=False))
        self.layer_idx = PointNetSetAbstraction(npoint=npts[3], radius=2.0, nsample=cfg.lidar_group_samples[4],
                                                in_channel=cfg.cost_volume_mlps[-1][-1] + 3, mlp=enc[4], group_all=False)
        for i in range(3):
            setattr(self, "RGB_net%d" % (i + 1), createCNNs(*cfg.rgb_encoder_channels[i]))

        def cost_volume(k):
            return CostVolume(radius=10.0, nsample=cfg.cost_volume_nsamples[0], nsample_q=cfg.cost_volume_nsamples[1][k],
                              rgb_in_channels=cfg.rgb_encoder_channels[-1][1][-1], lidar_in_channels=enc[-3][-1],
                              mlp1=cfg.cost_volume_mlps[0], mlp2=cfg.cost_volume_mlps[1], is_training=self.training,
                              bn_decay=bn_decay, bn=True, pooling='max', knn=True, corr_func=cfg.cost_volume_corr_func,
                              backward_validation=cfg.backward_validation[k], max_cost=cfg.max_cost,
                              backward_fc=cfg.backward_fc)
        self.cost_volume1, self.cost_volume2 = cost_volume(0), cost_volume(1)
        fp = cfg.flow_predictor_mlps
        self.flow_predictor0 = FlowPredictor(in_channels=enc[-2][-1] + enc[-1][-1], mlp=fp[0], is_training=self.training,
                                             bn_decay=bn_decay)
        self.set_upconv0_w_upsample = SetUpconvModule(
            nsample=cfg.setupconv_nsamples[0], radius=2.4, in_channels=[enc[-3][-1], fp[0][-1]], mlp=cfg.setupconv_mlps[0][0],
            mlp2=cfg.setupconv_mlps[0][1], is_training=self.training, bn_decay=bn_decay, knn=True)
        self.set_upconv0_upsample = SetUpconvModule(
            nsample=cfg.setupconv_nsamples[1], radius=2.4, in_channels=[enc[-3][-1], enc[-1][-1]], mlp=cfg.setupconv_mlps[1][0],
            mlp2=cfg.setupconv_mlps[1][1], is_training=self.training, bn_decay=bn_decay, knn=True)
        self.flow_predictor0_predict = FlowPredictor(
            in_channels=enc[-3][-1] + cfg.setupconv_mlps[1][1][-1] + cfg.cost_volume_mlps[-1][-1], mlp=fp[1],
            is_training=self.training, bn_decay=bn_decay)
        self.flow_predictor0_w = FlowPredictor(in_channels=enc[-3][-1] + cfg.setupconv_mlps[0][-1][-1] + fp[1][-1], mlp=fp[2],
                                               is_training=self.training, bn_decay=bn_decay)
        head = dict(hidden=cfg.head_hidden_dim, q_dim=cfg.rotation_quat_head_dim, t_dim=cfg.transition_vec_head_dim,
                    dropout_rate=cfg.head_dropout_rate, split_dp=cfg.split_dp, corr_func=cfg.head_corr_func,
                    pos_embed=cfg.head_pos_embedding, sigmoid=cfg.mask_sigmoid, maxhead=cfg.max_head)
        self.l4_head = PoseHead(in_channels=[enc[-1][-1], enc[-2][-1]], mlp1=cfg.pose_head_mlps[0][0],
                                mlp2=cfg.pose_head_mlps[0][1], **head)
        self.l3_head = PoseHead(in_channels=[fp[1][-1], enc[-3][-1]], mlp1=cfg.pose_head_mlps[1][0],
                                mlp2=cfg.pose_head_mlps[1][1], **head)
        if cfg.use_projection_mask:
            if cfg.layer_mask[0]:
                self.l4_projection_mask = ProjectMask(enc[-1][-1] + enc[-2][-1], cfg.projection_mask_mlps[0], cfg.mask_sigmoid)
                self.l4_delay = DelayWeight(cfg.mask_delay_step, cfg.mask_delay, cfg.ab_delay)
            if cfg.layer_mask[1]:
                self.l3_projection_mask = ProjectMask(enc[-3][-1] + fp[1][-1], cfg.projection_mask_mlps[1], cfg.mask_sigmoid)
                self.l3_delay = DelayWeight(cfg.mask_delay_step, cfg.mask_delay, cfg.ab_delay)
        # learnable loss weights
        self.sq = nn.Parameter(torch.tensor([cfg.sq_init]), requires_grad=True)
        self.sx = nn.Parameter(torch.tensor([cfg.sx_init]), requires_grad=True)

    def forward(self, rgb_img, lidar_img, H_initial, intrinsic, resize_img, gt_project=None, calib=None, lidar_feature=None,
                cfg=cfg_default, lidar_img_raw=None):
        """rgb_img (B,3,h,w), lidar_img (B,N,3) in the camera frame, intrinsic (B,3,3), lidar_feature (B,N,D) or None,
        lidar_img_raw (B,N,3) raw sensor coordinates (cfg.raw_feat_point)
        -> out_3 (B,7) = (q, t) composed over both levels, result_4 (B,7), pm3, pm4, sx, sq"""
        device = rgb_img.device
        intrinsic = intrinsic.float()
        B = rgb_img.shape[0]
        N = lidar_img.shape[1]
        with Fork(rgb_img) as rgb_branch:          # image pyramid beside the (sequential, FPS-bound) point pyramid: streams.py
            RF3 = self.RGB_net3(self.RGB_net2(self.RGB_net1(rgb_img)))

        lidar_img = lidar_img.permute(0, 2, 1).float()
        if lidar_feature is None:
            lidar_norm = torch.zeros(B, 3, N, device=device)
        else:
            lidar_norm = lidar_feature.permute(0, 2, 1).float()
        raw = dict(raw_feat_point=cfg.raw_feat_point)
        lv1 = dict(feat_mode=cfg.featmode) if cfg.featmode is not None else {}
        # Furthest point sampling is a chain of M dependent steps per level (~1 us each) and depends on coordinates
        # only: the four samplings run back to back on a side stream, one level ahead of the grouping + shared MLP of the
        # previous level on the current stream (same indices as sampling inside each level, :183 of pointnet_util.py).
        levels = [self.LiDAR_lv1, self.LiDAR_lv2, self.LiDAR_lv3, self.LiDAR_lv4]
        xyz0 = lidar_img.permute(0, 2, 1).contiguous()
        fps = Fork(xyz0)

        def sample(xyz, npoint):
            with fps:
                idx = farthest_point_sample(xyz, npoint)
                return idx, index_points(xyz, idx)
        idx_next, xyz_next = fps.join(*sample(xyz0, levels[0].npoint))
        xyz_l, feat_l, raw_l, outs = lidar_img, lidar_norm, lidar_img_raw, []
        for lv, level in enumerate(levels):
            idx_cur = idx_next
            if lv + 1 < len(levels):
                ahead = sample(xyz_next, levels[lv + 1].npoint)              # issued before this level's own work
            xyz_l, feat_l, _, _, raw_l = level(xyz_l, feat_l, sample_idx=idx_cur, raw_xyz=raw_l, **(lv1 if lv == 0 else {}), **raw)
            outs.append((xyz_l, feat_l, idx_cur, raw_l))
            if lv + 1 < len(levels):
                idx_next, xyz_next = fps.join(*ahead)
        (P1, LF1, fps_idx_1, P1_raw), (P2, LF2, fps_idx_2, P2_raw), (P3, LF3, fps_idx_3, P3_raw), (P4, LF4, fps_idx_4, P4_raw) = outs

        # pixels on the normalised camera plane
        RF3 = rgb_branch.join(RF3)
        RF3_index = pixel_rays(intrinsic, RF3.shape[2], RF3.shape[3], rgb_img.shape[2], rgb_img.shape[3])   # B,h3*w3,3
        lidar_uv, lidar_z, LF3 = warp_utils.projection_initial(P3, None, None, None, LF3)
        _, C, H, W = RF3.shape
        RF3 = RF3.reshape(B, C, H * W).permute(0, 2, 1)                   # B,h3*w3,C
        P3_t, P4_t, LF3_t, LF4_t = P3.permute(0, 2, 1), P4.permute(0, 2, 1), LF3.permute(0, 2, 1), LF4.permute(0, 2, 1)

        # level 4
        concat_4 = self.cost_volume1(lidar_uv, LF3_t, RF3_index, RF3, lidar_z)
        P4, l4_cv, _, _, _ = self.layer_idx(P3, concat_4.permute(0, 2, 1), sample_idx=fps_idx_4, raw_xyz=P3_raw, **raw)
        P4_t = P4.permute(0, 2, 1)
        l4_points_predict = l4_cv.permute(0, 2, 1)
        l4_cost_volume_w = self.flow_predictor0(LF4_t, None, l4_points_predict)
        l4_projection_mask = None
        if cfg.use_projection_mask and cfg.layer_mask[0]:
            l4_projection_mask = self.l4_projection_mask(LF4_t, l4_points_predict)
        if gt_project is not None:
            gt_project_l1 = index_points(gt_project, fps_idx_1)
            gt_project_l2 = index_points(gt_project_l1, fps_idx_2)
            gt_project_l3 = index_points(gt_project_l2, fps_idx_3)
            gt_project_l4 = index_points(gt_project_l3, fps_idx_4)
            if cfg.ground_truth_mask_layer[0]:
                l4_projection_mask_predict = l4_projection_mask
                l4_projection_mask = (self.l4_delay(gt_project_l4, l4_projection_mask) if l4_projection_mask is not None
                                      else gt_project_l4)
        if cfg.ground_truth_mask_layer[0]:
            assert l4_projection_mask is not None
        q4, t4, _ = self.l4_head(l4_points_predict, l4_cost_volume_w, P4_t, LF4_t, l4_projection_mask)
        if gt_project is not None and cfg.ground_truth_mask_layer[0]:
            l4_projection_mask = l4_projection_mask_predict
        result_4 = torch.cat([q4, t4], dim=1)

        # level 3: warp the level-3 points with the level-4 estimate, second cost volume, refinement
        t4_quat = torch.cat([torch.zeros((B, 1), device=device), t4], -1)
        lidar_uv, lidar_z, LF3 = warp_utils.warp_quat(P3, q4, t4_quat, None, None, LF3)
        concat_3 = self.cost_volume2(lidar_uv, LF3_t, RF3_index, RF3, lidar_z)
        up = dict(raw_feat_point=True, raw_xyz1=P3_raw, raw_xyz2=P4_raw) if cfg.raw_feat_point else {}
        with Fork(P3_t, P4_t, LF3_t, l4_cost_volume_w, P3_raw, P4_raw) as up_branch:     # independent up-convolutions
            l3_cost_volume_w_upsample = self.set_upconv0_w_upsample(P3_t, P4_t, LF3_t, l4_cost_volume_w, **up)
        l3_cost_volume_upsample = self.set_upconv0_upsample(P3_t, P4_t, LF3_t, l4_points_predict, **up)
        l3_cost_volume_w_upsample = up_branch.join(l3_cost_volume_w_upsample)
        l3_cost_volume_predict = self.flow_predictor0_predict(LF3_t, l3_cost_volume_upsample, concat_3)
        l3_cost_volume_w = self.flow_predictor0_w(LF3_t, l3_cost_volume_w_upsample, l3_cost_volume_predict)
        l3_prediction_mask = None
        if cfg.use_projection_mask and cfg.layer_mask[1]:
            l3_prediction_mask = self.l3_projection_mask(LF3_t, l3_cost_volume_predict)
        if gt_project is not None and cfg.ground_truth_mask_layer[1]:
            l3_prediction_mask_predict = l3_prediction_mask
            l3_prediction_mask = (self.l3_delay(gt_project_l3, l3_prediction_mask) if l3_prediction_mask is not None
                                  else gt_project_l3)
        q3, t3, W_l3_cost_volume = self.l3_head(l3_cost_volume_predict, l3_cost_volume_w, P3_t, LF3_t, l3_prediction_mask)
        if gt_project is not None and cfg.ground_truth_mask_layer[1]:
            l3_prediction_mask = l3_prediction_mask_predict

        # compose: q = q3 q4, t = q3 [0,t4] q3^-1 + t3
        out_3_real = warp_utils.mul_q(q3.view(B, 1, 4), q4.view(B, 1, 4))
        t3_quat = torch.cat([torch.zeros((B, 1), device=device), t3], 1).view(B, 1, 4)
        out_3_dual = warp_utils.mul_q(warp_utils.mul_q(q3, t4_quat.view(B, 1, 4)), warp_utils.inv_q(q3)) + t3_quat
        out_3 = torch.cat((out_3_real.squeeze(1), out_3_dual.squeeze(1)[:, 1:]), 1)

        if self.eval_info:
            if gt_project is not None:
                return out_3.float(), result_4.float(), self.sx, self.sq, W_l3_cost_volume, P3_t, gt_project_l3, gt_project_l4, P4_t
            return out_3.float(), result_4.float(), self.sx, self.sq, W_l3_cost_volume, P3_t, l3_prediction_mask, l4_projection_mask, P4_t
        pm3 = [l3_prediction_mask, P3_t] if l3_prediction_mask is not None else None
        if gt_project is not None and pm3 is not None:
            pm3.append(gt_project_l3)
        pm4 = [l4_projection_mask, P4_t] if l4_projection_mask is not None and not cfg.one_head_mask else None
        if gt_project is not None and pm4 is not None:
            pm4.append(gt_project_l4)
        return out_3.float(), result_4.float(), pm3, pm4, self.sx, self.sq
