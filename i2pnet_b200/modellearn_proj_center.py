"""Large-range image-to-point-cloud registration network (mirror of the reference's
src/modellearn_proj_center.py: RegNet_v2 :24-440, set_id_grid :441, change_intrinsic :459).

Same constructor, forward signature, return tuple and state_dict keys; the reference's own
file also runs unchanged on this package through the `dropin/` extension modules.  This
assembly exists because the reference file forces one host round-trip per forward
(`torch.inverse(intrinsic_3.cpu())`, :282) and a Python loop of index_put_ (project_seq), both
of which keep the step from being captured into a CUDA graph; here the whole forward is
asynchronous on the current stream.
"""
import numpy as np
import torch
import torch.nn as nn

from .config_proj_lidarcenter import I2PNetConfig as cfg_default
from .modules import warp_utils
from .modules.basicConv import createCNNs
from .projectPN.PPBackbone_center import CostVolume, FlowPredictor, PoseHead, ProjectPointNet, ProjSetUpconvModule
from .modules.basicConv import pyramid_out_hw
from .projectPN.utils import StrideGrid, check_valid, inverse3x3, pixel_rays, project_seq  # noqa: F401
from .streams import Fork


def set_id_grid(rf):
    """rf (B,h,w,C) -> homogeneous pixel coordinates (u=col, v=row, 1), (B, h*w, 3)"""
    B, h, w, _ = rf.shape
    v, u = torch.meshgrid(torch.arange(h, device=rf.device, dtype=rf.dtype),
                          torch.arange(w, device=rf.device, dtype=rf.dtype), indexing="ij")
    coords = torch.stack([u, v, torch.ones_like(u)], dim=-1).reshape(1, h * w, 3)
    return coords.expand(B, -1, -1)


def change_intrinsic(intrinsic, RF, rgb_img):
    """Rescale K (B,3,3) from the input image to the feature map RF (B,C,h,w)."""
    sx, sy = RF.shape[3] / rgb_img.shape[3], RF.shape[2] / rgb_img.shape[2]
    # row scaling only (no host tensor: the forward must stay capturable into a CUDA graph);
    # K[0,1] and K[1,0] are zero for a pinhole intrinsic, as the reference assumes too
    return torch.cat([intrinsic[:, 0:1] * sx, intrinsic[:, 1:2] * sy, intrinsic[:, 2:3]], dim=1)


def _mask_fill(x, valid):
    return x * valid + -1e10 * (1 - valid)


class RegNet_v2(nn.Module):
    def __init__(self, bn_decay=None, eval_info=False, cfg=cfg_default):
        super().__init__()
        self.eval_info = eval_info
        Hs = self.lidar_Hs = [int(np.ceil(cfg.init_H / s)) for s in np.cumprod(cfg.stride_Hs)]
        Ws = self.lidar_Ws = [int(np.ceil(cfg.init_W / s)) for s in np.cumprod(cfg.stride_Ws)]
        enc, bnkw = cfg.lidar_encoder_mlps, dict(use_trans=cfg.use_trans, use_bn_p=cfg.use_bn_p,
                                                 use_bn_input=cfg.use_bn_input)
        in_H, in_W = [cfg.init_H] + Hs[:3], [cfg.init_W] + Ws[:3]
        in_ch = [cfg.lidar_feature_size + (4 if cfg.using_intens else 3)] + [m[-1] + 3 for m in enc[:3]]
        for lv in range(4):
            setattr(self, "LiDAR_lv%d" % (lv + 1), ProjectPointNet(
                H=in_H[lv], W=in_W[lv], out_h=Hs[lv], out_w=Ws[lv], stride_H=cfg.stride_Hs[lv],
                stride_W=cfg.stride_Ws[lv], kernel_size=cfg.kernel_sizes[lv], nsample=cfg.lidar_group_samples[lv],
                distance=cfg.down_conv_dis[lv], in_channel=in_ch[lv], mlp=enc[lv], **bnkw))
        self.layer_idx = ProjectPointNet(
            H=Hs[2], W=Ws[2], out_h=Hs[3], out_w=Ws[3], stride_H=cfg.stride_Hs[3], stride_W=cfg.stride_Ws[3],
            kernel_size=cfg.kernel_sizes[3], nsample=cfg.lidar_group_samples[4], distance=cfg.down_conv_dis[3],
            in_channel=cfg.cost_volume_mlps[-1][-1] + 3, mlp=enc[4], **bnkw)

        for i, (cin, chans, strides) in enumerate(cfg.rgb_encoder_channels):
            setattr(self, "RGB_net%d" % (i + 1), createCNNs(cin, chans, strides))

        rgb_c, lf3_c = cfg.rgb_encoder_channels[-1][1][-1], enc[-3][-1]
        for i in range(2):
            setattr(self, "cost_volume%d" % (i + 1), CostVolume(
                H=Hs[2], W=Ws[2], kernel_size=cfg.cost_volume_kernel_size[i], distance=cfg.cost_volume_dis[i],
                nsample=cfg.cost_volume_nsamples[0], nsample_q=cfg.cost_volume_nsamples[1][i],
                rgb_in_channels=rgb_c, lidar_in_channels=lf3_c, mlp1=cfg.cost_volume_mlps[0],
                mlp2=cfg.cost_volume_mlps[1], backward_validation=cfg.backward_validation[i], **bnkw))

        fp = cfg.flow_predictor_mlps
        fpkw = dict(is_training=self.training, bn_decay=bn_decay, bn=cfg.use_bn_p, use_bn_input=cfg.use_bn_input)
        self.flow_predictor0 = FlowPredictor(in_channels=enc[-2][-1] + enc[-1][-1], mlp=fp[0], **fpkw)
        up = dict(H=Hs[-1], W=Ws[-1], out_h=Hs[-2], out_w=Ws[-2], stride_H=cfg.stride_Hs[-1],
                  stride_W=cfg.stride_Ws[-1], **bnkw)
        self.set_upconv0_w_upsample = ProjSetUpconvModule(
            kernel_size=cfg.up_conv_kernel_size[0], nsample=cfg.setupconv_nsamples[0], distance=cfg.up_conv_dis[0],
            in_channels=[lf3_c, fp[0][-1]], mlp=cfg.setupconv_mlps[0][0], mlp2=cfg.setupconv_mlps[0][1], **up)
        self.set_upconv0_upsample = ProjSetUpconvModule(
            kernel_size=cfg.up_conv_kernel_size[1], nsample=cfg.setupconv_nsamples[1], distance=cfg.up_conv_dis[1],
            in_channels=[lf3_c, enc[-1][-1]], mlp=cfg.setupconv_mlps[1][0], mlp2=cfg.setupconv_mlps[1][1], **up)
        self.flow_predictor0_predict = FlowPredictor(
            in_channels=lf3_c + cfg.setupconv_mlps[1][1][-1] + cfg.cost_volume_mlps[-1][-1], mlp=fp[1], **fpkw)
        self.flow_predictor0_w = FlowPredictor(
            in_channels=lf3_c + cfg.setupconv_mlps[0][-1][-1] + fp[1][-1], mlp=fp[2], **fpkw)

        head = dict(hidden=cfg.head_hidden_dim, q_dim=cfg.rotation_quat_head_dim, t_dim=cfg.transition_vec_head_dim,
                    dropout_rate=cfg.head_dropout_rate, split_dp=cfg.split_dp, pos_embed=cfg.head_pos_embedding,
                    sigmoid=cfg.mask_sigmoid, maxhead=cfg.max_head)
        self.l4_head = PoseHead(in_channels=[enc[-1][-1], enc[-2][-1]], mlp1=cfg.pose_head_mlps[0][0],
                                mlp2=cfg.pose_head_mlps[0][1], **head)
        self.l3_head = PoseHead(in_channels=[fp[1][-1], lf3_c], mlp1=cfg.pose_head_mlps[1][0],
                                mlp2=cfg.pose_head_mlps[1][1], **head)
        self.sq = nn.Parameter(torch.tensor([cfg.sq_init]), requires_grad=True)
        self.sx = nn.Parameter(torch.tensor([cfg.sx_init]), requires_grad=True)

    def forward(self, rgb_img, lidar_img, lidar_img_raw, H_initial, intrinsic, resize_img, gt_project=None,
                calib=None, lidar_feature=None, cfg=cfg_default):
        """rgb_img (B,3,h,w); lidar_img (B,N,3) points in the (mis-calibrated) camera frame;
        lidar_img_raw (B,N,3) the same points in the LiDAR frame (drives the range image);
        intrinsic (B,3,3); lidar_feature (B,N,D) or None.
        -> out_3 (B,7) refined [q,t], result_4 (B,7) coarse [q,t], None, None, sx, sq"""
        s = self._coarse(rgb_img, lidar_img, lidar_img_raw, intrinsic, lidar_feature, cfg)
        out_3, W_l3, _, _ = self._refine(s, s["q4"], s["t4"], cfg)
        return self._result(s, out_3, W_l3)

    def _result(self, s, out_3, W_l3):
        if self.eval_info:
            return (out_3.float(), s["result_4"].float(), self.sx, self.sq, W_l3, s["P3_l4"], None, None,
                    s["P4"].view(s["B"], -1, 3))
        return out_3.float(), s["result_4"].float(), None, None, self.sx, self.sq

    def _coarse(self, rgb_img, lidar_img, lidar_img_raw, intrinsic, lidar_feature, cfg):
        """Everything up to the coarse (level-4) pose and the two up-convolutions (:216-343 of the reference):
        returns the tensors the level-3 refinement needs."""
        dev = rgb_img.device
        intrinsic = intrinsic.float()
        B = rgb_img.shape[0]
        N = lidar_img.shape[1]
        rfkw = dict(cfg=cfg, raw_feat_point=cfg.raw_feat_point)

        # The image pyramid and the LiDAR pyramid are independent up to the first cost volume: the image branch goes
        # to a side stream (streams.py), joined right before its output is first read.
        with Fork(rgb_img) as rgb_branch:
            # pixel centres of RF3 on the normalised camera plane, K3^-1 [u, v, 1]: they depend on RF3's SHAPE only, so
            # they are issued ahead of the pyramid instead of on the critical path between the pyramids and cost volume 1
            h3, w3 = pyramid_out_hw((self.RGB_net1, self.RGB_net2, self.RGB_net3), rgb_img.shape[2], rgb_img.shape[3])
            RF3_index = pixel_rays(intrinsic, h3, w3, rgb_img.shape[2], rgb_img.shape[3])
            RF3 = self.RGB_net3(self.RGB_net2(self.RGB_net1(rgb_img)))

        lidar_norm = torch.zeros(B, N, 3, device=dev) if lidar_feature is None else lidar_feature
        lidar_img_raw, (lidar_norm, lidar_img) = project_seq(
            lidar_img_raw.float(), [lidar_norm.float(), lidar_img.float()], cfg.init_H, cfg.init_W, cfg.rank,
            cfg.fup, cfg.fdown)

        P1_raw, P1, LF1, _, _ = self.LiDAR_lv1.forward_center(lidar_img_raw, lidar_img, lidar_norm,
                                                              using_intens=cfg.using_intens, **rfkw)
        P2_raw, P2, LF2, _, _ = self.LiDAR_lv2(P1_raw, P1, LF1, **rfkw)
        P3_raw, P3, LF3, _, _ = self.LiDAR_lv3(P2_raw, P2, LF2, **rfkw)
        P4_raw, P4, LF4, _, sample_idx_4 = self.LiDAR_lv4(P3_raw, P3, LF3, **rfkw)

        RF3, RF3_index = rgb_branch.join(RF3, RF3_index)
        assert RF3.shape[2:] == (h3, w3), (RF3.shape, h3, w3)

        H3, W3 = self.lidar_Hs[2], self.lidar_Ws[2]
        H4, W4 = self.lidar_Hs[-1], self.lidar_Ws[-1]
        P3_l4 = P3.reshape(B, H3 * W3, 3)
        LF3_cv = LF3.reshape(B, H3 * W3, -1)
        lidar_z = P3_l4[:, :, 2:]
        lidar_uv = P3_l4 / (lidar_z + 1e-10)
        RF3 = RF3.reshape(B, RF3.shape[1], -1).permute(0, 2, 1)            # B,hw,C
        l3_grid = StrideGrid(B, H3, W3, 1, 1, dev)                          # every level-3 pixel is a centre

        # ---- level 4: coarse pose from the all-pixel cost volume
        concat_4 = self.cost_volume1(P3_raw, lidar_uv, LF3_cv, l3_grid, RF3_index, RF3, lidar_z, cfg=cfg)
        _, _, l4_points_predict, _, _ = self.layer_idx(P3_raw, P3, concat_4, sample_idx=sample_idx_4, **rfkw)
        # The two up-convolutions need only the level-4 embedding / mask, not the coarse pose: each is forked as soon as
        # its input exists and joined where the level-3 refinement first reads it (_refine), so they run beside the
        # mask predictor, the coarse pose head and the second cost volume.
        with Fork(P3_raw, P4_raw, P3, P4, LF3, l4_points_predict) as up_branch:
            l3_up = self.set_upconv0_upsample(P3_raw, P4_raw, P3, P4, l3_grid, LF3, l4_points_predict, **rfkw)
        l4_valid = check_valid(P4_raw).view(B, -1, 1)
        l4_w = self.flow_predictor0(LF4.view(B, H4 * W4, -1), None, l4_points_predict.view(B, H4 * W4, -1))
        l4_w = _mask_fill(l4_w, l4_valid)
        l4_w_img = l4_w.view(B, H4, W4, -1)
        with Fork(P3_raw, P4_raw, P3, P4, LF3, l4_w_img) as up_w_branch:
            l3_w_up = self.set_upconv0_w_upsample(P3_raw, P4_raw, P3, P4, l3_grid, LF3, l4_w_img, **rfkw)
        q4, t4, _ = self.l4_head(l4_points_predict.view(B, H4 * W4, -1), l4_w, P4.view(B, H4 * W4, 3),
                                 LF4.view(B, H4 * W4, -1), None)
        result_4 = torch.cat([q4, t4], dim=1)
        return dict(B=B, H3W3=H3 * W3, P3_raw=P3_raw, P3_l4=P3_l4, P4=P4, LF3_cv=LF3_cv, l3_grid=l3_grid,
                    RF3_index=RF3_index, RF3=RF3, l3_w_up=l3_w_up, l3_up=l3_up, q4=q4, t4=t4, result_4=result_4,
                    pending={"l3_up": up_branch, "l3_w_up": up_w_branch})

    def _refine(self, s, q_in, t_in, cfg):
        """Level 3 (:331-404): warp the level-3 points by the pose (q_in, t_in), correlate with the image again,
        regress the residual pose (q3, t3) and compose  q = q3 q_in,  t = R(q3) t_in + t3.
        -> out_3 (B,7), attention weights, q3, t3"""
        B, n3 = s["B"], s["H3W3"]
        P3_l4, LF3_cv = s["P3_l4"], s["LF3_cv"]
        P3_warped = warp_utils.rigid_warp(P3_l4, q_in, t_in, mask_invalid=True)      # warp_quat_xyz(...) * check_valid(P3_l4)
        lidar_z = P3_warped[:, :, 2:]
        lidar_uv = P3_warped / (lidar_z + 1e-10)
        concat_3 = self.cost_volume2(s["P3_raw"], lidar_uv, LF3_cv, s["l3_grid"], s["RF3_index"], s["RF3"], lidar_z, cfg=cfg)
        for key in ("l3_up", "l3_w_up"):                       # first refinement: the up-convolution branches end here
            branch = s["pending"].pop(key, None)
            if branch is not None:
                s[key] = branch.join(s[key])
        l3_predict = self.flow_predictor0_predict(LF3_cv, s["l3_up"].view(B, n3, -1), concat_3.view(B, n3, -1))
        l3_w = self.flow_predictor0_w(LF3_cv, s["l3_w_up"].view(B, n3, -1), l3_predict)
        l3_w = _mask_fill(l3_w, check_valid(s["P3_raw"]).view(B, -1, 1))
        q3, t3, W_l3 = self.l3_head(l3_predict, l3_w, P3_warped, LF3_cv, None)

        q = warp_utils.mul_q(q3.view(B, 1, 4), q_in.view(B, 1, 4)).squeeze(1)
        t = warp_utils.rigid_warp(t_in.view(B, 1, 3), q3, t3).view(B, 3)             # t = R(q3) t_in + t3  (:414-421)
        return torch.cat([q, t], 1), W_l3, q3, t3

    def set_bn(self):
        for name in ("flow_predictor0", "flow_predictor0_w", "flow_predictor0_predict", "LiDAR_lv1", "LiDAR_lv2",
                     "LiDAR_lv3", "LiDAR_lv4", "layer_idx", "set_upconv0_upsample", "set_upconv0_w_upsample",
                     "cost_volume1", "cost_volume2"):
            getattr(self, name).set_bn()


def get_num_parameters(model, trainable=False):
    return sum(p.numel() for p in model.parameters() if p.requires_grad or not trainable)
