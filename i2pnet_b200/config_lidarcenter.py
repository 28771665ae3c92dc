"""Hyper-parameters of the small-range model (the values of the reference's src/config_lidarcenter.py): plain class
attributes passed as `cfg=` through the forward, like the reference does.  Grouped by the module that reads them."""
from .modules.MainModules import CostVolume, PoseHead

_RGB = [(3, [16, 16, 16, 16, 32], [2, 1, 1, 1, 2]),        # (in channels, 3x3 conv channels, max-pool strides) per pyramid net
        (32, [32, 32, 32, 32, 64], [2, 1, 1, 1, 2]),
        (64, [64, 64, 64, 64, 128], [1, 1, 1, 1, 2])]
_ENCODER_A = [[8, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 128], [128, 64, 64]]
_ENCODER_B = [[16, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 256], [128, 64, 64]]   # last entry: resampling set conv


class I2PNetConfig:
    # image pyramid
    rgb_encoder_channels = _RGB
    # point pyramid: 8192 -> /4 -> /2 -> /4 -> /4 points, k nearest neighbours per level (last: the cost-volume resampling)
    lidar_in_points, lidar_downsample_rate = 8192, [4, 2, 4, 4]
    lidar_group_samples = [32, 16, 16, 16, 16]
    lidar_feature_size, featmode, raw_feat_point = 7, 'dim10feat', True
    lidar_encoder_mlps_planA, lidar_encoder_mlps_planB = _ENCODER_A, _ENCODER_B
    lidar_encoder_mlps = _ENCODER_B
    # cost volumes: mlp1 on (point, pixel) pairs, mlp2 for the attention weights of both stages; 4 3-D neighbours;
    # pixels per point: all of level 3 in the first volume, the 32 nearest in the second
    cost_volume_mlps = [[128, 64, 64], [128, 64]]
    cost_volume_nsamples = [4, [-1, 32]]
    cost_volume_corr_func = CostVolume.CorrFunc.ELEMENTWISE_PRODUCT
    backward_validation, backward_fc, max_cost = [True, False], False, False
    # up-convolutions (mask, embedding) and the three per-point predictors (l4 mask, l3 embedding, l3 mask)
    setupconv_mlps = [[[128, 64], [64]], [[128, 64], [64]]]
    setupconv_nsamples = [8, 8]
    flow_predictor_mlps = [[128, 64], [128, 64], [128, 64]]
    # pose heads
    pose_head_mlps = [[[], []], [[], []]]
    head_hidden_dim, rotation_quat_head_dim, transition_vec_head_dim = 256, 4, 3
    head_dropout_rate, split_dp, max_head = 0.5, False, False
    head_corr_func, head_pos_embedding = PoseHead.CorrFunc.CONCAT, False
    # projection mask branch: off in the shipped configuration
    use_projection_mask, mask_sigmoid, one_head_mask = False, False, False
    layer_mask, ground_truth_mask_layer = [False, True], [False, True]
    projection_mask_mlps = [[128, 64], [128, 64]]
    ground_truth_projection_mask = ground_truth_projection_mask_eval = False
    ab_delay, mask_delay, mask_delay_step = False, False, 1904 * 8 * 30
    # loss
    sq_init, sx_init = -2.5, 0.0
    l1_trans_loss, pointwise_reproject_loss, focal_mask_loss, focal_gamma = True, False, True, 2
