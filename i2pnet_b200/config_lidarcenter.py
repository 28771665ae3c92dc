"""Hyper-parameters of the small-range model (mirror of src/config_lidarcenter.py): plain class attributes passed
as `cfg=` through the forward, like the reference does."""
from .modules.MainModules import CostVolume, PoseHead


class I2PNetConfig:
    rgb_encoder_channels = [
        # in_channel, channels of the 3x3 convolutions, strides of the max-pools
        (3, [16, 16, 16, 16, 32], [2, 1, 1, 1, 2]),
        (32, [32, 32, 32, 32, 64], [2, 1, 1, 1, 2]),
        (64, [64, 64, 64, 64, 128], [1, 1, 1, 1, 2]),
    ]
    lidar_downsample_rate = [4, 2, 4, 4]
    lidar_in_points = 8192
    lidar_feature_size = 7
    featmode = 'dim10feat'
    raw_feat_point = True
    lidar_group_samples = [32, 16, 16, 16, 16]
    lidar_encoder_mlps_planA = [[8, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 128], [128, 64, 64]]
    lidar_encoder_mlps_planB = [[16, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 256], [128, 64, 64]]
    lidar_encoder_mlps = lidar_encoder_mlps_planB
    # cost volume
    backward_fc = False
    cost_volume_mlps = [[128, 64, 64],      # mlp1: per (point, pixel) features
                        [128, 64]]          # mlp2: attention weights (pixel stage and point stage)
    cost_volume_nsamples = [4,              # 3-D neighbours of the second stage
                            [-1, 32]]       # pixels per point: all of level 3 / the 32 nearest
    cost_volume_corr_func = CostVolume.CorrFunc.ELEMENTWISE_PRODUCT
    backward_validation = [True, False]
    max_cost = False
    setupconv_mlps = [[[128, 64], [64]], [[128, 64], [64]]]     # mask / embedding up-sampling
    setupconv_nsamples = [8, 8]
    flow_predictor_mlps = [[128, 64], [128, 64], [128, 64]]     # l4 mask, l3 refined embedding, l3 mask
    pose_head_mlps = [[[], []], [[], []]]
    head_hidden_dim = 256
    rotation_quat_head_dim = 4
    transition_vec_head_dim = 3
    head_dropout_rate = 0.5
    head_corr_func = PoseHead.CorrFunc.CONCAT
    head_pos_embedding = False
    split_dp = False
    max_head = False
    # projection mask (all off in the shipped configuration)
    use_projection_mask = False
    layer_mask = [False, True]
    projection_mask_mlps = [[128, 64], [128, 64]]
    mask_sigmoid = False
    ground_truth_projection_mask = False
    ground_truth_projection_mask_eval = False
    ground_truth_mask_layer = [False, True]
    ab_delay = False
    mask_delay = False
    mask_delay_step = 1904 * 8 * 30
    one_head_mask = False
    # loss
    sq_init = -2.5
    sx_init = 0.0
    l1_trans_loss = True
    pointwise_reproject_loss = False
    focal_mask_loss = True
    focal_gamma = 2
