"""Hyper-parameters of the large-range model (mirror of src/config_proj_lidarcenter.py and
src/config_proj_lidarcenter_nus.py): plain class attributes, passed as `cfg=` through every
forward like the reference does.  Only the attributes the forward/backward path reads are
kept; the reference's debug-dump machinery (debug_dict, Timings) is out of scope."""


def make_config(dataset_type=0, name="I2PNetConfig"):
    """dataset_type 0: KITTI (64x1800 range image), 1: nuScenes (32x1800 by the formula;
    the shipped nuScenes config overrides init_H, see I2PNetConfigNus)."""

    class _Cfg:
        use_bn_p = True
        use_bn_input = True
        use_trans = True
        rgb_encoder_channels = [
            (3, [16, 16, 16, 16, 32], [2, 1, 1, 1, 2]),
            (32, [32, 32, 32, 32, 64], [2, 1, 1, 1, 2]),
            (64, [64, 64, 64, 64, 128], [1, 1, 1, 1, 2]),
        ]
        stride_Hs = [2 ** (2 - dataset_type), 2, 2, 1]
        stride_Ws = [8, 2, 2, 2]
        rank = False
        debug = False
        debug_time = False
        down_conv_dis = [0.75, 3.0, 6.0, 12.0]
        init_H = 16 * 2 ** (2 - dataset_type)
        init_W = 1800
        fup, fdown = {0: (2.0, -24.8), 1: (10.0, -30.0), 2: (15.0, -15.0)}[dataset_type]
        kernel_sizes = [[9, 15], [9, 15], [5, 9], [5, 9]]
        lidar_feature_size = 7
        using_intens = False
        raw_feat_point = True
        lidar_group_samples = [32, 16, 16, 16, 16]
        lidar_encoder_mlps = [[16, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 256], [128, 64, 64]]
        cost_volume_dis = [4.5, 4.5]
        cost_volume_kernel_size = [[3, 5], [3, 5]]
        cost_volume_mlps = [[128, 64, 64], [128, 64]]
        cost_volume_nsamples = [4, [-1, 32]]
        backward_validation = [True, False]
        max_cost = False
        up_conv_dis = [9.0, 9.0]
        up_conv_kernel_size = [[5, 9], [5, 9]]
        setupconv_mlps = [[[128, 64], [64]], [[128, 64], [64]]]
        setupconv_nsamples = [8, 8]
        flow_predictor_mlps = [[128, 64], [128, 64], [128, 64]]
        pose_head_mlps = [[[], []], [[], []]]
        head_hidden_dim = 256
        rotation_quat_head_dim = 4
        transition_vec_head_dim = 3
        head_dropout_rate = 0.5
        head_pos_embedding = False
        split_dp = False
        max_head = False
        mask_sigmoid = False
        sq_init = -2.5
        sx_init = 0.0
        l1_trans_loss = True
        efgh = False  # read by train20v2learn_wandb_proj.py:453, defined by no shipped config

    _Cfg.__name__ = _Cfg.__qualname__ = name
    return _Cfg


I2PNetConfig = make_config(0)


def _nus():
    cfg = make_config(1, "I2PNetConfigNus")
    # src/config_proj_lidarcenter_nus.py:53,62-63: 21 rows, KITTI's vertical field of view
    cfg.init_H = 21
    cfg.fup, cfg.fdown = 2.0, -24.8
    return cfg


I2PNetConfigNus = _nus()
