"""i2pnet_b200 -- the I2PNet per-sample forward/backward hot path as hand-written sm_100a CUDA
behind the reference's operator / module surface.

    i2pnet_b200.pointnet2.pointnet2_utils     PointNet++ ops (FPS, ball query, group, gather, 3-NN, ...)
    i2pnet_b200.projectPN.utils               window select, row gather, projection, kNN grouping
    i2pnet_b200.projectPN.PPBackbone_center   ProjectPointNet, ProjSetUpconvModule, CostVolume, heads
    i2pnet_b200.modellearn_proj_center        RegNet_v2
    dropin/                                   `pointnet2.pointnet2_cuda`, `fused_conv_select_k_cuda`
                                              for the unchanged reference Python

Every operator is a launch into i2pnet_b200/lib/libi2p_b200.so (include/i2p_b200.h); there is
no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
