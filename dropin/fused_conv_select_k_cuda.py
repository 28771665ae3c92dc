"""Drop-in for the reference's pybind module `fused_conv_select_k_cuda`
(src/projectPN/fused_conv_select/fused_conv_g.cpp:15-73), imported top-level by
src/projectPN/fused_conv_select/fused_conv_select_k.py:5.  With `<repo>/dropin` on sys.path
that file -- and everything above it -- runs unchanged on libi2p_b200.so.
"""
from i2pnet_b200 import _cabi


def fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K, flag,
                        distance, stride_h, stride_w, selected_b_idx, selected_h_idx, selected_w_idx,
                        valid_idx, valid_in_dis_idx, selected_mask, small_h, small_w):
    # valid_idx / valid_in_dis_idx: accepted and left untouched, exactly like the reference
    # kernel, whose writes to them are commented out (fused_conv_go.cu:150,167).
    _cabi.fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kernel_size_H, kernel_size_W, K,
                              flag, distance, stride_h, stride_w, selected_b_idx, selected_h_idx,
                              selected_w_idx, selected_mask, small_h, small_w)
