"""ctypes binding of libi2p_b200.so (include/i2p_b200.h) for torch tensors.

This is the only place where tensors become raw device pointers.  Every wrapper checks
device / dtype / contiguity (the reference only asserts contiguity, pointnet2_utils.py:24-25,
and CHECK_INPUTs two of its ten entry points), launches on torch's CURRENT stream under the
tensor's device guard, and raises I2PError on a non-zero return code -- the reference calls
exit(-1) instead (pointnet2/src/group_points_gpu.cu:81-85).

There is deliberately no fallback: if the library is missing or a tensor is not on a CUDA
device the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libi2p_b200.so")


class I2PError(RuntimeError):
    pass


_lib = None

_vp = ctypes.c_void_p
_int = ctypes.c_int
_flt = ctypes.c_float
_ll = ctypes.c_longlong

_SIGNATURES = {
    "i2p_furthest_point_sampling": [_int, _int, _int, _vp, _vp, _vp, _vp],
    "i2p_gather_points": [_int] * 4 + [_vp] * 4,
    "i2p_gather_points_grad": [_int] * 4 + [_vp] * 4,
    "i2p_ball_query": [_int, _int, _int, _flt, _int, _vp, _vp, _vp, _vp],
    "i2p_group_points": [_int] * 5 + [_vp] * 4,
    "i2p_group_points_grad": [_int] * 5 + [_vp] * 4,
    "i2p_three_nn": [_int] * 3 + [_vp] * 5,
    "i2p_three_interpolate": [_int] * 4 + [_vp] * 5,
    "i2p_three_interpolate_grad": [_int] * 4 + [_vp] * 5,
    "i2p_knn": [_int] * 4 + [_vp] * 5,
    "i2p_fused_conv_select_k": [_int] * 8 + [_flt, _int, _int] + [_vp] * 8 + [_int, _int, _vp],
    "i2p_select_k_flat": [_int] * 8 + [_flt, _int, _int] + [_vp] * 3 + [_int] * 3 + [_vp] * 2 + [_int, _int, _vp],
    "i2p_gather_rows": [_int] * 4 + [_vp] * 4,
    "i2p_sa_geometry": [_int] * 4 + [_vp] * 6,
    "i2p_gather_rows_grad": [_int] * 4 + [_vp] * 4,
    "i2p_knn_point": [_int] * 4 + [_vp] * 5,
    "i2p_project_seq": [_int] * 4 + [_flt, _flt, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "i2p_pw_linear_fwd": [_int] * 3 + [_vp] * 3 + [_flt] + [_vp] * 5,
    "i2p_bn_finalize": [_int, _int, _vp, _vp, _vp, _flt, _vp, _vp, _vp, _vp, _vp],
    "i2p_bn_act": [_ll, _int, _vp, _vp, _vp, _flt, _vp, _vp],
    "i2p_bn_act_maxk": [_ll, _int, _int, _vp, _vp, _vp, _flt, _vp, _vp, _vp],
    "i2p_bn_bwd_reduce": [_ll, _int, _vp, _vp, _vp, _int] + [_vp] * 5 + [_flt, _vp, _vp],
    "i2p_pw_linear_bwd_dx": [_int] * 3 + [_vp] * 3 + [_int] + [_vp] * 5 + [_flt] + [_vp] * 3 + [_vp] * 5 + [_flt, _vp, _vp],
    "i2p_pw_pack_weights": [_int, _int, _vp, _vp, _vp],
    "i2p_pw_pack_weights_multi": [_int, _vp, _vp],
    "i2p_pw_linear_fwd_tc": [_int] * 3 + [_vp] * 3 + [_flt] + [_vp] * 5,
    "i2p_pw_linear_bwd_dx_tc": [_int] * 3 + [_vp] * 3 + [_int] + [_vp] * 5 + [_flt] + [_vp] * 3 + [_vp] * 5 + [_flt, _vp, _vp],
    "i2p_pw_linear_bwd_dw_tc": [_int] * 3 + [_vp] * 3 + [_int] + [_vp] * 5 + [_flt] + [_vp] * 4 + [_flt, _vp, _vp],
    "i2p_cv_build": [_int] * 5 + [_vp] * 9,
    "i2p_cv_build_bwd": [_int] * 6 + [_vp] * 11,
    "i2p_softmax_wsum": [_ll, _int, _int] + [_vp] * 6,
    "i2p_softmax_wsum_bwd": [_ll, _int, _int] + [_vp] * 9,
    "i2p_cv_prep_fwd": [_int] * 5 + [_vp] * 16,
    "i2p_cv_prep_scratch_floats": [_int] * 4,
    "i2p_cv_prep_bwd": [_int] * 5 + [_vp] * 19,
    "i2p_pixel_rays": [_int, _int, _int, _flt, _flt, _vp, _vp, _vp],
    "i2p_quat_mul": [_int] * 6 + [_vp] * 4,
    "i2p_quat_warp_fwd": [_int] * 3 + [_vp] * 5,
    "i2p_quat_warp_bwd": [_int] * 3 + [_vp] * 7,
    "i2p_clip_adam_step": [_ll] + [_vp] * 5 + [_flt] * 6 + [_int, _vp],
    "i2p_rgb_bn_stats": [_int] * 4 + [_vp, _vp, _vp],
    "i2p_rgb_bn_finalize": [_int, _int, _vp, _vp, _vp, _flt, _flt, _vp, _vp, _vp, _vp, _vp, _vp],
    "i2p_rgb_bn_from_running": [_int, _vp, _vp, _flt, _vp, _vp, _vp, _vp, _vp],
    "i2p_rgb_bn_act_pool_fwd": [_int] * 5 + [_vp, _vp, _flt, _vp, _vp],
    "i2p_rgb_bn_act_pool_bwd": [_int] * 6 + [_vp, _vp, _flt] + [_vp] * 6,
    "i2p_pose_head_fwd": [_int] * 4 + [_vp] * 16,
    "i2p_pose_head_bwd": [_int] * 4 + [_vp] * 20,
    "i2p_pose_loss_fwd": [_int, _int] + [_vp] * 8,
    "i2p_pose_loss_bwd": [_int, _int] + [_vp] * 12,
    "i2p_conv3x3_pack": [_int, _int, _int, _vp, _vp, _vp],
    "i2p_conv3x3_pack_multi": [_int, _vp, _vp],
    "i2p_conv3x3_tc": [_int] * 5 + [_vp] * 6,
    "i2p_conv3x3_wgrad": [_int] * 5 + [_vp] * 4,
    "i2p_pw_linear_bwd_dw": [_int] * 3 + [_vp] * 3 + [_int] + [_vp] * 5 + [_flt] + [_vp] * 4 + [_flt, _vp, _vp],
}


def exported_symbols():
    """Every entry point include/i2p_b200.h declares."""
    return sorted(list(_SIGNATURES) + ["i2p_last_error", "i2p_abi_version", "i2p_launch_count", "i2p_pw_num_tiles",
                   "i2p_set_mlp_tensor_cores", "i2p_get_mlp_tensor_cores", "i2p_rgb_num_chunks", "i2p_rgb_pool_out", "i2p_rgb_s12_slots", "i2p_pw_tc_supported", "i2p_pw_pack_floats", "i2p_optim_state_bytes", "i2p_optim_lr_offset", "i2p_conv3x3_pack_floats", "i2p_conv3x3_tiles", "i2p_conv3x3_stat_slots"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise I2PError(
                "%s is missing: build it with `python -m i2pnet_b200._build` (nvcc, sm_100a). "
                "i2pnet_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _int
        L.i2p_last_error.restype = ctypes.c_char_p
        L.i2p_abi_version.restype = _int
        L.i2p_launch_count.restype = ctypes.c_uint64
        L.i2p_pw_num_tiles.argtypes = [_int]
        L.i2p_pw_num_tiles.restype = _int
        L.i2p_pw_tc_supported.argtypes = [_int] * 4
        L.i2p_pw_tc_supported.restype = _int
        L.i2p_pw_pack_floats.argtypes = [_int, _int]
        L.i2p_pw_pack_floats.restype = _ll
        L.i2p_optim_state_bytes.restype = _int
        L.i2p_optim_lr_offset.restype = _int
        L.i2p_conv3x3_pack_floats.argtypes = [_int] * 3
        L.i2p_conv3x3_pack_floats.restype = _ll
        L.i2p_conv3x3_tiles.argtypes = [_int, _int]
        L.i2p_conv3x3_tiles.restype = _int
        L.i2p_conv3x3_stat_slots.argtypes = [_int] * 4
        L.i2p_conv3x3_stat_slots.restype = _int
        L.i2p_rgb_num_chunks.argtypes = [_int]
        L.i2p_rgb_num_chunks.restype = _int
        L.i2p_rgb_s12_slots.restype = _int
        L.i2p_rgb_pool_out.argtypes = [_int, _int]
        L.i2p_rgb_pool_out.restype = _int
        L.i2p_set_mlp_tensor_cores.argtypes = [_int]
        L.i2p_set_mlp_tensor_cores.restype = None
        L.i2p_get_mlp_tensor_cores.restype = _int
        _lib = L
    return _lib


def launch_count():
    return int(lib().i2p_launch_count())


def _check(rc, name):
    if rc != 0:
        raise I2PError("%s failed (code %d): %s" % (name, rc, lib().i2p_last_error().decode()))


def _ptr(t, dtype, name, device=None):
    if not isinstance(t, torch.Tensor):
        raise I2PError("%s: expected a tensor, got %r" % (name, type(t)))
    if not t.is_cuda:
        raise I2PError("%s: tensor is on %s; i2pnet_b200 operators run on CUDA only" % (name, t.device))
    if t.dtype != dtype:
        raise I2PError("%s: expected %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise I2PError("%s: tensor must be contiguous" % name)
    if device is not None and t.device != device:
        raise I2PError("%s: tensor is on %s, expected %s" % (name, t.device, device))
    return t.data_ptr()


def call(name, device, *args):
    """Invoke `name` on torch's current stream of `device` (appended as the last argument)."""
    fn = getattr(lib(), name)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = fn(*args, stream)
    _check(rc, name)


f32, i32, i64 = torch.float32, torch.int32, torch.int64


# ---- thin typed wrappers (shapes documented in include/i2p_b200.h) ---------------------------

def furthest_point_sampling(b, n, m, points, temp, idx):
    d = points.device
    call("i2p_furthest_point_sampling", d, b, n, m, _ptr(points, f32, "points"), _ptr(temp, f32, "temp", d),
         _ptr(idx, i32, "idx", d))


def gather_points(b, c, n, npoints, points, idx, out):
    d = points.device
    call("i2p_gather_points", d, b, c, n, npoints, _ptr(points, f32, "points"), _ptr(idx, i32, "idx", d),
         _ptr(out, f32, "out", d))


def gather_points_grad(b, c, n, npoints, grad_out, idx, grad_points):
    d = grad_out.device
    call("i2p_gather_points_grad", d, b, c, n, npoints, _ptr(grad_out, f32, "grad_out"), _ptr(idx, i32, "idx", d),
         _ptr(grad_points, f32, "grad_points", d))


def ball_query(b, n, m, radius, nsample, new_xyz, xyz, idx):
    d = xyz.device
    call("i2p_ball_query", d, b, n, m, float(radius), nsample, _ptr(new_xyz, f32, "new_xyz", d),
         _ptr(xyz, f32, "xyz"), _ptr(idx, i32, "idx", d))


def group_points(b, c, n, npoints, nsample, points, idx, out):
    d = points.device
    call("i2p_group_points", d, b, c, n, npoints, nsample, _ptr(points, f32, "points"), _ptr(idx, i32, "idx", d),
         _ptr(out, f32, "out", d))


def group_points_grad(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    d = grad_out.device
    call("i2p_group_points_grad", d, b, c, n, npoints, nsample, _ptr(grad_out, f32, "grad_out"),
         _ptr(idx, i32, "idx", d), _ptr(grad_points, f32, "grad_points", d))


def three_nn(b, n, m, unknown, known, dist2, idx):
    d = unknown.device
    call("i2p_three_nn", d, b, n, m, _ptr(unknown, f32, "unknown"), _ptr(known, f32, "known", d),
         _ptr(dist2, f32, "dist2", d), _ptr(idx, i32, "idx", d))


def three_interpolate(b, c, m, n, points, idx, weight, out):
    d = points.device
    call("i2p_three_interpolate", d, b, c, m, n, _ptr(points, f32, "points"), _ptr(idx, i32, "idx", d),
         _ptr(weight, f32, "weight", d), _ptr(out, f32, "out", d))


def three_interpolate_grad(b, c, n, m, grad_out, idx, weight, grad_points):
    d = grad_out.device
    call("i2p_three_interpolate_grad", d, b, c, n, m, _ptr(grad_out, f32, "grad_out"), _ptr(idx, i32, "idx", d),
         _ptr(weight, f32, "weight", d), _ptr(grad_points, f32, "grad_points", d))


def knn(b, n, m, k, unknown, known, dist2, idx):
    d = unknown.device
    call("i2p_knn", d, b, n, m, k, _ptr(unknown, f32, "unknown"), _ptr(known, f32, "known", d),
         _ptr(dist2, f32, "dist2", d), _ptr(idx, i32, "idx", d))


def fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kH, kW, K, flag, distance, stride_h,
                        stride_w, sel_b, sel_h, sel_w, sel_mask, small_h, small_w):
    d = xyz1.device
    call("i2p_fused_conv_select_k", d, xyz1.shape[0], H, W, npoints, kH, kW, K, int(flag), float(distance),
         stride_h, stride_w, _ptr(xyz1, f32, "xyz1"), _ptr(xyz2, f32, "xyz2", d), _ptr(idx_n2, i32, "idx_n2", d),
         _ptr(random_hw, i32, "random_hw", d), _ptr(sel_b, i64, "selected_b_idx", d),
         _ptr(sel_h, i64, "selected_h_idx", d), _ptr(sel_w, i64, "selected_w_idx", d),
         _ptr(sel_mask, f32, "selected_mask", d), small_h, small_w)


def select_k_flat(xyz1, xyz2, idx_n2, grid, kernel, K, flag, distance, stride, flat_idx, mask):
    """grid = (out_h, out_w, stride_ch, stride_cw) when idx_n2 is None."""
    d = xyz1.device
    B, H, W, _ = xyz1.shape
    sh, sw = xyz2.shape[1:3]
    if idx_n2 is None:
        out_h, out_w, sch, scw = grid
        npoints = out_h * out_w
        pidx = None
    else:
        npoints = idx_n2.shape[1]
        out_w, sch, scw = 1, 1, 1
        pidx = _ptr(idx_n2, i32, "idx_n2", d)
    call("i2p_select_k_flat", d, B, H, W, npoints, kernel[0], kernel[1], K, int(flag), float(distance),
         stride[0], stride[1], _ptr(xyz1, f32, "xyz1"), _ptr(xyz2, f32, "xyz2", d), pidx, out_w, sch, scw,
         _ptr(flat_idx, i32, "flat_idx", d), _ptr(mask, f32, "mask", d), sh, sw)


def gather_rows(b, hw, c, m, feature, flat_idx, out):
    d = feature.device
    call("i2p_gather_rows", d, b, hw, c, m, _ptr(feature, f32, "feature"), _ptr(flat_idx, i32, "flat_idx", d),
         _ptr(out, f32, "out", d))


def gather_rows_grad(b, hw, c, m, grad_out, flat_idx, grad_feature):
    d = grad_out.device
    call("i2p_gather_rows_grad", d, b, hw, c, m, _ptr(grad_out, f32, "grad_out"),
         _ptr(flat_idx, i32, "flat_idx", d), _ptr(grad_feature, f32, "grad_feature", d))


def knn_point(b, n, s, nsample, xyz, new_xyz, group_idx, dist_out=None):
    d = xyz.device
    call("i2p_knn_point", d, b, n, s, nsample, _ptr(xyz, f32, "xyz"), _ptr(new_xyz, f32, "new_xyz", d),
         _ptr(group_idx, i64, "group_idx", d), None if dist_out is None else _ptr(dist_out, f32, "dist_out", d))


def project_seq(xyz, feats, H, W, fup, fdown, xyz_proj, feat_projs, owner):
    d = xyz.device
    B, N, _ = xyz.shape
    nf = len(feats)
    fp = (ctypes.c_void_p * max(nf, 1))(*[_ptr(f, f32, "feature", d) for f in feats])
    op = (ctypes.c_void_p * max(nf, 1))(*[_ptr(o, f32, "feature_proj", d) for o in feat_projs])
    dims = (ctypes.c_int * max(nf, 1))(*[int(f.shape[-1]) for f in feats])
    call("i2p_project_seq", d, B, N, H, W, float(fup), float(fdown), _ptr(xyz, f32, "xyz"), nf,
         ctypes.cast(fp, ctypes.c_void_p), ctypes.cast(dims, ctypes.c_void_p), _ptr(xyz_proj, f32, "xyz_proj", d),
         ctypes.cast(op, ctypes.c_void_p), _ptr(owner, i32, "owner", d))
