"""ctypes binding of the CPU oracle (oracle/i2p_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(i2pnet_b200/) never imports this module.

All functions take and return numpy arrays (float32 / int32 / int64) with the
reference's layouts; see i2p_oracle.c for the reference file:line each follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libi2p_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)
_l = ctypes.POINTER(ctypes.c_int64)
_int = ctypes.c_int
_flt = ctypes.c_float


def build(force=False):
    src = os.path.join(_HERE, "i2p_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libi2p_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_fps_block_size.argtypes = [_int]
        L.orc_fps_block_size.restype = _int
        L.orc_fps.argtypes = [_int, _int, _int, _f, _f, _i]
        L.orc_gather_points.argtypes = [_int] * 4 + [_f, _i, _f]
        L.orc_gather_points_grad.argtypes = [_int] * 4 + [_f, _i, _f]
        L.orc_ball_query.argtypes = [_int, _int, _int, _flt, _int, _f, _f, _i]
        L.orc_group_points.argtypes = [_int] * 5 + [_f, _i, _f]
        L.orc_group_points_grad.argtypes = [_int] * 5 + [_f, _i, _f]
        L.orc_three_nn.argtypes = [_int] * 3 + [_f, _f, _f, _i]
        L.orc_three_interpolate.argtypes = [_int] * 4 + [_f, _i, _f, _f]
        L.orc_three_interpolate_grad.argtypes = [_int] * 4 + [_f, _i, _f, _f]
        L.orc_fused_conv_select_k.argtypes = [_int] * 8 + [_flt, _int, _int, _f, _f, _i, _i, _l, _l, _l, _f, _int, _int]
        L.orc_fused_conv_select_k.restype = _int
        L.orc_knn.argtypes = [_int] * 4 + [_f, _f, _l, _f]
        L.orc_gather_rows.argtypes = [_int] * 4 + [_f, _i, _f]
        L.orc_gather_rows_grad.argtypes = [_int] * 4 + [_f, _i, _f]
        _lib = L
    return _lib


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_f)


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i)


def _lp(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_l)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fps_block_size(n):
    return lib().orc_fps_block_size(int(n))


def furthest_point_sample(xyz, npoint):
    """xyz (B,N,3) -> idx (B,npoint) int32; start index 0; reference tie order."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32)
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib().orc_fps(B, N, npoint, _fp(xyz), _fp(temp), _ip(idx))
    return idx


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    lib().orc_gather_points(B, C, N, M, _fp(points), _ip(idx), _fp(out))
    return out


def gather_points_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M = grad_out.shape
    g = np.zeros((B, C, N), dtype=np.float32)
    lib().orc_gather_points_grad(B, C, N, M, _fp(grad_out), _ip(idx), _fp(g))
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib().orc_ball_query(B, N, M, float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(idx))
    return idx


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, P, S = idx.shape
    out = np.empty((B, C, P, S), dtype=np.float32)
    lib().orc_group_points(B, C, N, P, S, _fp(points), _ip(idx), _fp(out))
    return out


def group_points_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, P, S = grad_out.shape
    g = np.zeros((B, C, N), dtype=np.float32)
    lib().orc_group_points_grad(B, C, N, P, S, _fp(grad_out), _ip(idx), _fp(g))
    return g


def three_nn(unknown, known):
    """-> (dist2 (B,N,3) squared, idx (B,N,3))"""
    unknown, known = _f32(unknown), _f32(known)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = np.empty((B, N, 3), dtype=np.float32)
    idx = np.empty((B, N, 3), dtype=np.int32)
    lib().orc_three_nn(B, N, M, _fp(unknown), _fp(known), _fp(d2), _ip(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, M = points.shape
    N = idx.shape[1]
    out = np.empty((B, C, N), dtype=np.float32)
    lib().orc_three_interpolate(B, C, M, N, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, M):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, N = grad_out.shape
    g = np.zeros((B, C, M), dtype=np.float32)
    lib().orc_three_interpolate_grad(B, C, N, M, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def fused_conv_select_k(xyz1, xyz2, idx_n2, random_hw, kH, kW, K, flag, distance, stride_h=1, stride_w=1):
    """-> (sel_b, sel_h, sel_w int64 (B,n,K), mask f32 (B,n,K)); outputs start at zero."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    idx_n2, random_hw = _i32(idx_n2), _i32(random_hw)
    B, H, W, _ = xyz1.shape
    sh, sw = xyz2.shape[1:3]
    n = idx_n2.shape[1]
    sb = np.zeros((B, n, K), dtype=np.int64)
    shh = np.zeros((B, n, K), dtype=np.int64)
    sww = np.zeros((B, n, K), dtype=np.int64)
    mask = np.zeros((B, n, K), dtype=np.float32)
    rc = lib().orc_fused_conv_select_k(B, H, W, n, kH, kW, K, int(flag), float(distance), stride_h, stride_w,
                                       _fp(xyz1), _fp(xyz2), _ip(idx_n2), _ip(random_hw),
                                       _lp(sb), _lp(shh), _lp(sww), _fp(mask), sh, sw)
    if rc != 0:
        raise ValueError("kernel window or K exceeds the reference's 150-entry arrays")
    return sb, shh, sww, mask


def knn(k, xyz, new_xyz, return_dist=False):
    """k nearest of xyz (B,N,3) for each new_xyz (B,S,3); sorted by (dist, index)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = np.zeros((B, S, k), dtype=np.int64)
    dist = np.zeros((B, S, k), dtype=np.float32)
    lib().orc_knn(B, N, S, k, _fp(xyz), _fp(new_xyz), _lp(idx), _fp(dist))
    return (idx, dist) if return_dist else idx


def gather_rows(feat, idx):
    """feat (B,HW,C), idx (B,M) flat -> (B,M,C)"""
    feat, idx = _f32(feat), _i32(idx)
    B, HW, C = feat.shape
    M = idx.shape[1]
    out = np.empty((B, M, C), dtype=np.float32)
    lib().orc_gather_rows(B, HW, C, M, _fp(feat), _ip(idx), _fp(out))
    return out


def gather_rows_grad(grad_out, idx, HW):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, M, C = grad_out.shape
    g = np.zeros((B, HW, C), dtype=np.float32)
    lib().orc_gather_rows_grad(B, HW, C, M, _fp(grad_out), _ip(idx), _fp(g))
    return g


def project_seq(xyz, feats, H, W, fup=2.0, fdown=-24.8):
    """xyz (B,N,3), feats list of (B,N,D) -> xyz_proj (B,H,W,3), [feat_proj (B,H,W,D)]; rank=False."""
    xyz = _f32(xyz)
    feats = [_f32(f) for f in feats]
    B, N, _ = xyz.shape
    nf = len(feats)
    xyz_proj = np.empty((B, H, W, 3), dtype=np.float32)
    outs = [np.empty((B, H, W, f.shape[-1]), dtype=np.float32) for f in feats]
    L = lib()
    fp = (_f * max(nf, 1))(*[_fp(f) for f in feats])
    op = (_f * max(nf, 1))(*[_fp(o) for o in outs])
    dims = (ctypes.c_int * max(nf, 1))(*[f.shape[-1] for f in feats])
    L.orc_project_seq.argtypes = [_int] * 4 + [_flt, _flt, _f, _int, ctypes.c_void_p, ctypes.c_void_p, _f,
                                              ctypes.c_void_p]
    L.orc_project_seq(B, N, H, W, float(fup), float(fdown), _fp(xyz), nf, ctypes.cast(fp, ctypes.c_void_p),
                      ctypes.cast(dims, ctypes.c_void_p), _fp(xyz_proj), ctypes.cast(op, ctypes.c_void_p))
    return xyz_proj, outs
