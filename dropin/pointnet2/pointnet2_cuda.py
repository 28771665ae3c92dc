"""Drop-in for the reference's pybind module `pointnet2.pointnet2_cuda`
(pointnet2/src/pointnet2_api.cpp:10-24).  Put `<repo>/dropin` on sys.path AHEAD of the
reference checkout's own (unbuilt) `pointnet2/` directory and the reference's
pointnet2/pointnet2_utils.py runs unchanged on the sm_100a kernels of libi2p_b200.so.

Same function names, positional signatures and in-place output convention; `knn_wrapper`,
which pointnet2_utils.py:32 calls but the reference never binds, is provided too.
"""
from i2pnet_b200 import _cabi


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _cabi.ball_query(b, n, m, radius, nsample, new_xyz, xyz, idx)
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _cabi.group_points(b, c, n, npoints, nsample, points, idx, out)
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _cabi.group_points_grad(b, c, n, npoints, nsample, grad_out, idx, grad_points)
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _cabi.gather_points(b, c, n, npoints, points, idx, out)
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _cabi.gather_points_grad(b, c, n, npoints, grad_out, idx, grad_points)
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _cabi.furthest_point_sampling(b, n, m, points, temp, idx)
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _cabi.three_nn(b, n, m, unknown, known, dist2, idx)


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _cabi.three_interpolate(b, c, m, n, points, idx, weight, out)


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _cabi.three_interpolate_grad(b, c, n, m, grad_out, idx, weight, grad_points)


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    _cabi.knn(b, n, m, k, unknown, known, dist2, idx)
