"""Seeded synthetic (image, point cloud) pairs with the shapes of the reference's large-range
loaders (src/kitti_odometry_corr_lidarnone_proj.py:770-789: rgb, lidar, raw_point_xyz,
lidar_feats, init_intrinsic, decalib_*_gt).  There are no datasets offline; this fixes the INPUT
CONTRACT of the hot path, nothing more.

Every point is placed at the centre of its own cell of the init_H x init_W range image (cells
drawn without replacement), so that
  * the truncating cell formulas of project_seq (src/projectPN/utils.py:147-155) are evaluated
    far from any rounding boundary -- CPU and GPU libm agree on the cell;
  * no two points share a cell -- the reference's duplicate-cell scatter is unordered on CUDA
    (cfg.rank = False) and is fenced off from parity inputs.
numpy's PCG64 stream is platform independent, so the same seed gives the same pairs here and
on the GPU box.
"""
import numpy as np
import torch


def _quat_from_matrix(R):
    """3x3 rotation -> (w,x,y,z), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q if q[0] >= 0 else -q


def make_pairs(batch, n_points=20480, image_hw=(160, 512), init_H=64, init_W=1800, fup=2.0, fdown=-24.8, seed=0,
               max_yaw_deg=360.0, max_trans=10.0, occupy_centres=None):
    """-> dict of CPU float32 tensors: rgb (B,3,h,w) in [0,255]; lidar (B,N,3) camera frame after
    the random decalibration; raw_point_xyz (B,N,3) LiDAR frame; lidar_feats (B,N,1); intrinsic
    (B,3,3); q_gt (B,4), t_gt (B,3) the pose that undoes the decalibration.
    occupy_centres=(stride_h, stride_w): additionally guarantee a point in every cell of that
    stride grid (the level-1 centres, a superset of every coarser level's).  Parity inputs use it:
    an empty centre is an all-zero query point, which ties EXACTLY between mirror-image pixels in
    the cost volume's kNN, where torch.topk's choice is unspecified."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w = image_hw
    cells = init_H * init_W
    replace = n_points > cells  # nuScenes-shape stress config: 40960 points on 37800 cells
    deg = np.pi / 180.0
    az = 360.0 / init_W * deg
    vres = (fup - fdown) * deg / (init_H - 1)
    # LiDAR (x fwd, y left, z up) -> camera (x right, y down, z fwd)
    Pc = np.array([[0., -1., 0.], [0., 0., -1.], [1., 0., 0.]])
    out = {k: [] for k in ("rgb", "lidar", "raw_point_xyz", "lidar_feats", "intrinsic", "q_gt", "t_gt")}
    for _ in range(batch):
        out["rgb"].append(rng.integers(0, 256, size=(3, h, w)).astype(np.float32))
        if occupy_centres is not None:
            rr, cc = np.meshgrid(np.arange(0, init_H, occupy_centres[0]), np.arange(0, init_W, occupy_centres[1]),
                                 indexing="ij")
            forced = (rr * init_W + cc).reshape(-1)
            rest = np.setdiff1d(np.arange(cells), forced)
            cell = np.concatenate([forced, rng.choice(rest, size=n_points - forced.size, replace=False)])
            cell = cell[rng.permutation(n_points)]
        else:
            cell = rng.choice(cells, size=n_points, replace=replace)
        row, col = cell // init_W, cell % init_W
        alpha = np.pi - (col + 0.5) * az                     # iCol = (pi - atan2(y,x)) / az
        beta = fdown * deg + (init_H - row + 0.5) * vres     # iRow = H - int((beta - down) / vres)
        r = np.clip(rng.lognormal(np.log(12.0), 0.7, size=n_points), 2.0, 80.0)
        raw = np.stack([r * np.cos(beta) * np.cos(alpha), r * np.cos(beta) * np.sin(alpha), r * np.sin(beta)], -1)
        yaw = rng.uniform(0.0, max_yaw_deg) * deg
        Rr = np.array([[np.cos(yaw), 0., np.sin(yaw)], [0., 1., 0.], [-np.sin(yaw), 0., np.cos(yaw)]])
        tr = np.array([rng.uniform(-max_trans, max_trans), 0.0, rng.uniform(-max_trans, max_trans)])
        cam = raw @ Pc.T
        out["lidar"].append((cam @ Rr.T + tr).astype(np.float32))
        out["raw_point_xyz"].append(raw.astype(np.float32))
        out["lidar_feats"].append(rng.uniform(0.0, 1.0, size=(n_points, 1)).astype(np.float32))
        out["intrinsic"].append(np.array([[360.8, 0., w / 2.0], [0., 360.8, h / 2.0], [0., 0., 1.]], np.float32))
        out["q_gt"].append(_quat_from_matrix(Rr.T).astype(np.float32))
        out["t_gt"].append((-Rr.T @ tr).astype(np.float32))
    return {k: torch.from_numpy(np.stack(v)) for k, v in out.items()}


def make_pairs_small(batch, n_points=8192, image_hw=(160, 512), seed=0, max_rot_deg=10.0, max_trans=1.0):
    """Input contract of the small-range model (src/kitti_odometry_cmr.py: rgb, lidar, raw_point_xyz,
    init_intrinsic, decalib_*_gt): the cloud is cropped to the camera frustum, in front of the camera, and
    mis-calibrated by a small rotation / translation.  -> the same keys as make_pairs (lidar_feats is zeros).
    Points: depth log-uniform in [3, 50] m, bearing uniform over the image's field of view."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w = image_hw
    f = 360.8
    Pc = np.array([[0., -1., 0.], [0., 0., -1.], [1., 0., 0.]])      # LiDAR -> camera axes
    deg = np.pi / 180.0
    out = {k: [] for k in ("rgb", "lidar", "raw_point_xyz", "lidar_feats", "intrinsic", "q_gt", "t_gt")}
    for _ in range(batch):
        out["rgb"].append(rng.integers(0, 256, size=(3, h, w)).astype(np.float32))
        z = np.exp(rng.uniform(np.log(3.0), np.log(50.0), size=n_points))
        u = rng.uniform(-0.5 * w / f, 0.5 * w / f, size=n_points)
        v = rng.uniform(-0.5 * h / f, 0.5 * h / f, size=n_points)
        cam = np.stack([u * z, v * z, z], -1)
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(0.0, max_rot_deg) * deg
        Kx = np.array([[0., -axis[2], axis[1]], [axis[2], 0., -axis[0]], [-axis[1], axis[0], 0.]])
        Rr = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        tr = rng.uniform(-max_trans, max_trans, size=3)
        out["lidar"].append((cam @ Rr.T + tr).astype(np.float32))
        out["raw_point_xyz"].append((cam @ Pc).astype(np.float32))    # Pc^-1 = Pc^T, row vectors: cam @ Pc
        out["lidar_feats"].append(np.zeros((n_points, 1), np.float32))
        out["intrinsic"].append(np.array([[f, 0., w / 2.0], [0., f, h / 2.0], [0., 0., 1.]], np.float32))
        out["q_gt"].append(_quat_from_matrix(Rr.T).astype(np.float32))
        out["t_gt"].append((-Rr.T @ tr).astype(np.float32))
    return {k: torch.from_numpy(np.stack(v)) for k, v in out.items()}
