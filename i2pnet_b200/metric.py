"""Pose metrics of the evaluation path (mirror of the reference's metric.py: quat_to_rotmat_batch :9,
mult_extrinsic_batch :38, inv_extrinsic :54, cal_rete_once :125, RteRreEval :205-272).

Host-side numpy, like the reference's: relative translation error RTE = |t| of P_pred^-1 P_gt (metres) and relative
rotation error RRE = sum of the absolute 'xzy' Euler angles of its rotation (degrees).  BASELINE.json's metric pairs the
throughput with "pose RTE/RRE vs ref": tests use this module to express the distance between this implementation's
regressed pose and the reference's in those units.
"""
import math

import numpy as np
from scipy.spatial.transform import Rotation


def quat_to_rotmat_batch(q):
    """q (B,4) as (w,x,y,z) -> (B,3,3); not normalised first, exactly like the reference"""
    q = np.asarray(q)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = np.stack([1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
                    2 * x * y + 2 * z * w, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * x * w,
                    2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x ** 2 - 2 * y ** 2], axis=-1)
    return rot.reshape(-1, 3, 3)


def _homogeneous(m):
    m = np.asarray(m)
    bottom = np.broadcast_to(np.array([0., 0., 0., 1.]).reshape(1, 1, 4), (m.shape[0], 1, 4))
    return np.concatenate([m, bottom], axis=-2)


def mult_extrinsic_batch(m1, m2):
    """(B,3,4) x (B,3,4) -> (B,3,4), as 4x4 rigid transforms"""
    return (_homogeneous(m1) @ _homogeneous(m2))[:, :3, :]


def inv_extrinsic(m):
    return np.linalg.inv(_homogeneous(m))[:, :3, :]


def pose_to_extrinsic(pose7):
    """pose (B,7) = (q (w,x,y,z), t) as the network regresses it -> (B,3,4) (getExtrinsic :104-113, `out_raw` form)"""
    pose7 = np.asarray(pose7)          # the rotation matrix is formed in the pose's own precision, like the reference
    return np.concatenate([quat_to_rotmat_batch(pose7[:, :4]), pose7[:, 4:].reshape(-1, 3, 1)], axis=-1)


def rre_rte(pred_extrinsic, gt_extrinsic):
    """-> (angles_diff (B,) degrees, t_diff (B,) metres) of P_pred^-1 P_gt"""
    diff = mult_extrinsic_batch(inv_extrinsic(pred_extrinsic), gt_extrinsic)
    t_diff = np.linalg.norm(diff[:, :3, 3], 2, -1)
    angles = Rotation.from_matrix(diff[:, :3, :3]).as_euler('xzy', degrees=True)
    return np.sum(np.abs(angles), -1), t_diff


def cal_rete_once(out3, q_gt, t_gt):
    """Mean RRE (deg) and RTE (m) of a batch of regressed poses out3 (B,7) against the ground-truth decalibration."""
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    gt = np.concatenate([to_np(q_gt).reshape(-1, 4), to_np(t_gt).reshape(-1, 3)], axis=1)
    r, t = rre_rte(pose_to_extrinsic(to_np(out3)), pose_to_extrinsic(gt))
    return r.mean(), t.mean()


class RteRreEval(object):
    """Accumulates RRE / RTE over batches; with `threshold`, only pairs below (rre_th, rte_th) enter the statistics
    and get_recall() is their share."""

    def __init__(self, threshold=False, rre_th=10., rte_th=5.):
        self.t_diff, self.r_diff, self.t_diff_all, self.r_diff_all = [], [], [], []
        self.threshold, self.rre_th, self.rte_th = threshold, rre_th, rte_th
        self.acc_count = self.all_count = 0

    def reset(self):
        self.t_diff.clear()
        self.r_diff.clear()
        self.acc_count = self.all_count = 0

    def get_recall(self):
        return self.acc_count / self.all_count

    def addBatch(self, pred_extrinsic, gt_extrinsic):
        """pred / gt (B,3,4) -> (list of RRE, list of RTE) of this batch"""
        angles_diff, t_diff = rre_rte(pred_extrinsic, gt_extrinsic)
        self.all_count += len(angles_diff)
        keep = np.ones(len(angles_diff), dtype=bool)
        if self.threshold:
            keep = np.logical_and(t_diff < self.rte_th, angles_diff < self.rre_th)
        self.acc_count += int(keep.sum())
        self.t_diff.extend(list(t_diff[keep]))
        self.r_diff.extend(list(angles_diff[keep]))
        self.t_diff_all.extend(list(t_diff))
        self.r_diff_all.extend(list(angles_diff))
        return list(angles_diff), list(t_diff)

    def evalSeq(self):
        """-> RTE mean, RTE sigma, RRE mean, RRE sigma"""
        t_diff, r_diff = np.array(self.t_diff), np.array(self.r_diff)
        return t_diff.mean(), math.sqrt(np.var(t_diff)), r_diff.mean(), math.sqrt(np.var(r_diff))

    def save_metric(self, path):
        np.savez(path, RRE=np.array(self.r_diff_all), RTE=np.array(self.t_diff_all))
