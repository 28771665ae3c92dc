"""Pose metrics (i2pnet_b200/metric.py) against the reference's own metric.py (RteRreEval, cal_rete_once) on seeded
random poses (tests/golden/make_golden.py metric), and the distance between this implementation's regressed pose and
the reference's expressed in those units (BASELINE.json: "pose RTE/RRE vs ref")."""
import os

import numpy as np
import torch

from tests.conftest import GOLDEN


def test_rte_rre_match_reference_metric():
    from i2pnet_b200 import metric
    g = np.load(os.path.join(GOLDEN, "ref_metric.npz"))
    pred, gt = g["pred"], g["gt"]
    for name, kw in (("plain", {}), ("thresholded", dict(threshold=True, rre_th=10., rte_th=5.))):
        ev = metric.RteRreEval(**kw)
        for i in range(0, 64, 16):
            ev.addBatch(metric.pose_to_extrinsic(pred[i:i + 16]), metric.pose_to_extrinsic(gt[i:i + 16]))
        assert np.allclose(ev.r_diff_all, g[name + "_rre_all"], rtol=1e-9, atol=1e-9)
        assert np.allclose(ev.t_diff_all, g[name + "_rte_all"], rtol=1e-9, atol=1e-12)
        assert np.allclose(np.array(ev.evalSeq()), g[name + "_seq"], rtol=1e-9)
        assert ev.get_recall() == float(g[name + "_recall"])
    assert 0.3 < float(g["thresholded_recall"]) < 1.0                     # the fixture has hits and misses
    once = metric.cal_rete_once(torch.from_numpy(pred), torch.from_numpy(gt[:, :4]), torch.from_numpy(gt[:, 4:]))
    assert np.allclose(np.array(once), g["once"], rtol=1e-9)


def pose_distance_to_reference(out3, ref_out3):
    """-> (max RRE in degrees, max RTE in metres) between two sets of regressed poses"""
    from i2pnet_b200 import metric
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    r, t = metric.rre_rte(metric.pose_to_extrinsic(to_np(out3)), metric.pose_to_extrinsic(to_np(ref_out3)))
    return float(r.max()), float(t.max())
