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


def test_evaluation_loop_host_logic(oracle_backend):
    """i2pnet_b200.evaluation.evaluate (evaluation_proj.py's loop) on the reference-initialised model and the golden
    batch.  Like evaluation_proj.py:214 it switches the model to eval() (dropout off, the image branch's tracked
    BatchNorms on their running statistics, which it must not update) and restores the caller's mode: its RTE / RRE
    equal cal_rete_once of the model's own eval-mode forward; with running statistics equal to the batch's, that of
    the golden (train-mode) outputs of the reference model."""
    from i2pnet_b200 import metric
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg
    from i2pnet_b200.evaluation import evaluate
    from tests.test_host_logic_cpu import build_model, load_golden_model
    g, state = load_golden_model()
    model = build_model(state)
    t = lambda k: torch.from_numpy(g[k])
    batch = dict(rgb=torch.from_numpy(g["rgb_u8"]).float(), lidar=t("lidar"), raw_point_xyz=t("raw_point_xyz"),
                 lidar_feats=t("lidar_feats"), intrinsic=t("intrinsic"), q_gt=t("q_gt"), t_gt=t("t_gt"))
    assert model.training
    buffers = {k: v.clone() for k, v in model.named_buffers()}
    res = evaluate(model, [batch], cfg)
    assert model.training, "evaluate() must restore the caller's mode"
    assert all(torch.equal(v, buffers[k]) for k, v in model.named_buffers()), "evaluate() updated running statistics"
    model.eval()
    with torch.no_grad():
        own = model(batch["rgb"], batch["lidar"], batch["raw_point_xyz"], None, batch["intrinsic"], None, None, None,
                    batch["lidar_feats"], cfg)[0]
    model.train()
    want_rre, want_rte = metric.cal_rete_once(own, t("q_gt"), t("t_gt"))
    assert abs(res["rre_mean"] - want_rre) < 1e-3 and abs(res["rte_mean"] - want_rte) < 1e-4
    assert res["recall"] == 1.0 and len(res["ms_per_batch"]) == 1 and len(res["rre"]) == 2
