"""Evaluation loop of the large-range models (the core of the reference's evaluation_proj.py:229-300: per batch a
no-grad forward timed with a synchronise on both sides, then RTE / RRE of the regressed decalibration against the
ground truth through metric.RteRreEval).  Dataset access, logging and visualisation of the reference script are out
of scope; batches are dicts with the loader's keys (i2pnet_b200.synthetic.make_pairs emits them)."""
import time

import torch

from .metric import RteRreEval, pose_to_extrinsic


def evaluate(model, batches, cfg, device=None, threshold=False):
    """model: RegNet_v2 (single-pass or iterative-refinement); batches: iterable of dicts of CPU or device tensors
    (rgb, lidar, raw_point_xyz, lidar_feats, intrinsic, q_gt, t_gt).
    -> dict(rte_mean, rte_sigma, rre_mean, rre_sigma, recall, ms_per_batch, rre, rte)"""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    evaluator = RteRreEval(threshold=threshold)
    times = []
    was_training = model.training
    model.eval()            # evaluation_proj.py:214: dropout off, tracked BatchNorms on their running statistics
    try:
        _run(model, batches, cfg, device, evaluator, times)
    finally:
        model.train(was_training)
    rte_mean, rte_sigma, rre_mean, rre_sigma = evaluator.evalSeq()
    return dict(rte_mean=rte_mean, rte_sigma=rte_sigma, rre_mean=rre_mean, rre_sigma=rre_sigma,
                recall=evaluator.get_recall(), ms_per_batch=[1e3 * t for t in times],
                rre=list(evaluator.r_diff_all), rte=list(evaluator.t_diff_all))


def _run(model, batches, cfg, device, evaluator, times):
    with torch.no_grad():
        for data in batches:
            if device.type == "cuda":
                torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            x = {k: v.to(device, non_blocking=True) for k, v in data.items()}
            out3 = model(x["rgb"], x["lidar"], x["raw_point_xyz"].float(), None, x["intrinsic"], None, None, None,
                         x["lidar_feats"].float(), cfg)[0]
            if device.type == "cuda":
                torch.cuda.synchronize(device)
            times.append(time.perf_counter() - t0)          # includes the host-to-device copies, like the reference's
            gt = torch.cat([data["q_gt"].reshape(-1, 4), data["t_gt"].reshape(-1, 3)], dim=1)
            evaluator.addBatch(pose_to_extrinsic(out3.cpu().numpy()), pose_to_extrinsic(gt.cpu().numpy()))
