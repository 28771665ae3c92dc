"""Generate the golden vectors of tests/golden/ by running the REAL reference.

Runs in the build container only (needs /root/reference; no GPU).  The reference's Python is
imported unchanged; its two CUDA extension modules are replaced by CPU stand-ins backed by the
C oracle (oracle/ref_shims.py), everything else -- project_seq, gather_torch, knn_point,
CostVolume, ProjectPointNet, RegNet_v2, Get_loss -- is the reference's own code executing on
CPU tensors.  The resulting fixtures pin (a) oracle/model_cpu.py and (b) the CUDA path of
i2pnet_b200 against the reference's module-level behaviour.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

ref_shims.install("/root/reference")

from i2pnet_b200.synthetic import make_pairs  # noqa: E402  (input generator only; no operators)


def _save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote %s (%.2f MB)" % (path, os.path.getsize(path) / 1e6))


def golden_ops():
    """Op-level vectors from the reference's own Python helpers (src/projectPN/utils.py)."""
    from src.projectPN import utils as U
    g = torch.Generator().manual_seed(1)
    d = make_pairs(2, n_points=4096, seed=3)
    xyz_proj, (f1, f2) = U.project_seq(d["raw_point_xyz"], [d["lidar_feats"], d["lidar"]], 64, 1800, False, 2.0, -24.8)
    nz = xyz_proj.abs().sum(-1) > 0
    # project_seq: store sparsely (cell index + value) to keep the fixture small
    _save("ref_project_seq.npz", raw=d["raw_point_xyz"], feats=d["lidar_feats"], cam=d["lidar"],
          cells=nz.nonzero(), xyz=xyz_proj[nz], f1=f1[nz], f2=f2[nz])

    # get_neighbor_copy / get_neighbor_att + gather_torch on a small range image
    B, H, W = 2, 16, 225
    img = torch.zeros(B, H, W, 3)
    keep = torch.rand(B, H, W, generator=g) < 0.6
    img[keep] = (torch.rand(int(keep.sum()), 3, generator=g) - 0.5) * 20
    idx_n2 = U.get_stride_idx_cuda(B, 8, 113, 2, 2, "cpu")
    feat = torch.randn(B, H, W, 5, generator=g)
    res = {}
    for name, fn in (("copy", U.get_neighbor_copy), ("att", U.get_neighbor_att)):
        b, h, w, m = fn(img, img, idx_n2, [5, 9], 8, distance=6.0)
        res["%s_h" % name], res["%s_w" % name], res["%s_mask" % name] = h, w, m
        res["%s_gather" % name] = U.gather_torch(feat, b, h, w, B, H, W)
    # up-conv form: xyz1 at 4x57, xyz2 at 4x29, stride (1,2)
    x1 = img[:, :4, :57].contiguous()
    x2 = x1[:, :, ::2].contiguous()
    b, h, w, m = U.get_neighbor_copy(x1, x2, U.get_idx_cuda(B, 4, 57, "cpu"), [5, 9], 8, 1, 2, distance=9.0)
    res.update(up_h=h, up_w=w, up_mask=m)
    q = torch.randn(B, 50, 3, generator=g)
    s = torch.randn(B, 300, 3, generator=g)
    res["knn_idx_sorted"] = torch.sort(U.knn_point(16, s, q), dim=-1)[0]
    _save("ref_select_gather.npz", img=img, idx_n2=idx_n2, feat=feat, knn_q=q, knn_s=s, **res)


def golden_model():
    """Whole-model forward + loss + backward of the reference RegNet_v2 (KITTI shape, B=2)."""
    import compute_loss
    from src.config_proj_lidarcenter import I2PNetConfig as cfg
    from src.modellearn_proj_center import RegNet_v2
    cfg.efgh = False
    torch.manual_seed(0)
    model = RegNet_v2(cfg=cfg)
    model.train()
    for head in (model.l4_head, model.l3_head):   # training-mode BN everywhere, dropout silenced
        head.DP1.p = 0.0
    # non-trivial BN affine parameters so that their gradients are exercised
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "bn" in n or (".1." in n and "RGB" in n) or ".5." in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    state = {k: v.clone() for k, v in model.state_dict().items()}

    d = make_pairs(2, n_points=20480, seed=11, occupy_centres=(4, 8))
    inter = {}
    for name in ("LiDAR_lv2", "LiDAR_lv3", "cost_volume1", "layer_idx", "set_upconv0_upsample", "cost_volume2"):
        def hook(mod, args, out, name=name):
            t = out[2] if isinstance(out, tuple) else out
            inter["inter_" + name] = t.detach().clone()
        getattr(model, name).register_forward_hook(hook)
    out3, out4, _, _, sx, sq = model(d["rgb"], d["lidar"], d["raw_point_xyz"], None, d["intrinsic"], None, None,
                                     None, d["lidar_feats"], cfg)
    loss, lq, lx = compute_loss.Get_loss(out3, out4, d["q_gt"], d["t_gt"], sx, sq, cfg)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    names = sorted(grads)
    keep = ["sq", "sx", "l3_head.quat_head.composed_module.0.weight", "cost_volume1.mlp1_convs.0.conv.weight",
            "LiDAR_lv1.mlp_convs.0.conv.weight", "RGB_net1.0.weight", "cost_volume2.pc_encoding.bn_linear.weight"]
    inter["inter_LiDAR_lv2"] = inter["inter_LiDAR_lv2"][:, ::2, ::7]   # subsample the big one
    _save("ref_model_kitti_b2.npz",
          rgb_u8=d["rgb"].to(torch.uint8), lidar=d["lidar"], raw_point_xyz=d["raw_point_xyz"],
          lidar_feats=d["lidar_feats"], intrinsic=d["intrinsic"], q_gt=d["q_gt"], t_gt=d["t_gt"],
          out3=out3, out4=out4, loss=loss, grad_names=np.array(names),
          grad_norms=np.array([float(grads[n].norm()) for n in names]),
          **{"grad__" + n: grads[n] for n in keep}, **{"state__" + k: v for k, v in state.items()}, **inter)


def golden_iter():
    """Forward of the reference's iterative-refinement model (src/modellearn_proj_center_iter.py, six level-3
    iterations) on the inputs and the state dict of ref_model_kitti_b2.npz; only the outputs are stored."""
    from src.config_proj_lidarcenter import I2PNetConfig as cfg
    from src.modellearn_proj_center_iter import RegNet_v2
    cfg.efgh = False
    g = np.load(os.path.join(HERE, "ref_model_kitti_b2.npz"))
    state = {k[len("state__"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state__")}
    model = RegNet_v2(cfg=cfg)
    model.load_state_dict(state, strict=True)
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    t = lambda k: torch.from_numpy(g[k])
    # the neighbour sets the reference itself selected in its six refinements (knn_point of the second cost volume: the 32
    # pixels nearest to every point warped by the previous pose): the GPU test pins them, so that it compares arithmetic
    # and not which side of a near-tie a 1e-6 pose difference falls on
    from src.projectPN import utils as RP       # grouping() -> knn_point(), src/projectPN/utils.py:329, 369
    sets, orig = [], RP.knn_point

    def record(nsample, xyz, new_xyz):
        idx = orig(nsample, xyz, new_xyz)
        sets.append(idx.clone())
        return idx
    RP.knn_point = record
    try:
        with torch.no_grad():
            out3, out4, _, _, _, _ = model(t("rgb_u8").float(), t("lidar"), t("raw_point_xyz"), None, t("intrinsic"), None, None,
                                           None, t("lidar_feats"), cfg)
    finally:
        RP.knn_point = orig
    assert len(sets) == 6 and all(s.shape == sets[0].shape for s in sets) and int(torch.stack(sets).max()) < 32768
    _save("ref_model_iter_kitti_b2.npz", out3=out3, out4=out4, knn_sets=torch.stack(sets).to(torch.int16))


def golden_nus():
    """Forward + loss + gradient norms of the reference RegNet_v2 at the nuScenes shape (BASELINE.json configs[3]:
    320x640 image, 21x1800 range image of src/config_proj_lidarcenter_nus.py), B = 2, 20480 distinct cells."""
    import compute_loss
    from src.config_proj_lidarcenter_nus import I2PNetConfig as cfg
    from src.modellearn_proj_center import RegNet_v2
    cfg.efgh = False
    torch.manual_seed(1)
    model = RegNet_v2(cfg=cfg)
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "bn" in n or (".1." in n and "RGB" in n) or ".5." in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    d = make_pairs(2, n_points=20480, image_hw=(320, 640), init_H=cfg.init_H, init_W=cfg.init_W, fup=cfg.fup, fdown=cfg.fdown,
                   seed=13, occupy_centres=(cfg.stride_Hs[0], cfg.stride_Ws[0]))
    inter = {}
    for name in ("LiDAR_lv2", "LiDAR_lv3", "cost_volume1", "layer_idx", "set_upconv0_upsample", "cost_volume2"):
        def hook(mod, args, out, name=name):
            t = out[2] if isinstance(out, tuple) else out
            inter["inter_" + name] = t.detach().clone()
        getattr(model, name).register_forward_hook(hook)
    out3, out4, _, _, sx, sq = model(d["rgb"], d["lidar"], d["raw_point_xyz"], None, d["intrinsic"], None, None,
                                     None, d["lidar_feats"], cfg)
    loss, lq, lx = compute_loss.Get_loss(out3, out4, d["q_gt"], d["t_gt"], sx, sq, cfg)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    names = sorted(grads)
    keep = ["sq", "sx", "l3_head.quat_head.composed_module.0.weight", "cost_volume1.mlp1_convs.0.conv.weight",
            "LiDAR_lv1.mlp_convs.0.conv.weight", "cost_volume2.pc_encoding.bn_linear.weight"]
    inter["inter_LiDAR_lv2"] = inter["inter_LiDAR_lv2"][:, ::2, ::7]
    _save("ref_model_nus_b2.npz",
          rgb_u8=d["rgb"].to(torch.uint8), lidar=d["lidar"], raw_point_xyz=d["raw_point_xyz"],
          lidar_feats=d["lidar_feats"], intrinsic=d["intrinsic"], q_gt=d["q_gt"], t_gt=d["t_gt"],
          out3=out3, out4=out4, loss=loss, grad_names=np.array(names),
          grad_norms=np.array([float(grads[n].norm()) for n in names]),
          **{"grad__" + n: grads[n] for n in keep}, **{"state__" + k: v for k, v in state.items()}, **inter)


def golden_metric():
    """RTE / RRE of the reference's own metric.py (RteRreEval, cal_rete_once) on seeded random poses."""
    import types
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    msee = types.ModuleType("src.util.lie_metric.MSEE")          # needs geomstats (absent); unused by RteRreEval
    msee.SE3_to_se3 = msee.cal_metric = None
    sys.modules.setdefault("src.util.lie_metric.MSEE", msee)
    import metric as ref_metric
    rng = np.random.Generator(np.random.PCG64(77))

    def poses(n, ang, tr):
        axis = rng.standard_normal((n, 3))
        axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        half = np.deg2rad(rng.uniform(0, ang, n)) / 2
        q = np.concatenate([np.cos(half)[:, None], axis * np.sin(half)[:, None]], 1)
        return np.concatenate([q, rng.uniform(-tr, tr, (n, 3))], 1).astype(np.float32)
    pred, gt = poses(64, 40.0, 8.0), poses(64, 40.0, 8.0)
    pred[:32] = gt[:32] + rng.normal(0, 2e-3, (32, 7)).astype(np.float32)      # near-hits as well as misses
    to_ext = lambda p: np.concatenate([ref_metric.quat_to_rotmat_batch(p[:, :4]), p[:, 4:].reshape(-1, 3, 1)], -1)
    out = {}
    for name, kw in (("plain", {}), ("thresholded", dict(threshold=True, rre_th=10., rte_th=5.))):
        ev = ref_metric.RteRreEval(**kw)
        for i in range(0, 64, 16):
            ev.addBatch(to_ext(pred[i:i + 16]), to_ext(gt[i:i + 16]))
        out[name + "_seq"] = np.array(ev.evalSeq())
        out[name + "_recall"] = ev.get_recall()
        out[name + "_rre_all"], out[name + "_rte_all"] = np.array(ev.r_diff_all), np.array(ev.t_diff_all)
    data_valid = {"decalib_real_gt": torch.from_numpy(gt[:, :4]), "decalib_dual_gt": torch.from_numpy(gt[:, 4:])}
    out["once"] = np.array(ref_metric.cal_rete_once(torch.from_numpy(pred), data_valid))
    _save("ref_metric.npz", pred=pred, gt=gt, **out)


def golden_configs():
    """Every plain hyper-parameter of the reference's three shipped configurations, as JSON."""
    import json
    import types
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    from src.config_lidarcenter import I2PNetConfig as small
    from src.config_proj_lidarcenter import I2PNetConfig as kitti
    from src.config_proj_lidarcenter_nus import I2PNetConfig as nus

    def plain(cls):
        out = {}
        for k, v in vars(cls).items():
            if k.startswith("__") or callable(v):
                continue
            if hasattr(v, "name") and hasattr(v, "value"):
                v = "enum:" + v.name
            try:
                json.dumps(v)
            except TypeError:
                continue
            out[k] = v
        return out
    path = os.path.join(HERE, "ref_configs.json")
    with open(path, "w") as fh:
        json.dump({"config_lidarcenter": plain(small), "config_proj_lidarcenter": plain(kitti),
                   "config_proj_lidarcenter_nus": plain(nus)}, fh, indent=1, sort_keys=True)
    print("wrote", path)


def golden_small():
    """Forward + loss + backward of the reference's small-range RegNet_v2 (src/modellearn.py with
    src/config_lidarcenter.py: 8192 points -> 2048 / 1024 / 256 / 64, 160x512 image), B = 2, training mode."""
    import types
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))      # src/utils.py imports it; unused on this path
    import compute_loss
    from src.config_lidarcenter import I2PNetConfig as cfg
    from src.modellearn import RegNet_v2
    from i2pnet_b200.synthetic import make_pairs_small
    torch.manual_seed(0)
    model = RegNet_v2(cfg=cfg)
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "bn" in n or (".1." in n and "RGB" in n) or ".5." in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    d = make_pairs_small(2, n_points=8192, seed=21)
    inter = {}
    for name in ("LiDAR_lv1", "LiDAR_lv3", "cost_volume1", "layer_idx", "set_upconv0_upsample", "cost_volume2"):
        def hook(mod, args, out, name=name):
            t = out[1] if isinstance(out, tuple) else out
            inter["inter_" + name] = t.detach().clone()
        getattr(model, name).register_forward_hook(hook)
    out3, out4, _, _, sx, sq = model(d["rgb"], d["lidar"], None, d["intrinsic"], None, None, None, None, cfg=cfg,
                                     lidar_img_raw=d["raw_point_xyz"])
    loss, lq, lx = compute_loss.Get_loss(out3, out4, d["q_gt"], d["t_gt"], sx, sq, cfg)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    names = sorted(grads)
    keep = ["sq", "sx", "l3_head.quat_head.composed_module.0.weight", "cost_volume1.mlp1_convs.0.conv.weight",
            "LiDAR_lv1.mlp_convs.0.weight", "RGB_net1.0.weight", "cost_volume2.pc_encoding.bn_linear.weight"]
    inter["inter_LiDAR_lv1"] = inter["inter_LiDAR_lv1"][:, :, ::8]
    after = {k: v for k, v in model.state_dict().items() if "running" in k and ("LiDAR_lv3" in k or "cost_volume1.mlp1_convs.0" in k)}
    # eval(): every norm uses the running statistics this one training step has just blended in
    state_after = {k: v.clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        e3, e4, _, _, _, _ = model(d["rgb"], d["lidar"], None, d["intrinsic"], None, None, None, None, cfg=cfg,
                                   lidar_img_raw=d["raw_point_xyz"])
    _save("ref_model_small_eval_b2.npz", out3=e3, out4=e4,
          **{"state__" + k: v for k, v in state_after.items() if "running" in k or "num_batches" in k})
    _save("ref_model_small_b2.npz",
          rgb_u8=d["rgb"].to(torch.uint8), lidar=d["lidar"], raw_point_xyz=d["raw_point_xyz"], intrinsic=d["intrinsic"],
          q_gt=d["q_gt"], t_gt=d["t_gt"], out3=out3, out4=out4, loss=loss, grad_names=np.array(names),
          grad_norms=np.array([float(grads[n].norm()) for n in names]),
          **{"grad__" + n: grads[n] for n in keep}, **{"state__" + k: v for k, v in state.items()},
          **{"after__" + k: v for k, v in after.items()}, **inter)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "iter":
        golden_iter()
    elif len(sys.argv) > 1 and sys.argv[1] == "small":
        golden_small()
    elif len(sys.argv) > 1 and sys.argv[1] == "nus":
        golden_nus()
    elif len(sys.argv) > 1 and sys.argv[1] == "metric":
        golden_metric()
    elif len(sys.argv) > 1 and sys.argv[1] == "configs":
        golden_configs()
    else:
        golden_ops()
        golden_model()
        golden_iter()
        golden_small()
        golden_nus()
        golden_metric()
        golden_configs()
