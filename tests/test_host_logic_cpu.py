"""Host-side logic of the product modules (state_dict layout, data flow, autograd wiring),
checked WITHOUT a GPU against the golden vectors recorded from the real reference model
(tests/golden/make_golden.py).  The CUDA kernels are replaced by the C oracle through the
`oracle_backend` fixture; the same comparison with the real kernels is tests/test_model_gpu.py."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

REL = 1e-4  # north_star: fp32 features and regressed pose within 1e-4 relative


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden_model():
    g = np.load(os.path.join(GOLDEN, "ref_model_kitti_b2.npz"))
    state = {k[len("state__"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state__")}
    return g, state


def build_model(state, device="cpu"):
    from i2pnet_b200.modellearn_proj_center import RegNet_v2
    model = RegNet_v2()
    missing = model.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    return model.to(device)


def run_model(model, g, device="cpu"):
    from i2pnet_b200.compute_loss import Get_loss
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg
    t = lambda k: torch.from_numpy(g[k]).to(device)
    rgb = torch.from_numpy(g["rgb_u8"]).float().to(device)
    inter = {}
    for name in ("LiDAR_lv2", "LiDAR_lv3", "cost_volume1", "layer_idx", "set_upconv0_upsample", "cost_volume2"):
        def hook(mod, args, out, name=name):
            inter[name] = (out[2] if isinstance(out, tuple) else out).detach()
        getattr(model, name).register_forward_hook(hook)
    out3, out4, _, _, sx, sq = model(rgb, t("lidar"), t("raw_point_xyz"), None, t("intrinsic"), None, None, None,
                                     t("lidar_feats"), cfg)
    loss, _, _ = Get_loss(out3, out4, t("q_gt"), t("t_gt"), sx, sq, cfg)
    loss.backward()
    return out3, out4, loss, inter


def _bias_cancelled_by_bn(name):
    """A conv bias feeding a batch-statistics BatchNorm has an exactly zero gradient; what both
    implementations produce there is rounding noise (1e-10..1e-5) and is not compared."""
    if name.endswith(".conv.bias"):
        return True
    parts = name.split(".")
    return parts[0].startswith("RGB_net") and parts[-1] == "bias" and int(parts[1]) % 4 == 0


def _points_outside(val, ref, tol):
    """Number of points (all axes but the last) whose largest channel error exceeds tol * max |ref|."""
    val, ref = torch.as_tensor(val).double(), torch.as_tensor(ref).double()
    err = (val - ref).abs().amax(-1) / ref.abs().max().clamp_min(1e-30)
    return int((err >= tol).sum()), err.numel()


def check_against_golden(model, g, out3, out4, loss, inter, tol=REL, grad_tol=2e-3, knn_flip_tol=None):
    """knn_flip_tol (GPU runs against the CPU recording only): the second cost volume selects, per point, the 32 image
    pixels nearest to the point warped by the COARSE POSE the network has just regressed (PPBackbone_center.py:369).
    That pose agrees with the recording to ~5e-6, not bit for bit, and at this input one of the 456 selections sits on
    a near-tie that such a perturbation flips (tools/debug_model_conv.py: library convolution vs own convolution, both
    f32, 1 of 456 sets).  Everything upstream of that selection is held to `tol`; downstream, at most 10 % of the points
    may leave `tol` (a
    flipped set reaches its 3 x 5 window of 3-D neighbours and, through the batch statistics, everything a little), and the refined pose / loss are held to knn_flip_tol.  The GPU <-> GPU comparison with the live
    reference at batch 8 / 32 (tests/test_reference_live_gpu.py) has no such exception."""
    inter["LiDAR_lv2"] = inter["LiDAR_lv2"][:, ::2, ::7]
    flipped = 0
    for name, val in inter.items():
        ref = g["inter_" + name]
        bad, total = _points_outside(val.cpu().reshape(ref.shape), ref, tol)
        if name == "cost_volume2" and knn_flip_tol is not None:
            assert bad <= total // 10, (name, bad, total)
            flipped = bad
        else:
            assert bad == 0, (name, bad, total, _rel(val.cpu().reshape(ref.shape), ref))
    assert _rel(out4.detach().cpu(), g["out4"]) < tol
    tol3 = knn_flip_tol if flipped else tol
    assert _rel(out3.detach().cpu(), g["out3"]) < tol3, (flipped, _rel(out3.detach().cpu(), g["out3"]))
    assert abs(float(loss) - float(g["loss"])) < tol3 * abs(float(g["loss"]))
    check_pose_distance(out3, g["out3"])
    check_pose_distance(out4, g["out4"])
    grads = {n: p.grad for n, p in model.named_parameters()}
    names = [str(n) for n in g["grad_names"]]
    assert sorted(grads) == names
    for n, ref_norm in zip(names, g["grad_norms"]):
        if _bias_cancelled_by_bn(n):
            assert float(grads[n].norm()) < 1e-3
            continue
        mine = float(grads[n].norm())
        assert abs(mine - ref_norm) <= grad_tol * max(ref_norm, 1e-6) + 1e-7, (n, mine, ref_norm)
    for k in g.files:
        if k.startswith("grad__"):
            assert _rel(grads[k[len("grad__"):]].cpu(), g[k]) < 10 * grad_tol, k


def check_pose_distance(pose, ref_pose):
    """BASELINE.json's "pose RTE/RRE vs ref": the regressed pose against the reference's own, in the units of the
    reference's evaluation (metric.RteRreEval): well below a milli-degree / a tenth of a millimetre at 1e-5 relative."""
    from tests.test_metric_cpu import pose_distance_to_reference
    rre, rte = pose_distance_to_reference(pose, ref_pose)
    scale = max(1.0, float(np.abs(np.asarray(ref_pose)[:, 4:]).max()))
    assert rre < 0.05 and rte < 1e-3 * scale, (rre, rte)


def test_state_dict_matches_reference_layout():
    g, state = load_golden_model()
    from i2pnet_b200.modellearn_proj_center import RegNet_v2
    mine = RegNet_v2().state_dict()
    assert list(mine.keys()) == list(state.keys())           # same names, same registration order
    assert all(mine[k].shape == state[k].shape for k in mine)
    assert sum(p.numel() for p in RegNet_v2().parameters()) == 844896


def test_model_host_logic_matches_reference(oracle_backend):
    g, state = load_golden_model()
    model = build_model(state)
    out3, out4, loss, inter = run_model(model, g)
    check_against_golden(model, g, out3, out4, loss, inter)


def test_product_refuses_cpu_tensors():
    """No fallback: the real binding raises on a CPU tensor instead of computing anything."""
    from i2pnet_b200 import _cabi
    from i2pnet_b200.projectPN.utils import gather_rows
    with pytest.raises(_cabi.I2PError):
        gather_rows(torch.zeros(1, 4, 4), torch.zeros(1, 2, dtype=torch.int32))


def _run_iter_model(state, g, device):
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg
    from i2pnet_b200.modellearn_proj_center_iter import RegNet_v2
    model = RegNet_v2()
    model.load_state_dict(state, strict=True)         # same parameters as the single-pass model
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    model.to(device)
    t = lambda k: torch.from_numpy(g[k]).to(device)
    with torch.no_grad():
        out3, out4, _, _, _, _ = model(torch.from_numpy(g["rgb_u8"]).float().to(device), t("lidar"), t("raw_point_xyz"), None,
                                       t("intrinsic"), None, None, None, t("lidar_feats"), cfg)
    return out3, out4


def check_iter_model(device, knn_flip_tol=None, pin_knn_sets=False):
    """The six-iteration inference model against the reference's own
    src/modellearn_proj_center_iter.py run on the same inputs and weights (tests/golden/make_golden.py iter).
    Every refinement selects, per point, the 32 pixels nearest to the point warped by the PREVIOUS pose: 2736 selections
    in all, of which a handful sit on near-ties that a 1e-6 pose difference flips (1 of 456 in the single-pass model, see
    check_against_golden), each flip worth up to 2e-3 of the refined pose.
    pin_knn_sets: the selections are replaced by the ones the reference made (recorded in the fixture), so the run
    compares arithmetic and is held to the north-star tolerance; otherwise out3 is held to knn_flip_tol."""
    g, state = load_golden_model()
    ref = np.load(os.path.join(GOLDEN, "ref_model_iter_kitti_b2.npz"))
    from i2pnet_b200.projectPN import PPBackbone_center as P
    orig, calls = P.knn_point, []

    def replay(nsample, xyz, new_xyz):
        idx = torch.from_numpy(ref["knn_sets"][len(calls)].astype(np.int64)).to(new_xyz.device)
        calls.append(idx)
        assert idx.shape == (new_xyz.shape[0], new_xyz.shape[1], nsample)
        return idx
    try:
        if pin_knn_sets:
            P.knn_point = replay
        out3, out4 = _run_iter_model(state, g, device)
    finally:
        P.knn_point = orig
    assert not pin_knn_sets or len(calls) == 6
    assert _rel(out4.cpu(), ref["out4"]) < REL
    tol3 = REL if pin_knn_sets else (knn_flip_tol or REL)
    assert _rel(out3.cpu(), ref["out3"]) < tol3, _rel(out3.cpu(), ref["out3"])
    if pin_knn_sets or knn_flip_tol is None:
        check_pose_distance(out3, ref["out3"])
    assert _rel(out3.cpu(), g["out3"]) > 1e-3          # and it is not the single-pass answer


def test_iter_model_host_logic_matches_reference(oracle_backend):
    check_iter_model("cpu")


def test_pixel_rays_and_pyramid_shape_host_spelling():
    """pixel_rays (host spelling) = change_intrinsic -> inverse -> set_id_grid -> batched product of the reference
    (src/modellearn_proj_center.py:275-287); pyramid_out_hw predicts the feature-map size the pyramids produce."""
    from i2pnet_b200.modellearn_proj_center import change_intrinsic, set_id_grid
    from i2pnet_b200.modules.basicConv import createCNNs, pyramid_out_hw
    from i2pnet_b200.projectPN.utils import pixel_rays
    torch.manual_seed(0)
    nets = (createCNNs(3, [4, 4, 8], [2, 1, 2]), createCNNs(8, [8, 8], [2, 2]), createCNNs(8, [8], [1]))
    x = torch.randn(2, 3, 37, 50)
    y = x
    with torch.no_grad():
        for n in nets:
            y = n(y)
    assert tuple(y.shape[2:]) == pyramid_out_hw(nets, 37, 50)
    K = torch.tensor([[700.0, 0.0, 25.0], [0.0, 650.0, 18.0], [0.0, 0.0, 1.0]]).repeat(2, 1, 1)
    K[1, 0, 2] += 2.5
    h, w = y.shape[2:]
    rays = pixel_rays(K, h, w, 37, 50)
    K3 = change_intrinsic(K.double(), y.double(), x)
    ref = torch.bmm(torch.linalg.inv(K3), set_id_grid(y.double().permute(0, 2, 3, 1)).permute(0, 2, 1)).permute(0, 2, 1)
    assert rays.shape == (2, h * w, 3) and torch.allclose(rays.double(), ref, rtol=1e-5, atol=1e-6)


def test_rigid_warp_host_spelling_is_the_reference_composition():
    """rigid_warp on CPU tensors is warp_quat_xyz (src/modules/warp_utils.py:78-94) times check_valid; N = 1 is the pose
    composition t = R(q3) t_in + t3 of src/modellearn_proj_center.py:414-421."""
    from i2pnet_b200.modules import warp_utils as W
    from i2pnet_b200.projectPN.utils import check_valid
    torch.manual_seed(1)
    p, q, t = torch.randn(2, 9, 3), torch.randn(2, 4), torch.randn(2, 3)
    p[:, 2] = 0
    ref = W.warp_quat_xyz(p, q, torch.cat([torch.zeros(2, 1), t], -1))
    assert torch.equal(W.rigid_warp(p, q, t), ref)
    assert torch.equal(W.rigid_warp(p, q, t, mask_invalid=True), ref * check_valid(p))
    tq = torch.cat([torch.zeros(2, 1), p[:, 0]], 1).view(2, 1, 4)
    comp = (W.mul_q(W.mul_q(q, tq), W.inv_q(q)) + torch.cat([torch.zeros(2, 1), t], 1).view(2, 1, 4)).squeeze(1)[:, 1:]
    assert torch.allclose(W.rigid_warp(p[:, :1], q, t).view(2, 3), comp, atol=1e-6)


def test_step_scratch_is_plain_zeros_without_an_engine():
    from i2pnet_b200 import scratch
    scratch.begin_step("cpu")             # no arena on the host: a no-op
    z = scratch.zeros((2, 3), torch.float32, "cpu")
    scratch.end_step()
    assert z.shape == (2, 3) and z.dtype == torch.float32 and float(z.abs().sum()) == 0.0
