"""nuScenes-shape configuration (BASELINE.json configs[3]: 320x640 image, 21x1800 range image ->
11x225 / 6x113 / 3x57 / 3x29 centres, 10x20 level-3 pixels): RegNet_v2 with I2PNetConfigNus against the reference's
own model under src/config_proj_lidarcenter_nus.py (tests/golden/make_golden.py nus)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN
from tests.test_host_logic_cpu import check_against_golden

HOOKS = ("LiDAR_lv2", "LiDAR_lv3", "cost_volume1", "layer_idx", "set_upconv0_upsample", "cost_volume2")


def _load():
    g = np.load(os.path.join(GOLDEN, "ref_model_nus_b2.npz"))
    state = {k[len("state__"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state__")}
    return g, state


def _run(device):
    from i2pnet_b200.compute_loss import Get_loss
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfigNus as cfg
    from i2pnet_b200.modellearn_proj_center import RegNet_v2
    g, state = _load()
    model = RegNet_v2(cfg=cfg)
    res = model.load_state_dict(state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.train()
    for head in (model.l4_head, model.l3_head):
        head.DP1.p = 0.0
    model.to(device)
    assert (model.lidar_Hs, model.lidar_Ws) == ([11, 6, 3, 3], [225, 113, 57, 29])
    t = lambda k: torch.from_numpy(g[k]).to(device)
    inter = {}
    for name in HOOKS:
        def hook(mod, args, out, name=name):
            inter[name] = (out[2] if isinstance(out, tuple) else out).detach()
        getattr(model, name).register_forward_hook(hook)
    out3, out4, _, _, sx, sq = model(torch.from_numpy(g["rgb_u8"]).float().to(device), t("lidar"), t("raw_point_xyz"), None,
                                     t("intrinsic"), None, None, None, t("lidar_feats"), cfg)
    loss, _, _ = Get_loss(out3, out4, t("q_gt"), t("t_gt"), sx, sq, cfg)
    loss.backward()
    return model, g, out3, out4, loss, inter


def test_nuscenes_shape_host_logic_matches_reference(oracle_backend):
    check_against_golden(*_run("cpu"), grad_tol=5e-3)


@pytest.mark.gpu
def test_nuscenes_shape_model_matches_reference_on_gpu():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    check_against_golden(*_run("cuda:0"), grad_tol=2e-2)      # gradients vs a CPU recording: see tests/test_model_gpu.py
