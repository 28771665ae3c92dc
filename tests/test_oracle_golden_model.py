"""CPU: pin oracle/model_cpu.py (the functional CPU restatement of the reference forward used by
bench.py's cpu_baseline / --impl reference) against vectors recorded from the real reference model."""
import numpy as np
import torch

from oracle import model_cpu
from tests.test_host_logic_cpu import _rel, load_golden_model


def test_model_cpu_matches_reference_python():
    g, state = load_golden_model()
    sd = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in state.items()}
    t = lambda k: torch.from_numpy(g[k])
    inter = {}
    out3, out4 = model_cpu.forward(sd, torch.from_numpy(g["rgb_u8"]).float(), t("lidar"), t("raw_point_xyz"),
                                   t("intrinsic"), t("lidar_feats"), intermediates=inter)
    inter["LiDAR_lv2"] = inter["LiDAR_lv2"][:, ::2, ::7]
    for name, val in inter.items():
        assert _rel(val.reshape(g["inter_" + name].shape), g["inter_" + name]) < 1e-4, name
    assert _rel(out3.detach(), g["out3"]) < 1e-4 and _rel(out4.detach(), g["out4"]) < 1e-4
    loss = model_cpu.loss_fn(out3, out4, t("q_gt"), t("t_gt"), sd["sx"], sd["sq"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for k in g.files:
        if k.startswith("grad__"):
            assert _rel(sd[k[len("grad__"):]].grad, g[k]) < 2e-3, k


def test_random_state_has_reference_layout():
    _, state = load_golden_model()
    sd = model_cpu.random_state(0)
    params = {k: v for k, v in state.items() if "running" not in k and "num_batches" not in k}
    assert set(sd) == set(params) and all(sd[k].shape == params[k].shape for k in sd)
