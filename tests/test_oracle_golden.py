"""CPU: pin the C oracle.
(1) against tests/golden/ref_gpu_golden.npz -- outputs of the reference's own CUDA kernels
    (oracle/_ref, compiled unmodified) recorded on a B200 by tests/golden/make_ref_gpu_golden.py;
(2) against vectors recorded from the reference's own Python (tests/golden/make_golden.py);
(3) against the hand-derived answers for the reference's only fixture, the print-only __main__ of
    fused_conv_select_k.py:29-139."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import ref_cases
from tests.conftest import GOLDEN

CASES = ref_cases.all_cases()
REF_GPU = os.path.join(GOLDEN, "ref_gpu_golden.npz")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_cuda_kernels(name):
    if not os.path.exists(REF_GPU):
        pytest.skip("ref_gpu_golden.npz not recorded yet (needs one gpurun of make_ref_gpu_golden.py)")
    g = np.load(REF_GPU)
    want = {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(name + "__")}
    assert want, name
    ref_cases.compare(CASES[name], ref_cases.run_oracle(CASES[name]), want, "oracle vs reference CUDA kernel")


def test_select_known_answers_of_reference_main():
    h = lambda tag: ref_cases.run_oracle(CASES["select_main_" + tag])
    r = h("shift")
    assert (r["h"] == 0).all()
    assert r["w"].tolist() == [[[0, 1, 2, 0, 0], [4, 0, 1, 0, 0]]]
    assert r["mask"].tolist() == [[[1, 1, 1, 0, 0], [1, 1, 1, 0, 0]]]
    r = h("shift_copy")
    assert r["w"].tolist() == [[[0, 1, 2, 0, 0], [4, 0, 1, 4, 4]]]
    assert (r["mask"] == 1).all()
    r = h("none")
    assert r["w"].tolist() == [[[0, 1, 2, 0, 0], [0, 1, 0, 0, 0]]]
    assert r["mask"].tolist() == [[[1, 1, 1, 0, 0], [1, 1, 0, 0, 0]]]


def test_fps_tie_rule_is_bit_reversed_thread_id():
    """Equal distances: the reference's tree reduction keeps the lower slot at every level, so the
    winner is the thread (k mod bs) with the smallest bit-reversed id, not the lowest index."""
    n = 1024
    xyz = np.zeros((1, n, 3), np.float32)
    xyz[0, 3] = xyz[0, 600] = (5, 0, 0)        # two equally far candidates in threads 3 and 600
    idx = orc.furthest_point_sample(xyz, 2)
    rev = lambda t: int(format(t, "010b")[::-1], 2)
    assert rev(600) < rev(3) and idx[0, 1] == 600


def test_project_seq_matches_reference_python():
    g = np.load(os.path.join(GOLDEN, "ref_project_seq.npz"))
    xyz_proj, (f1, f2) = orc.project_seq(g["raw"], [g["feats"], g["cam"]], 64, 1800, 2.0, -24.8)
    nz = np.abs(xyz_proj).sum(-1) > 0
    assert np.array_equal(np.argwhere(nz), g["cells"])
    assert np.array_equal(xyz_proj[nz], g["xyz"]) and np.array_equal(f1[nz], g["f1"]) and np.array_equal(f2[nz], g["f2"])


def test_select_gather_knn_match_reference_python():
    g = np.load(os.path.join(GOLDEN, "ref_select_gather.npz"))
    img, idx_n2, feat = g["img"], g["idx_n2"], g["feat"]
    B, H, W, _ = img.shape
    for tag, flag in (("copy", 3), ("att", 2)):
        _, h, w, m = orc.fused_conv_select_k(img, img, idx_n2, np.arange(45, dtype=np.int32), 5, 9, 8, flag, 6.0)
        assert np.array_equal(h, g[tag + "_h"]) and np.array_equal(w, g[tag + "_w"])
        assert np.array_equal(m[..., None], g[tag + "_mask"])
        rows = orc.gather_rows(feat.reshape(B, H * W, -1), (h * W + w).reshape(B, -1).astype(np.int32))
        assert np.array_equal(rows.reshape(g[tag + "_gather"].shape), g[tag + "_gather"])
    idx = orc.knn(16, g["knn_s"], g["knn_q"])
    assert np.array_equal(np.sort(idx, -1), g["knn_idx_sorted"])     # set equality per query
