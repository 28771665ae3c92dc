"""Seeded operator cases shared by the parity tests and the golden generators.

Each case is a dict of numpy inputs plus an "op" tag.  Three runners evaluate a case and return
a dict of numpy outputs under the same keys:
    run_reference  the reference's own CUDA kernels (oracle/_ref/*.so, GPU box only)
    run_oracle     the C restatement (oracle/, CPU)
    run_product    the sm_100a kernels of libi2p_b200.so through i2pnet_b200._cabi (GPU)
EXACT lists, per op, the outputs that must agree bit for bit; the rest (atomic scatter-adds)
agree to rounding.
"""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXACT = {"fps": ["idx"], "ball_query": ["idx"], "three_nn": ["idx", "dist2"], "select": ["b", "h", "w", "mask"],
         "group": ["out", "gather_out", "interp_out"]}
CLOSE = {"group": ["grad", "gather_grad", "interp_grad"]}


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------
def _cloud(rng, b, n, dup=0.0, quant=None):
    """Points in a 80 x 80 x 4 m slab (SURVEY.md section 8d sweep shape); dup: fraction of exact
    duplicates; quant: snap coordinates to a grid so that many distances tie exactly."""
    xyz = np.stack([rng.uniform(-40, 40, (b, n)), rng.uniform(-40, 40, (b, n)), rng.uniform(-3, 1, (b, n))], -1)
    if quant:
        xyz = np.round(xyz / quant) * quant
    if dup > 0:
        src = rng.integers(0, n, size=(b, n))
        take = rng.uniform(size=(b, n)) < dup
        xyz = np.where(take[..., None], np.take_along_axis(xyz, src[..., None].repeat(3, -1), 1), xyz)
    return xyz.astype(np.float32)


def _range_image(rng, b, h, w, fill, scale=20.0):
    img = np.zeros((b, h, w, 3), np.float32)
    keep = rng.uniform(size=(b, h, w)) < fill
    img[keep] = ((rng.uniform(size=(int(keep.sum()), 3)) - 0.5) * scale).astype(np.float32)
    return img


def _grid(b, out_h, out_w, sh, sw):
    hh, ww = np.meshgrid(np.arange(out_h) * sh, np.arange(out_w) * sw, indexing="ij")
    return np.broadcast_to(np.stack([hh, ww], -1).reshape(1, -1, 2), (b, out_h * out_w, 2)).astype(np.int32).copy()


def all_cases():
    rng = np.random.Generator(np.random.PCG64(2024))
    c = {}
    # ---- furthest point sampling: every block-size class of the reference, ties, duplicates
    for name, (b, n, m, kw) in {
        "fps_n7": (2, 7, 5, {}), "fps_n100": (3, 100, 40, {}), "fps_n1000": (2, 1000, 300, {}),
        "fps_n1024": (2, 1024, 256, {}), "fps_n2500": (2, 2500, 700, {}), "fps_n8192": (2, 8192, 2048, {}),
        "fps_n20480": (1, 20480, 1024, {}), "fps_ties_grid": (2, 3000, 900, dict(quant=4.0)),
        "fps_dups": (2, 2048, 2048, dict(dup=0.5)), "fps_m_gt_distinct": (1, 300, 300, dict(quant=20.0)),
    }.items():
        c[name] = dict(op="fps", xyz=_cloud(rng, b, n, **kw), m=m)
    # ---- ball query / three_nn
    for name, (b, n, m, r, ns, kw) in {
        "ball_small": (2, 500, 77, 6.0, 16, {}), "ball_8k": (2, 8192, 2048, 2.0, 32, {}),
        "ball_none": (1, 300, 50, 0.01, 8, {}), "ball_ties": (2, 2000, 300, 4.0, 32, dict(quant=2.0)),
    }.items():
        xyz = _cloud(rng, b, n, **kw)
        q = xyz[:, rng.permutation(n)[:m]].copy() if name != "ball_none" else _cloud(rng, b, m) + 100.0
        c[name] = dict(op="ball_query", xyz=xyz, new_xyz=np.ascontiguousarray(q), radius=r, nsample=ns)
    for name, (b, n, m, kw) in {"nn3_small": (2, 333, 64, {}), "nn3_4k": (2, 4096, 1024, {}),
                                "nn3_ties": (2, 1500, 400, dict(quant=2.0)), "nn3_m2": (1, 50, 2, {})}.items():
        c[name] = dict(op="three_nn", unknown=_cloud(rng, b, n, **kw), known=_cloud(rng, b, m, **kw))
    # ---- group / gather / interpolate (+ grads)
    for name, (b, ch, n, p, s) in {"group_cv2": (2, 128, 80, 228, 32), "group_small": (3, 5, 17, 9, 4),
                                   "group_sa": (2, 35, 2048, 512, 16)}.items():
        w = rng.uniform(0.1, 1.0, (b, p, 3)).astype(np.float32)
        c[name] = dict(op="group", points=rng.standard_normal((b, ch, n)).astype(np.float32),
                       idx=rng.integers(0, n, (b, p, s)).astype(np.int32),
                       grad_out=rng.standard_normal((b, ch, p, s)).astype(np.float32),
                       idx3=rng.integers(0, n, (b, p, 3)).astype(np.int32), weight=(w / w.sum(-1, keepdims=True)),
                       grad3=rng.standard_normal((b, ch, p)).astype(np.float32))
    # ---- projection-window select: the nine call shapes of one forward, small batch
    def sel(b, H, W, oh, ow, sch, scw, kh, kw, K, flag, dist, sh=1, sw=1, fill=0.5, scale=20.0, small=None):
        x1 = _range_image(rng, b, H, W, fill, scale)
        x2 = x1 if small is None else np.ascontiguousarray(x1[:, ::sh, ::sw][:, :small[0], :small[1]])
        return dict(op="select", xyz1=x1, xyz2=x2, idx_n2=_grid(b, oh, ow, sch, scw),
                    random_hw=np.arange(kh * kw, dtype=np.int32), kH=kh, kW=kw, K=K, flag=flag, distance=dist,
                    stride_h=sh, stride_w=sw)
    c["select_sa1"] = sel(1, 64, 1800, 16, 225, 4, 8, 9, 15, 32, 3, 0.75, fill=0.3, scale=3.0)
    c["select_sa2"] = sel(2, 16, 225, 8, 113, 2, 2, 9, 15, 16, 3, 3.0, fill=0.7, scale=8.0)
    c["select_sa3"] = sel(2, 8, 113, 4, 57, 2, 2, 5, 9, 16, 3, 6.0, fill=0.8)
    c["select_sa4"] = sel(2, 4, 57, 4, 29, 1, 2, 5, 9, 16, 3, 12.0, fill=0.8)
    c["select_cv"] = sel(2, 4, 57, 4, 57, 1, 1, 3, 5, 4, 2, 4.5, fill=0.8, scale=10.0)
    c["select_up"] = sel(2, 4, 57, 4, 57, 1, 1, 5, 9, 8, 3, 9.0, sh=1, sw=2, fill=0.8, small=(4, 29))
    c["select_noflag"] = sel(2, 8, 113, 4, 57, 2, 2, 5, 9, 8, 0, 6.0, fill=0.6)
    c["select_copy_only"] = sel(2, 8, 113, 4, 57, 2, 2, 5, 9, 8, 1, 6.0, fill=0.6)
    c["select_sparse"] = sel(2, 16, 225, 8, 113, 2, 2, 9, 15, 16, 3, 3.0, fill=0.03)   # K > valid, empty centres
    q = sel(2, 16, 225, 8, 113, 2, 2, 9, 15, 16, 3, 1e3, fill=0.9)                      # exact-tie stress
    q["xyz1"] = np.round(q["xyz1"] / 5.0) * 5.0
    q["xyz2"] = q["xyz1"]
    c["select_ties"] = q
    perm = sel(2, 8, 113, 4, 57, 2, 2, 5, 9, 8, 3, 6.0, fill=0.7)
    perm["random_hw"] = rng.permutation(45).astype(np.int32)                            # shuffled visiting order
    c["select_randperm"] = perm
    # the reference's only fixture: the print-only __main__ of fused_conv_select_k.py:29-139
    x1 = np.ones((1, 4, 9, 3), np.float32)
    x2 = np.zeros((1, 4, 5, 3), np.float32)
    for r in range(4):
        x2[0, r, :, :] = np.array([1 + 4 * r, 2 + 4 * r, 3 + 4 * r, 4 + 4 * r, 1], np.float32)[:, None]
    for tag, flag in (("shift", 2), ("shift_copy", 3), ("none", 0)):
        c["select_main_" + tag] = dict(op="select", xyz1=x1, xyz2=x2, idx_n2=np.array([[[0, 2], [0, 0]]], np.int32),
                                       random_hw=np.arange(3, dtype=np.int32), kH=1, kW=3, K=5, flag=flag,
                                       distance=200.0, stride_h=1, stride_w=2)
    return c


def big_cases():
    """BASELINE configs[4] sizes (point-count sweep 32k ... 131k, M = N / 4): too slow for the CPU oracle and too big
    for a committed fixture, so they are checked LIVE against the reference's own kernels on the GPU box only."""
    rng = np.random.Generator(np.random.PCG64(4242))
    c = {}
    for n in (32768, 65536, 131072):
        tag = "%dk" % (n // 1024)
        m = n // 4
        xyz = _cloud(rng, 2, n)
        c["fps_" + tag] = dict(op="fps", xyz=xyz, m=m)
        c["fps_ties_" + tag] = dict(op="fps", xyz=_cloud(rng, 1, n, quant=0.5, dup=0.05), m=m // 4)
        q = np.ascontiguousarray(xyz[:, rng.permutation(n)[:m]])
        c["ball_" + tag] = dict(op="ball_query", xyz=xyz, new_xyz=q, radius=1.0, nsample=32)
        c["ball_ties_" + tag] = dict(op="ball_query", xyz=_cloud(rng, 1, n, quant=0.5), new_xyz=np.ascontiguousarray(q[:1, :m // 2]),
                                     radius=1.5, nsample=16)
        c["nn3_" + tag] = dict(op="three_nn", unknown=xyz, known=q)
        w = rng.uniform(0.1, 1.0, (2, m, 3)).astype(np.float32)
        c["group_" + tag] = dict(op="group", points=rng.standard_normal((2, 16, n)).astype(np.float32),
                                 idx=rng.integers(0, n, (2, m, 16)).astype(np.int32),
                                 grad_out=rng.standard_normal((2, 16, m, 16)).astype(np.float32),
                                 idx3=rng.integers(0, n, (2, m, 3)).astype(np.int32), weight=(w / w.sum(-1, keepdims=True)),
                                 grad3=rng.standard_normal((2, 16, m)).astype(np.float32))
    return c


# ----------------------------------------------------------------------------------------------
# runners
# ----------------------------------------------------------------------------------------------
def load_reference_extensions():
    """The reference's pybind modules compiled unmodified into oracle/_ref (see oracle/build_ref.py)."""
    mods = []
    for name in ("pointnet2_cuda", "fused_conv_select_k_cuda"):
        path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if not os.path.exists(path):
            return None, None
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _run_gpu(case, dev, K):
    """K: namespace with the nine pybind-style wrappers + fused_conv_select_k (reference or drop-in)."""
    op = case["op"]
    if op == "fps":
        xyz = _t(case["xyz"], dev)
        B, N, _ = xyz.shape
        temp = torch.full((B, N), 1e10, device=dev)
        idx = torch.zeros(B, case["m"], dtype=torch.int32, device=dev)
        K.furthest_point_sampling_wrapper(B, N, case["m"], xyz, temp, idx)
        return dict(idx=idx)
    if op == "ball_query":
        xyz, q = _t(case["xyz"], dev), _t(case["new_xyz"], dev)
        B, N, _ = xyz.shape
        M = q.shape[1]
        idx = torch.zeros(B, M, case["nsample"], dtype=torch.int32, device=dev)
        K.ball_query_wrapper(B, N, M, case["radius"], case["nsample"], q, xyz, idx)
        return dict(idx=idx)
    if op == "three_nn":
        u, k = _t(case["unknown"], dev), _t(case["known"], dev)
        B, N, _ = u.shape
        d2 = torch.zeros(B, N, 3, device=dev)
        idx = torch.zeros(B, N, 3, dtype=torch.int32, device=dev)
        K.three_nn_wrapper(B, N, k.shape[1], u, k, d2, idx)
        return dict(idx=idx, dist2=d2)
    if op == "group":
        pts, idx, go = _t(case["points"], dev), _t(case["idx"], dev), _t(case["grad_out"], dev)
        B, C, N = pts.shape
        _, P, S = idx.shape
        out = torch.zeros(B, C, P, S, device=dev)
        K.group_points_wrapper(B, C, N, P, S, pts, idx, out)
        grad = torch.zeros(B, C, N, device=dev)
        K.group_points_grad_wrapper(B, C, N, P, S, go, idx, grad)
        idx1 = idx[:, :, 0].contiguous()
        g_out = torch.zeros(B, C, P, device=dev)
        K.gather_points_wrapper(B, C, N, P, pts, idx1, g_out)
        g_grad = torch.zeros(B, C, N, device=dev)
        K.gather_points_grad_wrapper(B, C, N, P, _t(case["grad3"], dev), idx1, g_grad)
        idx3, w = _t(case["idx3"], dev), _t(case["weight"], dev)
        i_out = torch.zeros(B, C, P, device=dev)
        K.three_interpolate_wrapper(B, C, N, P, pts, idx3, w, i_out)
        i_grad = torch.zeros(B, C, N, device=dev)
        K.three_interpolate_grad_wrapper(B, C, P, N, _t(case["grad3"], dev), idx3, w, i_grad)
        return dict(out=out, grad=grad, gather_out=g_out, gather_grad=g_grad, interp_out=i_out, interp_grad=i_grad)
    if op == "select":
        x1, x2 = _t(case["xyz1"], dev), _t(case["xyz2"], dev)
        idx_n2, rhw = _t(case["idx_n2"], dev), _t(case["random_hw"], dev)
        B, H, W, _ = x1.shape
        n, Kk, tot = idx_n2.shape[1], case["K"], case["kH"] * case["kW"]
        sb, sh, sw = (torch.zeros(B, n, Kk, 1, dtype=torch.int64, device=dev) for _ in range(3))
        mask = torch.zeros(B, n, Kk, 1, device=dev)
        v1, v2 = torch.zeros(B, n, tot, 1, device=dev), torch.zeros(B, n, tot, 1, device=dev)
        K.fused_conv_select_k(x1, x2, idx_n2, rhw, H, W, n, case["kH"], case["kW"], Kk, case["flag"],
                              case["distance"], case["stride_h"], case["stride_w"], sb, sh, sw, v1, v2, mask,
                              x2.shape[1], x2.shape[2])
        assert float(v1.abs().sum()) == 0 and float(v2.abs().sum()) == 0   # never written
        return dict(b=sb.squeeze(-1).to(torch.int32), h=sh.squeeze(-1).to(torch.int32),
                    w=sw.squeeze(-1).to(torch.int32), mask=mask.squeeze(-1))
    raise KeyError(op)


class _NS:
    pass


def run_reference(case, pn2, fused, dev):
    ns = _NS()
    for name in dir(pn2):
        if name.endswith("_wrapper"):
            setattr(ns, name, getattr(pn2, name))
    ns.fused_conv_select_k = fused.fused_conv_select_k
    res = _run_gpu(case, dev, ns)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


def run_product(case, dev):
    """Through the drop-in modules, i.e. the exact call surface the reference's Python binds."""
    import sys
    if os.path.join(ROOT, "dropin") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "dropin"))
    import fused_conv_select_k_cuda as fused
    from pointnet2 import pointnet2_cuda as pn2
    return run_reference(case, pn2, fused, dev)


def run_oracle(case):
    from oracle import oracle as orc
    op = case["op"]
    if op == "fps":
        return dict(idx=orc.furthest_point_sample(case["xyz"], case["m"]))
    if op == "ball_query":
        return dict(idx=orc.ball_query(case["radius"], case["nsample"], case["xyz"], case["new_xyz"]))
    if op == "three_nn":
        d2, idx = orc.three_nn(case["unknown"], case["known"])
        return dict(idx=idx, dist2=d2)
    if op == "group":
        N = case["points"].shape[2]
        idx1 = np.ascontiguousarray(case["idx"][:, :, 0])
        return dict(out=orc.group_points(case["points"], case["idx"]),
                    grad=orc.group_points_grad(case["grad_out"], case["idx"], N),
                    gather_out=orc.gather_points(case["points"], idx1),
                    gather_grad=orc.gather_points_grad(case["grad3"], idx1, N),
                    interp_out=orc.three_interpolate(case["points"], case["idx3"], case["weight"]),
                    interp_grad=orc.three_interpolate_grad(case["grad3"], case["idx3"], case["weight"], N))
    if op == "select":
        b, h, w, m = orc.fused_conv_select_k(case["xyz1"], case["xyz2"], case["idx_n2"], case["random_hw"], case["kH"],
                                             case["kW"], case["K"], case["flag"], case["distance"], case["stride_h"],
                                             case["stride_w"])
        return dict(b=b.astype(np.int32), h=h.astype(np.int32), w=w.astype(np.int32), mask=m)
    raise KeyError(op)


def compare(case, got, want, who):
    op = case["op"]
    for k in EXACT.get(op, []):
        assert got[k].shape == want[k].shape, (who, k)
        if not np.array_equal(got[k], want[k]):
            bad = np.argwhere(got[k] != want[k])
            raise AssertionError("%s: %s differs at %d of %d positions, first %s: got %s want %s" % (
                who, k, len(bad), got[k].size, bad[0], got[k][tuple(bad[0])], want[k][tuple(bad[0])]))
    for k in CLOSE.get(op, []):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-4, atol=1e-4, err_msg="%s: %s" % (who, k))
