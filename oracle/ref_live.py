"""The UNCHANGED reference model, live: model-level oracle and "legacy kernels on B200" baseline.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, and executed by bench.py's
`legacy_b200` leg as a SUBPROCESS after the timed region.  Nothing in i2pnet_b200/ imports it.

What runs is the reference's own Python -- src/modellearn_proj_center.py:24-424 (RegNet_v2),
compute_loss.py:102-133 (Get_loss), src/projectPN/*, src/modules/*, pointnet2/pointnet2_utils.py -- byte for byte,
staged by oracle/build_ref.py under oracle/_ref/py/ (or read from /root/reference where that exists), on one of
three operator back ends for its two pybind modules `pointnet2.pointnet2_cuda` / `fused_conv_select_k_cuda`:

  "ref"     oracle/_ref/*.so: the reference's own CUDA kernels compiled unmodified for sm_100 -> the live model-level
            oracle on the GPU box and the step-level legacy baseline;
  "dropin"  <repo>/dropin: the same two module names served by libi2p_b200.so -> north_star's "drop in unchanged";
  "cpu"     oracle/ref_shims.py: the C oracle on CPU tensors (build container, no GPU).

    ns = load("ref")                       # ns.RegNet_v2, ns.RegNet_v2_iter, ns.Get_loss, ns.cfg, ns.cfg_nus
    python -m oracle.ref_live --bench --batch 8 --steps 10      # one JSON line: the trainer's iteration, timed
"""
import argparse
import importlib
import importlib.util
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, "_ref", "py")
_PURGE = ("src", "pointnet2", "pointnet_util", "compute_loss", "fused_conv_select_k_cuda")


def python_root():
    """Where the reference's Python is read from: the staged copy (GPU box), else the checkout (build container)."""
    if os.path.isfile(os.path.join(STAGED, "src", "modellearn_proj_center.py")):
        return STAGED
    ref = os.environ.get("I2P_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(ref, "src", "modellearn_proj_center.py")):
        return ref
    return None


def available(backend="ref"):
    if python_root() is None:
        return False
    if backend == "ref":
        return all(os.path.exists(os.path.join(HERE, "_ref", n + ".so")) for n in ("pointnet2_cuda", "fused_conv_select_k_cuda"))
    return True


def _purge():
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in _PURGE):
            del sys.modules[k]
    importlib.invalidate_caches()


def _load_so(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "_ref", name + ".so"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(backend="ref"):
    """Import the reference's model files on `backend`; -> namespace.  Each call imports FRESH module objects (the
    reference binds its extension modules at import time), so "ref" and "dropin" instances can live side by side."""
    import torch  # noqa: F401  (the .so files link against libtorch)
    root = python_root()
    if root is None:
        raise RuntimeError("reference Python not found: run oracle/build_ref.py where /root/reference exists")
    _purge()
    added = [root]
    for name in ("matplotlib", "matplotlib.pyplot", "cv2"):        # imported, unused on this path (utils.py:6, src/utils.py:2)
        sys.modules.setdefault(name, types.ModuleType(name))
    if backend == "cpu":
        from . import ref_shims
        ref_shims.install(root)
    elif backend == "ref":
        sys.path.insert(0, root)
        pn2, fused = _load_so("pointnet2_cuda"), _load_so("fused_conv_select_k_cuda")
        sys.modules["pointnet2.pointnet2_cuda"] = pn2
        sys.modules["fused_conv_select_k_cuda"] = fused
        import pointnet2
        pointnet2.pointnet2_cuda = pn2
    elif backend == "dropin":
        dropin = os.path.join(ROOT, "dropin")
        sys.path.insert(0, root)
        sys.path.insert(0, dropin)      # ahead of the reference's own (unbuilt) pointnet2/ directory
        added.append(dropin)
    else:
        raise ValueError(backend)
    try:
        ns = types.SimpleNamespace(backend=backend, root=root)
        ns.cfg = importlib.import_module("src.config_proj_lidarcenter").I2PNetConfig
        ns.cfg_nus = importlib.import_module("src.config_proj_lidarcenter_nus").I2PNetConfig
        for c in (ns.cfg, ns.cfg_nus):
            c.efgh = False          # read by the trainer (train20v2learn_wandb_proj.py:453), defined by no shipped config
        ns.RegNet_v2 = importlib.import_module("src.modellearn_proj_center").RegNet_v2
        ns.RegNet_v2_iter = importlib.import_module("src.modellearn_proj_center_iter").RegNet_v2
        ns.Get_loss = importlib.import_module("compute_loss").Get_loss
        ns.pointnet2_utils = importlib.import_module("pointnet2.pointnet2_utils")
        ns.pointnet2_modules = importlib.import_module("pointnet2.pointnet2_modules")
        ns.extension = sys.modules["pointnet2.pointnet2_cuda"]
        ns.fused_extension = sys.modules["fused_conv_select_k_cuda"]
    finally:
        for p in added:
            while p in sys.path:
                sys.path.remove(p)
        _purge()
    return ns


def forward_loss(ns, model, batch, cfg=None):
    """One forward + loss exactly as the trainer calls them (train20v2learn_wandb_proj.py:457-465)."""
    cfg = cfg or ns.cfg
    out3, out4, _, _, sx, sq = model(batch["rgb"], batch["lidar"], batch["raw_point_xyz"], None, batch["intrinsic"], None,
                                     None, None, batch["lidar_feats"], cfg=cfg)
    loss, _, _ = ns.Get_loss(out3, out4, batch["q_gt"], batch["t_gt"], sx, sq, cfg=cfg)
    return out3, out4, loss


def train_iteration(ns, model, opt, batch, clip=10.0, cfg=None):
    """The body of the trainer's loop (train20v2learn_wandb_proj.py:457-483): forward, zero_grad, loss, loss.item(),
    backward, clip_grad_norm_, Adam.step.  -> (out3, loss value)"""
    import torch
    out3, out4, loss = forward_loss(ns, model, batch, cfg)
    opt.zero_grad()
    value = loss.item()
    loss.backward()
    if clip > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    opt.step()
    return out3, value


def make_optimizer(model, lr=1e-3):
    import torch
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)   # :198-202


def bench(backend, batch, steps, warmup, n_points, image_hw, nus=False, cudnn=False, seed=0):
    """pairs/s of the unchanged reference trainer iteration on cuda:0 -> dict.  Device-resident inputs for `value`;
    `e2e` adds the trainer's own per-iteration `.to(device)` of the host batch (:438-449, from pinned memory here) and its
    `out3` read-back (cal_rete_once, :485)."""
    import torch
    from i2pnet_b200.synthetic import make_pairs   # input generator only (no operators)
    ns = load(backend)
    cfg = ns.cfg_nus if nus else ns.cfg
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.enabled = bool(cudnn)       # the trainer runs set_seed(0): cuDNN off (src/deterministic.py:36-38)
    torch.manual_seed(seed)
    model = ns.RegNet_v2(cfg=cfg).to(dev)
    model.train()
    opt = make_optimizer(model)
    kw = dict(init_H=cfg.init_H, init_W=cfg.init_W, fup=cfg.fup, fdown=cfg.fdown) if nus else {}
    host = [{k: v.pin_memory() for k, v in make_pairs(batch, n_points, image_hw, seed=100 + i, **kw).items()} for i in range(2)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(n, from_host):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            flush.fill_(i & 0xff)
            if from_host:
                data = {k: v.to(dev, non_blocking=True) for k, v in host[i & 1].items()}
                out3, _ = train_iteration(ns, model, opt, data, cfg=cfg)
                out3.detach().cpu()
            else:
                train_iteration(ns, model, opt, devb[i & 1], cfg=cfg)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    run(max(warmup, 1), False)
    ms = run(steps, False)
    ms_e2e = run(steps, True)
    return {"backend": backend, "cudnn": bool(cudnn), "batch": batch, "steps": steps,
            "pairs_per_s": batch * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "e2e_pairs_per_s": batch * steps / (ms_e2e * 1e-3), "e2e_ms_per_step": ms_e2e / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--backend", default="ref", choices=["ref", "dropin"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=20480)
    ap.add_argument("--image", type=int, nargs=2, default=[160, 512])
    ap.add_argument("--nus", action="store_true")
    ap.add_argument("--cudnn", action="store_true")
    args = ap.parse_args()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if args.bench:
        t0 = time.time()
        out = bench(args.backend, args.batch, args.steps, args.warmup, args.points, tuple(args.image), args.nus, args.cudnn)
        out["wall_s"] = time.time() - t0
        print(json.dumps(out))


if __name__ == "__main__":
    main()
