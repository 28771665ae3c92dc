"""Isolated timing of the cost-volume glue kernels (csrc/cv.cu) at the batch-8 KITTI shapes, L2 flushed, CUDA events.
`python tools/bench_cv.py profile` launches each once inside a profiler range (for ncu --profile-from-start off)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2pnet_b200.projectPN.PPBackbone_center import _CvBuild, _SoftmaxWSum  # noqa: E402

dev = torch.device("cuda:0")
PROFILE = len(sys.argv) > 1 and sys.argv[1] == "profile"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)


def timed(fn, reps=12, skip=2):
    evs = []
    for i in range(reps):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs[skip:]) * 1e3


for name, (B, N, K, N2, C, has_max, use_idx) in {"cv1": (8, 228, 80, 80, 128, True, False), "cv2": (8, 228, 32, 80, 128, False, True)}.items():
    xyz1, xyz2 = r(B, N, 3), r(B, N2, 3)
    pi, qi = r(B, N, C).requires_grad_(True), r(B, N2, C).requires_grad_(True)
    maxc = r(B, N2, C).requires_grad_(True) if has_max else None
    idx = torch.randint(0, N2, (B, N, K), device=dev, generator=g, dtype=torch.int32) if use_idx else None
    X, xyz6 = _CvBuild.apply(xyz1, xyz2, pi, qi, maxc, idx)
    dX, d6 = torch.randn_like(X), torch.randn_like(xyz6)
    fwd = lambda: _CvBuild.apply(xyz1, xyz2, pi, qi, maxc, idx)
    bwd = lambda: torch.autograd.grad([X, xyz6], [pi, qi] + ([maxc] if has_max else []), [dX, d6], retain_graph=True)
    if PROFILE:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fwd(); bwd()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    # the backward through the C ABI alone (autograd's host overhead exceeds the kernel's duration)
    from i2pnet_b200 import _cabi
    outs = [torch.zeros(B, N, 3, device=dev), torch.zeros(B, N2, 3, device=dev), torch.zeros(B, N, C, device=dev),
            torch.zeros(B, N2, C, device=dev), torch.zeros(B, N2, C, device=dev)]
    bwd = lambda: _cabi.call("i2p_cv_build_bwd", dev, B, N, K, N2, C, int(has_max), dX.data_ptr(), d6.data_ptr(), pi.data_ptr(),
                             qi.data_ptr(), idx.data_ptr() if use_idx else None, outs[0].data_ptr(), outs[1].data_ptr(),
                             outs[2].data_ptr(), outs[3].data_ptr(), outs[4].data_ptr() if has_max else None)
    tf, tb = timed(fwd), timed(bwd)
    mb = X.numel() * 4 / 1e6
    print("%s: X %.0f MB | build %.1f us (%.0f GB/s) | build_bwd %.1f us (%.0f GB/s)" % (name, mb, tf, mb / tf * 1e3, tb, mb / tb * 1e3))
