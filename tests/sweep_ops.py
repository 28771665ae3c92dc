"""Point-count sweep of the PointNet++ operator kernels (BASELINE.json configs[4]):
N in {4k ... 131k}, M = N / 4, B = 8, points uniform in a 80 x 80 x 4 m slab; furthest point sampling, ball query
(r = 0.5, nsample = 32), grouping (C = 64) forward / backward, 3-NN, kNN (k = 16).

    python tests/sweep_ops.py [--with-reference] [--out gpurun_out/sweep.md]

Every kernel is timed alone with CUDA events on torch's current stream (the stream the C ABI launches on), the L2
flushed before every launch, median of the repetitions.  Achieved GB/s = SURVEY.md section 8d's algorithmic bytes /
time, against the measured HBM peak of MEASURED_PEAKS.json.  --with-reference also times the reference's own CUDA
kernels (oracle/_ref, compiled unmodified for sm_100) on the same tensors: the "legacy kernel on B200" bar, a
measurement baseline only.
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # lives under tests/: it may time the reference kernels of oracle/_ref
sys.path.insert(0, ROOT)

from i2pnet_b200 import _cabi  # noqa: E402

B, C, NSAMPLE, RADIUS, KNN_K = 8, 64, 32, 0.5, 16


def event_ms(fn, flush, reps, skip=1):
    evs = []
    for i in range(reps + skip):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs[skip:])


def ops_for(K, dev, n, m, gen):
    """-> list of (name, launcher, algorithmic bytes, flops) over one set of tensors; K = namespace of wrappers with the
    reference's pybind signatures (i2pnet_b200._cabi or the reference module)."""
    xyz = torch.stack([torch.rand(B, n, device=dev, generator=gen) * 80 - 40, torch.rand(B, n, device=dev, generator=gen) * 80 - 40,
                       torch.rand(B, n, device=dev, generator=gen) * 4 - 3], -1).contiguous()
    temp = torch.empty(B, n, device=dev)
    fidx = torch.zeros(B, m, dtype=torch.int32, device=dev)

    def fps():
        temp.fill_(1e10)
        K.furthest_point_sampling_wrapper(B, n, m, xyz, temp, fidx)
    fps()
    new_xyz = torch.gather(xyz, 1, fidx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    bidx = torch.zeros(B, m, NSAMPLE, dtype=torch.int32, device=dev)
    ball = lambda: K.ball_query_wrapper(B, n, m, RADIUS, NSAMPLE, new_xyz, xyz, bidx)
    ball()
    feats = torch.randn(B, C, n, device=dev, generator=gen)
    gout = torch.empty(B, C, m, NSAMPLE, device=dev)
    grp = lambda: K.group_points_wrapper(B, C, n, m, NSAMPLE, feats, bidx, gout)
    ggrad = torch.randn(B, C, m, NSAMPLE, device=dev, generator=gen)
    gpts = torch.zeros(B, C, n, device=dev)
    grpb = lambda: K.group_points_grad_wrapper(B, C, n, m, NSAMPLE, ggrad, bidx, gpts)
    d3, i3 = torch.empty(B, n, 3, device=dev), torch.empty(B, n, 3, dtype=torch.int32, device=dev)
    nn3 = lambda: K.three_nn_wrapper(B, n, m, xyz, new_xyz, d3, i3)
    ops = [
        ("fps", fps, B * (12 * n + 4 * m + 8 * n), B * 9.0 * n * m),          # + the temp fill and read the ABI prescribes
        ("ball_query", ball, B * (12 * n + 12 * m + 4 * m * NSAMPLE), B * 8.0 * n * m),
        ("group", grp, B * (4 * C * min(n, m * NSAMPLE) + 4 * m * NSAMPLE + 4 * C * m * NSAMPLE), 0.0),
        ("group_grad", grpb, B * (4 * C * min(n, m * NSAMPLE) + 4 * m * NSAMPLE + 4 * C * m * NSAMPLE + 4 * C * n), 0.0),
        ("three_nn", nn3, B * (12 * n + 12 * m + 24 * n), B * 9.0 * n * m),
    ]
    if hasattr(K, "knn_wrapper"):
        kd, ki = torch.empty(B, m, KNN_K, device=dev), torch.empty(B, m, KNN_K, dtype=torch.int32, device=dev)
        ops.append(("knn", lambda: K.knn_wrapper(B, m, n, KNN_K, new_xyz, xyz, kd, ki), B * (12 * m + 12 * n + 8 * m * KNN_K),
                    B * 8.0 * n * m))
    return ops


class Ours:
    furthest_point_sampling_wrapper = staticmethod(_cabi.furthest_point_sampling)
    ball_query_wrapper = staticmethod(_cabi.ball_query)
    group_points_wrapper = staticmethod(_cabi.group_points)
    group_points_grad_wrapper = staticmethod(_cabi.group_points_grad)
    three_nn_wrapper = staticmethod(_cabi.three_nn)
    knn_wrapper = staticmethod(_cabi.knn)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--with-reference", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.md"))
    ap.add_argument("--sizes", default="4096,8192,16384,32768,65536,131072")
    ap.add_argument("--ops", default="", help="comma-separated subset of the operators (default: all)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    _cabi.lib()
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ref = None
    if args.with_reference:
        from tests.ref_cases import load_reference_extensions
        ref, _ = load_reference_extensions()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = ["# point-count sweep, B = 8, M = N / 4, C = 64, nsample = 32, r = 0.5, k = 16 (BASELINE.json configs[4])", "",
             "us per launch (CUDA events, L2 flushed, median); GB/s = algorithmic bytes (SURVEY.md 8d) / time; "
             "frac = of the measured %.0f GB/s HBM peak; GFLOP/s for the brute-force distance kernels; "
             "ref = the reference's CUDA kernel recompiled for sm_100, same tensors." % peak, "",
             "| op | N | M | us | GB/s | frac | GFLOP/s | ref us | speed-up |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for n in [int(s) for s in args.sizes.split(",")]:
        m = n // 4
        gen = torch.Generator(device=dev).manual_seed(n)
        ours = ops_for(Ours, dev, n, m, gen)
        refops = {}
        if ref is not None:
            gen = torch.Generator(device=dev).manual_seed(n)
            refops = {name: fn for name, fn, _, _ in ops_for(ref, dev, n, m, gen)}
        for name, fn, nbytes, flops in ours:
            if args.ops and name not in args.ops.split(","):
                continue
            reps = 3 if name == "fps" and n >= 32768 else 10
            ms = event_ms(fn, flush, reps)
            rms = event_ms(refops[name], flush, 2 if name == "fps" and n >= 32768 else 5) if name in refops else None
            lines.append("| %s | %d | %d | %.1f | %.1f | %.4f | %s | %s | %s |" % (
                name, n, m, ms * 1e3, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak,
                "%.0f" % (flops / ms / 1e6) if flops else "",
                "%.1f" % (rms * 1e3) if rms is not None else "", "%.1fx" % (rms / ms) if rms is not None else ""))
            print(lines[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
