"""One profiled training step for ncu (use with --profile-from-start off): a warm eager step
outside the capture range, then one eager step (or the select kernel alone) inside it.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py step
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from i2pnet_b200.engine import TrainStep  # noqa: E402
from i2pnet_b200.synthetic import make_pairs  # noqa: E402


def mlp_micro(dev):
    """The cost-volume-1 shared MLP (rows = 8 x 228 x 80, 262 -> 128 -> 64 -> 64) forward + backward and
    the SA1 select, alone: a small memory footprint keeps ncu's save/restore between replays cheap."""
    from i2pnet_b200.projectPN import PPBackbone_center as P
    from i2pnet_b200.projectPN.utils import FLAG_COPY, FLAG_SHIFT, StrideGrid, project_seq, select_flat
    torch.manual_seed(0)
    mods, c = [], 262
    for co in (128, 64, 64):
        mods.append(P.Conv2d(c, co, [1, 1], bn=True).to(dev))
        c = co
    x = torch.randn(8, 228, 80, 262, device=dev, requires_grad=True)
    d = make_pairs(8, seed=7)
    _, (cam,) = project_seq(d["raw_point_xyz"].to(dev), [d["lidar"].to(dev)], 64, 1800, False)
    grid = StrideGrid(8, 16, 225, 4, 8, dev)

    def body():
        out = P.run_mlp(mods, x)
        out.backward(torch.ones_like(out))
        select_flat(cam, cam, grid, [9, 15], 32, FLAG_SHIFT | FLAG_COPY, 0.75)
    body()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    body()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def marked_step(eng, out_json):
    """One eager step with a marker kernel (torch.cuda._sleep -> `spin_kernel`) launched at every top-level module
    boundary, forward and backward; the labels, in launch order, go to `out_json`.  tools/summarise_ncu.py zips them
    with the spin_kernel launches of the ncu launch list to attribute device time to modules."""
    import json
    labels = []

    def mark(label):
        labels.append(label)
        torch.cuda._sleep(1)

    handles = []
    for name, m in eng.model.named_children():
        handles.append(m.register_forward_pre_hook(lambda mod, inp, n=name: mark("fwd " + n)))
        handles.append(m.register_forward_hook(lambda mod, inp, out, n=name: mark("fwd (assembly)")))
        handles.append(m.register_full_backward_pre_hook(lambda mod, g, n=name: mark("bwd " + n)))
        handles.append(m.register_full_backward_hook(lambda mod, gi, go, n=name: mark("bwd (assembly)")))
    mark("step prologue")
    x = eng.inputs
    eng.bucket.release()
    out3, out4, _, _, sx, sq = eng.model(x["rgb"], x["lidar"], x["raw_point_xyz"], None, x["intrinsic"], None, None, None,
                                         x["lidar_feats"], eng.cfg)
    mark("loss")
    from i2pnet_b200.compute_loss import Get_loss
    loss, _, _ = Get_loss(out3, out4, x["q_gt"], x["t_gt"], sx, sq, eng.cfg)
    mark("bwd (assembly)")
    loss.backward()
    mark("gather + clip + adam")
    eng.bucket.gather()
    eng.bucket.all_reduce_mean(eng.group)
    eng.bucket.clip_(eng.clip)
    eng.opt.step()
    mark("end")
    for h in handles:
        h.remove()
    with open(out_json, "w") as fh:
        json.dump(labels, fh)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "step"
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if what == "mlp":
        return mlp_micro(dev)
    eng = TrainStep(8, device=dev, use_graph=False)
    eng.load({k: v.to(dev) for k, v in make_pairs(8, seed=0).items()})
    for _ in range(2):
        eng._step_body()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if what == "step":
        eng._step_body()
    elif what == "kineto":
        # real (warm, unserialised) kernel durations through CUPTI: sum per kernel name over 3 eager steps
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                eng._step_body()
            torch.cuda.synchronize()
        rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
        rows.sort(key=lambda r: -r[2])
        total = sum(r[2] for r in rows)
        out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/kineto_step.md"
        with open(out, "w") as fh:
            fh.write("# CUPTI kernel durations, eager training step (3 steps; per-step figures below)\n\n")
            fh.write("%.1f us of kernel time per step, %d launches per step\n\n" % (total / 3, sum(r[1] for r in rows) / 3))
            fh.write("| share | us / step | launches / step | kernel |\n|---:|---:|---:|---|\n")
            for k, n, t in rows[:80]:
                fh.write("| %.1f %% | %.1f | %.1f | `%s` |\n" % (100 * t / total, t / 3, n / 3, k[:120]))
        print("wrote", out)
    elif what == "kineto_graph":
        # the captured step as it is replayed (side streams overlapping): per-stream busy time, span, top kernels
        from torch.profiler import ProfilerActivity, profile
        eng2 = TrainStep(8, device=dev, use_graph=True)
        eng2.load({k: v.to(dev) for k, v in make_pairs(8, seed=0).items()})
        eng2.warmup_and_capture(eager_steps=3)
        for _ in range(3):
            eng2.step()
        torch.cuda.synchronize()
        reps = 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                eng2.step()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 0]
        streams, names = {}, {}
        for e in evs:
            tr = e.time_range
            st = getattr(e, "device_resource_id", None)     # the CUDA stream id of a kernel record
            rec = streams.setdefault(st, [0.0, 0, None, None])
            rec[0] += e.time_range.elapsed_us()
            rec[1] += 1
            rec[2] = tr.start if rec[2] is None else min(rec[2], tr.start)
            rec[3] = tr.end if rec[3] is None else max(rec[3], tr.end)
            n = names.setdefault(e.name, [0.0, 0])
            n[0] += e.time_range.elapsed_us()
            n[1] += 1
        t0 = min(r[2] for r in streams.values())
        t1 = max(r[3] for r in streams.values())
        out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/kineto_graph.md"
        with open(out, "w") as fh:
            fh.write("# CUPTI kernel records of the replayed CUDA-graph training step (%d replays; per-step figures)\n\n" % reps)
            fh.write("span of all kernels: %.1f us per step; sum of kernel durations: %.1f us per step\n\n" % (
                (t1 - t0) / reps, sum(r[0] for r in streams.values()) / reps))
            fh.write("| stream | busy us / step | kernels / step |\n|---|---:|---:|\n")
            for st, r in sorted(streams.items(), key=lambda kv: -kv[1][0]):
                fh.write("| %s | %.1f | %.1f |\n" % (st, r[0] / reps, r[1] / reps))
            fh.write("\n| share | us / step | launches / step | kernel |\n|---:|---:|---:|---|\n")
            total = sum(v[0] for v in names.values())
            for k, v in sorted(names.items(), key=lambda kv: -kv[1][0])[:60]:
                fh.write("| %.1f %% | %.1f | %.1f | `%s` |\n" % (100 * v[0] / total, v[0] / reps, v[1] / reps, k[:120]))
        print("wrote", out)
    elif what == "timeline":
        # one replay of the captured step as a time-ordered kernel list (start offset, duration, stream, name): the input
        # of the critical-path analysis (which stream is busy when, where the step waits)
        from torch.profiler import ProfilerActivity, profile
        batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
        eng2 = TrainStep(batch, device=dev, use_graph=True)
        eng2.load({k: v.to(dev) for k, v in make_pairs(batch, seed=0).items()})
        eng2.warmup_and_capture(eager_steps=3)
        for _ in range(3):
            eng2.step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng2.step()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 0]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/timeline.csv"
        with open(out, "w") as fh:
            fh.write("start_us,dur_us,stream,name\n")
            for e in evs:
                fh.write("%.2f,%.2f,%s,\"%s\"\n" % (e.time_range.start - t0, e.time_range.elapsed_us(),
                                                   getattr(e, "device_resource_id", None), e.name[:100].replace('"', "'")))
        print("wrote", out, len(evs), "kernels, span %.1f us" % (evs[-1].time_range.end - t0))
    elif what == "marked":
        marked_step(eng, sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/markers.json")
    else:  # forward only
        with torch.no_grad():
            x = eng.inputs
            eng.model(x["rgb"], x["lidar"], x["raw_point_xyz"], None, x["intrinsic"], None, None, None,
                      x["lidar_feats"], eng.cfg)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
