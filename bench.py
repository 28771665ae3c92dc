#!/usr/bin/env python
"""Benchmark of the I2PNet hot path: image + point-cloud pairs/s of one full RegNet_v2 training
step (forward + loss + backward + grad all-reduce + clip + Adam) on KITTI-shaped synthetic pairs
(160x512 RGB + 20480 points, batch 8 per GPU; BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # ours, one JSON line
    python bench.py --impl reference [--steps K] [--warmup W]        # the reference path on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N     # one rank per GPU, weak scaling

Keys follow the driver's contract: `value` = whole-job pairs/s with inputs resident in HBM,
`e2e` = the same through engine.HostPipeline (every batch from pinned host memory, copied while the
previous step runs; every loss back on the host), `roofline` = the dominant own kernel (the tensor-core dX GEMM, at the largest shared-MLP layer; the other hot
kernels follow in `roofline_others`) against the measured HBM peak, `cpu_baseline` = the
CPU oracle port (oracle/model_cpu.py) on a bounded sample.  The oracle is executed only in the
cpu_baseline leg and the --impl reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs: [1] = the default bench line; [2]'s shape (batch 32) in f32; [3] nuScenes shape, batch 16
CONFIGS = {
    "kitti_b8": dict(workload="kitti_160x512_rgb+20480pts_batch8_per_gpu_fwd+bwd+adam", points=20480, image=(160, 512),
                     batch=8, nus=False),
    "kitti_b32": dict(workload="kitti_160x512_rgb+20480pts_batch32_per_gpu_fwd+bwd+adam", points=20480, image=(160, 512),
                      batch=32, nus=False),
    "nus_b16": dict(workload="nuscenes_320x640_rgb+40960pts_batch16_per_gpu_fwd+bwd+adam", points=40960, image=(320, 640),
                    batch=16, nus=True),
}
WORKLOAD = CONFIGS["kitti_b8"]["workload"]
N_POINTS, IMAGE_HW, BATCH = 20480, (160, 512), 8


def _cfg_of(conf):
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig, I2PNetConfigNus
    return I2PNetConfigNus if conf["nus"] else I2PNetConfig


def _pairs(conf, batch, seed):
    from i2pnet_b200.synthetic import make_pairs
    if conf["nus"]:
        cfg = _cfg_of(conf)
        return make_pairs(batch, conf["points"], conf["image"], init_H=cfg.init_H, init_W=cfg.init_W, fup=cfg.fup,
                          fdown=cfg.fdown, seed=seed)
    return make_pairs(batch, conf["points"], conf["image"], seed=seed)


def _config_dict(conf, name, per_gpu_batch, world, strong):
    """The `config` object both arms print (same keys and values, so that the driver sees the same configuration)."""
    return {"workload": conf["workload"] if not strong else conf["workload"].replace("_per_gpu", "_global"),
            "name": name, "points": conf["points"], "image": list(conf["image"]), "per_gpu_batch": per_gpu_batch,
            "global_batch": per_gpu_batch * world, "parallelism": "dp%d" % world,
            "l2": "256 MB flush write between steps (GPU arm)"}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_steps(steps, warmup, batch, threads=None, seed=0, conf=None):
    """The reference path on host cores: oracle/model_cpu.py forward + loss + backward + Adam on
    `batch` pairs per step.  -> (seconds per step list, cores)"""
    import torch
    from oracle import model_cpu
    conf = conf or CONFIGS["kitti_b8"]
    if conf["nus"]:
        raise SystemExit("bench.py: the CPU port (oracle/model_cpu.py) covers the KITTI configuration only")
    # every host thread this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the reference arm into a single-thread run
    if not threads:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    sd = {k: v.requires_grad_(True) for k, v in model_cpu.random_state(seed).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-3, weight_decay=1e-4)
    times = []
    for i in range(warmup + steps):
        d = _pairs(conf, batch, 1000 + i)
        t0 = time.perf_counter()
        opt.zero_grad()
        out3, out4 = model_cpu.forward(sd, d["rgb"], d["lidar"], d["raw_point_xyz"], d["intrinsic"], d["lidar_feats"])
        loss = model_cpu.loss_fn(out3, out4, d["q_gt"], d["t_gt"], sd["sx"], sd["sq"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 10.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    conf = CONFIGS[args.config]
    per_gpu = _per_gpu_batch(conf, world, args.strong)
    # a bounded sample of the workload: batch <= 8 pairs per step (~1 s on 16 cores), the metric is per pair
    batch = min(per_gpu, 8)
    times, cores = cpu_reference_steps(args.steps, args.warmup, batch, conf=conf)
    total = sum(times)
    value = batch * len(times) / total
    sample = "%d steps of batch %d (of the batch-%d workload) on %d host threads, oracle/model_cpu.py" % (
        len(times), batch, per_gpu, cores)
    print(json.dumps({
        "impl": "reference", "metric": "pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config_dict(conf, args.config, per_gpu, world, args.strong),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _per_gpu_batch(conf, world, strong):
    if not strong:
        return conf["batch"]
    if conf["batch"] % world:
        raise SystemExit("bench.py --strong: global batch %d does not split over %d GPUs" % (conf["batch"], world))
    return conf["batch"] // world


def legacy_b200(conf, batch, steps=10):
    """The reference's own kernels recompiled for sm_100 under its UNCHANGED model / loss / optimiser loop on this GPU
    (oracle/ref_live.py on oracle/_ref, as a subprocess after the timed region): BASELINE.md B2(i), the comparator a
    GPU kernel rewrite has to beat.  cuDNN off is what the trainer runs (src/deterministic.py:36-38); on is kinder."""
    out = {}
    for tag, flag in (("cudnn_off", []), ("cudnn_on", ["--cudnn"])):
        cmd = [sys.executable, "-m", "oracle.ref_live", "--bench", "--batch", str(batch), "--steps", str(steps), "--warmup", "3",
               "--points", str(conf["points"]), "--image", str(conf["image"][0]), str(conf["image"][1])] + flag + \
              (["--nus"] if conf["nus"] else [])
        try:
            res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
            line = [l for l in res.stdout.splitlines() if l.startswith("{")]
            if res.returncode != 0 or not line:
                out[tag] = {"unavailable": (res.stderr.strip().splitlines() or ["no output"])[-1][:200]}
            else:
                r = json.loads(line[-1])
                out[tag] = {k: r[k] for k in ("pairs_per_s", "ms_per_step", "e2e_pairs_per_s", "e2e_ms_per_step")}
        except Exception as e:   # noqa: BLE001
            out[tag] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:160])}
    out["what"] = ("unchanged src/modellearn_proj_center.py + compute_loss.py + clip_grad_norm_ + Adam, eager, on the "
                   "reference's CUDA kernels compiled for sm_100 (oracle/_ref), batch %d, %d steps" % (batch, steps))
    return out


def _event_time(fn, flush, reps=25, skip=5):
    """Median CUDA-event duration (ms) of fn() on torch's current stream, L2 flushed before every launch."""
    import torch
    evs = []
    for i in range(reps):
        flush.fill_(i & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs[skip:])


def _traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel)
    except Exception:
        return None


def time_hot_kernels(device, peak):
    """The own kernels that dominate the step, each alone at its largest shape of the workload (batch 8):
    the three tensor-core GEMMs of cost-volume-1's first shared-MLP layer (rows = 8 x 228 x 80, 262 -> 128)
    and the projection-window select at SA1.  -> (roofline of the dominant kernel, list for the others)"""
    import torch
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    from i2pnet_b200.projectPN.utils import FLAG_COPY, FLAG_SHIFT, StrideGrid, project_seq, select_flat
    from i2pnet_b200.synthetic import make_pairs
    L = _cabi.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    rows, cin, cout = BATCH * 228 * 80, 262, 128
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.randn(rows, cin, device=device, generator=g)
    w, b = torch.randn(cout, cin, device=device, generator=g) * 0.1, torch.randn(cout, device=device, generator=g)
    gam, bet = torch.ones(cout, device=device), torch.zeros(cout, device=device)
    y, gr = torch.empty(rows, cout, device=device), torch.randn(rows, cout, device=device, generator=g)
    tiles = torch.empty(L.i2p_pw_num_tiles(rows), cout, 2, device=device)
    st = torch.empty(4, cout, device=device)
    pack = torch.empty(L.i2p_pw_pack_floats(cin, cout), device=device)
    s12 = torch.zeros(2, cout, dtype=torch.float64, device=device)
    dx, dw = torch.empty(rows, cin, device=device), torch.zeros(cout, cin, device=device)
    call("i2p_pw_pack_weights", device, cin, cout, w.data_ptr(), pack.data_ptr())
    fwd = lambda: call("i2p_pw_linear_fwd_tc", device, rows, cin, cout, x.data_ptr(), None, None, 1.0, pack.data_ptr(),
                       b.data_ptr(), y.data_ptr(), tiles.data_ptr())
    fwd()
    call("i2p_bn_finalize", device, rows, cout, tiles.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-5, st[0].data_ptr(),
         st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr())
    bn = (y.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), 0.1)
    none_prev = (None, None, None, None, None, 1.0, None)
    dxf = lambda: call("i2p_pw_linear_bwd_dx_tc", device, rows, cin, cout, gr.data_ptr(), None, None, 1, *bn, s12.data_ptr(), pack.data_ptr(),
                       dx.data_ptr(), *none_prev)
    dwf = lambda: call("i2p_pw_linear_bwd_dw_tc", device, rows, cin, cout, gr.data_ptr(), None, None, 1, *bn, s12.data_ptr(), x.data_ptr(),
                       None, None, 1.0, dw.data_ptr())
    d = make_pairs(BATCH, N_POINTS, IMAGE_HW, seed=7)
    _, (cam,) = project_seq(d["raw_point_xyz"].to(device), [d["lidar"].to(device)], 64, 1800, False)
    grid = StrideGrid(BATCH, 16, 225, 4, 8, device)
    sel = lambda: select_flat(cam, cam, grid, [9, 15], 32, FLAG_SHIFT | FLAG_COPY, 0.75)
    n, K = 3600, 32
    shape = "cost-volume-1 mlp1 layer 1 (rows 8x228x80 = %d, 262 -> 128), batch 8" % rows
    specs = [   # name, launcher, algorithmic bytes per launch
        ("tc::dx_kernel<128> @ " + shape, dxf, 4 * rows * (2 * cout + cin)),      # largest share of the step (profiles/r1_launches_step.md)
        ("tc::dw_kernel<128,2> @ " + shape, dwf, 4 * rows * (2 * cout + cin)),
        ("tc::fwd_kernel<128,2> @ " + shape, fwd, 4 * rows * (cin + cout)),
        ("select_k_kernel<5,flat> @ SA1 (64x1800, 3600 centres, 9x15, K=32), batch 8", sel, BATCH * (12 * 64 * 1800 + 8 * n * K)),
    ]
    # the image pyramid's convolution at its most frequent large shape (16 -> 16 channels at 80 x 256, three layers)
    cb, cc, ch_, cw_ = BATCH, 16, 80, 256
    cx = torch.randn(cb, cc, ch_, cw_, device=device, generator=g)
    cwt = torch.randn(cc, cc, 3, 3, device=device, generator=g) * 0.1
    cy, cdy, cdw = torch.empty_like(cx), torch.randn(cb, cc, ch_, cw_, device=device, generator=g), torch.zeros_like(cwt)
    cbias = torch.zeros(cc, device=device)
    cst = torch.empty(cc, L.i2p_conv3x3_stat_slots(cb, cc, ch_, cw_), 3, device=device)
    cpk = torch.empty(L.i2p_conv3x3_pack_floats(cc, cc, 0), device=device)
    call("i2p_conv3x3_pack", device, cc, cc, 0, cwt.data_ptr(), cpk.data_ptr())
    conv_f = lambda: call("i2p_conv3x3_tc", device, cb, cc, cc, ch_, cw_, cx.data_ptr(), cpk.data_ptr(), cbias.data_ptr(), cy.data_ptr(),
                          cst.data_ptr())
    conv_w = lambda: call("i2p_conv3x3_wgrad", device, cb, cc, cc, ch_, cw_, cx.data_ptr(), cdy.data_ptr(), cdw.data_ptr())
    conv_bytes = 4 * cb * ch_ * cw_ * (cc + cc)
    specs += [("conv::conv3x3_tc_kernel<16,16> @ RGB pyramid 16 -> 16 at 80x256 (tcgen05 3xTF32 implicit GEMM + BN statistics), batch 8", conv_f, conv_bytes),
              ("conv::wgrad_kernel<16,32> @ RGB pyramid 16 -> 16 at 80x256 (f32 FMA), batch 8", conv_w, conv_bytes)]
    out = []
    for name, fn, alg in specs:
        ms = _event_time(fn, flush)
        achieved = alg / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak[0], "peak_source": peak[1],
                    "unit": "GB/s", "frac": achieved / peak[0], "traffic": _traffic(name.split(" @")[0]),
                    "algorithmic_bytes": alg, "us_per_launch": ms * 1e3})
    return out[0], out[1:]


def run_ours(args):
    import torch
    import torch.distributed as dist
    from i2pnet_b200 import _cabi
    from i2pnet_b200.engine import INPUT_KEYS, HostPipeline, TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; i2pnet_b200 has no CPU path (use --impl reference)")
    _cabi.lib()  # fail loudly here if the sm_100a library is missing
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = False   # f32 parity configuration (north_star: 1e-4 relative)
    torch.backends.cudnn.allow_tf32 = False

    conf = CONFIGS[args.config]
    batch = _per_gpu_batch(conf, world, args.strong)
    eng = TrainStep(batch, conf["points"], conf["image"], cfg=_cfg_of(conf), device=device, seed=0,
                    use_graph=not args.no_graph, fused_optimizer=not args.stock_optimizer)
    nb = 4  # distinct batches, cycled
    host = [_pairs(conf, batch, 100 * rank + i) for i in range(nb)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    dev_batches = [{k: v.to(device) for k, v in b.items()} for b in host]
    eng.load(dev_batches[0])
    eng.warmup_and_capture(eager_steps=3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    pipe = HostPipeline(eng)

    def timed(n_steps, from_host):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        if from_host:
            # end to end: every step's batch comes from pinned host memory (copied while the previous step runs) and
            # every step's loss goes back to the host (read one step later); the region ends when the last loss has landed
            pipe.submit(host[0])
            for i in range(n_steps):
                flush.fill_(i & 0xff)
                pipe.step()
                if i + 1 < n_steps:
                    pipe.submit(host[(i + 1) % nb])
                pipe.loss()
                pipe.out3()
            pipe.drain()
        else:
            for i in range(n_steps):
                flush.fill_(i & 0xff)
                eng.load(dev_batches[i % nb])
                eng.step()
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    timed(max(args.warmup, 3), False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    timed(3, True)
    ms_e2e = timed(args.steps, True)
    loss = float(eng.loss.item())

    if rank == 0:
        peak = _peaks()
        roof, roof_others = time_hot_kernels(device, peak)
        value = world * batch * args.steps / (ms * 1e-3)
        e2e = world * batch * args.steps / (ms_e2e * 1e-3)
        h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in INPUT_KEYS)
        cpu = legacy = None
        if world == 1 and not args.no_cpu_baseline and not conf["nus"]:
            cb = min(batch, 8)
            t, cores = cpu_reference_steps(3, 1, cb, conf=conf)
            cpu = {"value": cb * len(t) / sum(t), "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": "3 steps of batch %d (of the batch-%d workload) after 1 warm-up, oracle/model_cpu.py "
                             "fwd+bwd+clip+Adam" % (cb, batch)}
        if world == 1 and not args.no_legacy:
            legacy = legacy_b200(conf, batch)
        config = _config_dict(conf, args.config, batch, world, args.strong)
        engine = dict({"cuda_graph": not args.no_graph, "optimizer": "torch.optim.Adam" if args.stock_optimizer else "fused clip+Adam (2 launches)",
                       "side_streams": os.environ.get("I2P_STREAMS", "1") != "0", "tf32": False,
                       "shared_mlp": "tcgen05 3xTF32 split (f32-accurate), mask %d" % _cabi.lib().i2p_get_mlp_tensor_cores(),
                       "rgb_convolutions": "own tcgen05 3xTF32 implicit GEMM (fwd, dgrad) + f32 FMA wgrad; no cuDNN",
                       "final_loss": loss})
        print(json.dumps({
            "metric": "pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "engine": engine,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": pipe.d2h_bytes_per_step,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(eng.launches_per_step) * args.steps,
            "gpu_launches_per_step": int(eng.launches_per_step),
            "clocks": clocks, "roofline": roof, "roofline_others": roof_others, "cpu_baseline": cpu,
            "legacy_b200": legacy,
        }))
    if world > 1:
        # Tear-down: the captured graph holds NCCL kernels, and destroying the communicator underneath it was
        # observed to hang (N = 2, after the JSON line was already printed).  All ranks meet at a barrier, the
        # graph is released first, and the processes leave without the communicator destructor.
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize(device)
        eng.graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of CUDA-graph replay")
    ap.add_argument("--config", default="kitti_b8", choices=sorted(CONFIGS),
                    help="kitti_b8: BASELINE configs[1] (default); kitti_b32: north_star's target batch; nus_b16: configs[3]")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the config's batch is the GLOBAL batch, split evenly over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legacy", action="store_true", help="skip the legacy-kernels-on-B200 leg (oracle/ref_live.py)")
    ap.add_argument("--stock-optimizer", action="store_true", help="torch.optim.Adam + clip instead of the fused flat step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
