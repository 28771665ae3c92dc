"""Training-step engine: forward + loss + backward + gradient all-reduce + clip + Adam for RegNet_v2,
one process per GPU.

What the reference does per iteration (train20v2learn_wandb_proj.py:435-483): eight separate
`.to(device)` copies, an eager forward with a host round-trip inside it, `loss.item()`,
`loss.backward()`, `clip_grad_norm_(10)`, `Adam.step()`; single GPU only.
What this does:
  * all gradients live in ONE flat f32 buffer (`p.grad` are views into it), so the data-parallel
    exchange is a single NCCL all-reduce of 3.4 MB over NVLink, and zeroing / clipping touch one
    tensor;
  * the whole step -- every torch kernel and every libi2p_b200.so launch -- is captured once
    into a CUDA graph and replayed; inputs are copied into static device buffers (from pinned host
    memory on the end-to-end path), the loss is read back from a static scalar.
The optimiser and its hyper-parameters are the reference's (Adam, lr 1e-3, weight decay 1e-4,
clip 10, train...proj.py:198-205,481).
"""
import torch
import torch.distributed as dist

from . import _cabi, scratch, streams
from .compute_loss import Get_loss
from .config_proj_lidarcenter import I2PNetConfig
from .modellearn_proj_center import RegNet_v2
from .projectPN import fused_mlp

INPUT_KEYS = ("rgb", "lidar", "raw_point_xyz", "lidar_feats", "intrinsic", "q_gt", "t_gt")


class FlatGradBucket:
    """One contiguous gradient buffer for all parameters.

    Autograd accumulates into an existing `p.grad` with one `add_` launch per parameter (299 tiny
    kernels per step here).  Instead the flat buffer is zeroed once per step (`begin_step()`), `p.grad` is dropped
    so that autograd simply keeps the gradient tensors its functions produce, and `gather()` adds those into their
    slices with a multi-tensor add; `p.grad` then alias slices of the flat buffer for the all-reduce, the clip and
    the optimiser.

    Gradient sinks: every parameter carries `p._i2p_sink`, its slice of the flat buffer.  A backward function of this
    package that computes a parameter gradient with its own kernel may accumulate straight into the sink (on whatever
    stream it likes, e.g. a weight-gradient side stream) and return None for that parameter: no tensor produced off
    the autograd node's stream is ever handed to autograd (whose AccumulateGrad would read it without waiting for
    that stream), and the copy into the flat buffer disappears.  `sinks_live` tells those functions the buffer has
    been zeroed for this step."""

    def __init__(self, params, align=1):
        """align: every parameter's slice starts at a multiple of `align` elements (zero padding in between), so that a
        second flat buffer with the same layout can hold the parameters themselves at a vector-load alignment."""
        self.all_params = list(params)
        self.params = [p for p in self.all_params if p.requires_grad]
        ref = self.params[0]
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // align) * align
        self.flat = torch.zeros(off, dtype=ref.dtype, device=ref.device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.sinks_live = False
        for p, v in zip(self.params, self.views):
            p._i2p_sink = v
            p._i2p_bucket = self
        self._alias()

    def _alias(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def release(self):
        for p in self.params:
            p.grad = None

    def begin_step(self):
        """Zero the flat buffer (one memset) and drop every p.grad: from here until gather() the sinks accumulate."""
        self.flat.zero_()
        self.release()
        self.sinks_live = True

    def gather(self):
        """Add the gradients autograd produced into their slices (what the sinks received is already there)."""
        if not self.sinks_live:        # no begin_step(): the slices still hold the previous step
            self.flat.zero_()
        self.sinks_live = False
        views, grads = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                views.append(v)
                grads.append(p.grad.reshape(v.shape) if p.grad.shape != v.shape else p.grad)
        if views:
            torch._foreach_add_(views, grads)
        self._alias()

    def zero(self):
        self.flat.zero_()
        self._alias()

    def all_reduce_mean(self, group=None):
        """Sum over ranks / world size: the single exchange step of the data-parallel path."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def all_reduce_sum(self, group=None):
        """Sum over ranks, left un-averaged (the fused optimiser step divides); -> world size."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def clip_(self, max_norm):
        """clip_grad_norm_ on the flat buffer, no host synchronisation."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


def grad_sink(p):
    """The flat-buffer slice a backward kernel may accumulate parameter p's gradient into, or None (no engine, or
    outside begin_step() .. gather())."""
    b = getattr(p, "_i2p_bucket", None)
    return p._i2p_sink if b is not None and b.sinks_live else None


class FlatAdam:
    """clip_grad_norm_(max_norm) + torch.optim.Adam(lr, betas, eps, weight_decay).step() of the reference trainer
    (train20v2learn_wandb_proj.py:198-205, 481-483) as two launches of csrc/optim.cu on flat buffers.

    The parameters are re-seated as views into one flat f32 buffer laid out like the gradient bucket (`bucket.offsets`),
    both moments are flat buffers of the same layout, and the step count AND the learning rate live on the device, so
    the update is capturable into a CUDA graph and still follows a learning-rate schedule (`set_lr`).  The gradient
    buffer is expected to hold the SUM over `world` ranks."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=10.0):
        self.bucket, self.betas, self.eps, self.weight_decay, self.max_norm = bucket, betas, eps, weight_decay, max_norm
        flat = bucket.flat
        self.param = torch.zeros_like(flat)
        with torch.no_grad():
            for p, o in zip(bucket.params, bucket.offsets):
                view = self.param[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        lib = _cabi.lib()
        self.state = torch.zeros(lib.i2p_optim_state_bytes(), dtype=torch.uint8, device=flat.device)
        off = lib.i2p_optim_lr_offset()
        self._lr_dev = self.state[off:off + 4].view(torch.float32)
        self.initial_lr = float(lr)
        self.set_lr(lr)
        # torch.optim.Adam numbers its state by position in the parameter list it was given -- model.parameters(),
        # frozen tensors included
        pos = {id(p): i for i, p in enumerate(bucket.all_params)}
        self.indices = [pos[id(p)] for p in bucket.params]

    @property
    def lr(self):
        return self._lr

    def set_lr(self, lr):
        """Write the learning rate into the device state: the next step -- eager or a replay of a captured graph --
        uses it (the reference steps ExponentialLR(0.99) once per epoch, train20v2learn_wandb_proj.py:205, 524)."""
        self._lr = float(lr)
        self._lr_dev.fill_(self._lr)

    # ---- checkpoint / resume in torch.optim.Adam's own format, so that the reference trainer's
    # `optimizer_state_dict` (train20v2learn_wandb_proj.py:220, 259) loads here and vice versa
    def _step_count(self):
        return self.state[:32].view(torch.float64)[1]

    def state_dict(self):
        b, step = self.bucket, self._step_count().to(torch.float32).cpu()
        state = {}
        if float(step) > 0:
            for i, p, o in zip(self.indices, b.params, b.offsets):
                state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     decoupled_weight_decay=False, initial_lr=self.initial_lr, params=list(range(len(b.all_params))))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        b = self.bucket
        group = sd["param_groups"][0]
        if len(sd["param_groups"]) != 1 or len(group["params"]) != len(b.all_params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group over %d tensors" % len(b.all_params))
        self.betas, self.eps, self.weight_decay = tuple(group["betas"]), group["eps"], group["weight_decay"]
        self.initial_lr = float(group.get("initial_lr", group["lr"]))
        self.set_lr(group["lr"])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        with torch.no_grad():
            for i, p, o in zip(self.indices, b.params, b.offsets):
                st = sd["state"].get(i)
                if st is None:
                    continue
                self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                steps.add(float(st["step"]))
            if len(steps) > 1:
                raise ValueError("FlatAdam.load_state_dict: parameters with different step counts %s" % sorted(steps))
            self._step_count().fill_(steps.pop() if steps else 0.0)

    def snapshot(self):
        return [self.param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.state.clone()]

    def restore(self, snap):
        for dst, src in zip((self.param, self.exp_avg, self.exp_avg_sq, self.state), snap):
            dst.copy_(src)

    def step(self, world=1):
        b = self.bucket
        _cabi.call("i2p_clip_adam_step", b.flat.device, b.flat.numel(), self.param.data_ptr(), b.flat.data_ptr(),
                   self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.state.data_ptr(), -1.0,   # lr: device state
                   float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), float(self.max_norm),
                   int(world))


class TrainStep:
    def __init__(self, batch, n_points=20480, image_hw=(160, 512), cfg=I2PNetConfig, device="cuda:0", seed=0,
                 use_graph=True, lr=1e-3, weight_decay=1e-4, clip=10.0, group=None, fused_optimizer=True, lr_gamma=0.99):
        self.device = torch.device(device)
        self.cfg, self.batch, self.clip, self.group, self.use_graph = cfg, batch, clip, group, use_graph
        torch.manual_seed(seed)
        self.model = RegNet_v2(cfg=cfg).to(self.device)
        self.model.train()
        self.fused_optimizer = fused_optimizer
        if self.fused_optimizer:
            self.bucket = FlatGradBucket(self.model.parameters(), align=64)
            self.opt = FlatAdam(self.bucket, lr=lr, weight_decay=weight_decay, max_norm=clip)
        else:   # the stock optimiser (A/B measurements, parity tests of the fused one)
            self.bucket = FlatGradBucket(self.model.parameters())
            self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay, capturable=use_graph,
                                        foreach=True)
        # the trainer's ExponentialLR(optimizer, 0.99), stepped once per epoch (train20v2learn_wandb_proj.py:205, 524)
        self.lr_gamma, self.base_lr, self.epoch, self.sched_steps = float(lr_gamma), float(lr), 0, 0
        h, w = image_hw
        shapes = dict(rgb=(batch, 3, h, w), lidar=(batch, n_points, 3), raw_point_xyz=(batch, n_points, 3),
                      lidar_feats=(batch, n_points, 1), intrinsic=(batch, 3, 3), q_gt=(batch, 4), t_gt=(batch, 3))
        self.inputs = {k: torch.zeros(s, device=self.device) for k, s in shapes.items()}
        self.loss = torch.zeros(1, device=self.device)
        # the refined and the coarse pose of the step: the trainer reads out_3 back every iteration for its running
        # RRE / RTE (cal_rete_once, train20v2learn_wandb_proj.py:485)
        self.out3 = torch.zeros(batch, 7, device=self.device)
        self.out4 = torch.zeros(batch, 7, device=self.device)
        self.graph = None
        self.launches_per_step = None
        self._loaded = False

    # ---- one eager step on the static buffers
    def _step_body(self):
        x = self.inputs
        self.bucket.begin_step()
        scratch.begin_step(self.device)          # the step's zero-initialised work space: one memset (scratch.py)
        fused_mlp.packs_begin_step(self.device)  # tensor-core copies of every shared-MLP weight: one launch (fused_mlp.py)
        try:
            self._forward_backward(x)
        finally:
            scratch.end_step()
            fused_mlp.packs_end_step()

    def _forward_backward(self, x):
        out3, out4, _, _, sx, sq = self.model(x["rgb"], x["lidar"], x["raw_point_xyz"], None, x["intrinsic"], None,
                                              None, None, x["lidar_feats"], self.cfg)
        loss, _, _ = Get_loss(out3, out4, x["q_gt"], x["t_gt"], sx, sq, self.cfg)
        streams.DEFER_JOINS = True           # weight-gradient branches float until the whole backward has been issued
        try:
            loss.backward()
        finally:
            streams.DEFER_JOINS = False
        streams.join_pending()
        self.bucket.gather()
        if self.fused_optimizer:
            world = self.bucket.all_reduce_sum(self.group)      # averaging, clipping and the update: one fused step
            self.opt.step(world)
        else:
            self.bucket.all_reduce_mean(self.group)
            self.bucket.clip_(self.clip)
            self.opt.step()
        self.loss.copy_(loss.detach())
        self.out3.copy_(out3.detach())
        self.out4.copy_(out4.detach())

    def load(self, batch_dict, non_blocking=True):
        """Copy a batch (device or pinned-host tensors) into the static input buffers."""
        for k in INPUT_KEYS:
            self.inputs[k].copy_(batch_dict[k], non_blocking=non_blocking)
        self._loaded = True

    # ---- learning-rate schedule (device state: no re-capture)
    def current_lr(self):
        return self.opt.lr if self.fused_optimizer else self.opt.param_groups[0]["lr"]

    def set_lr(self, lr):
        if self.fused_optimizer:
            self.opt.set_lr(lr)
        else:
            for g in self.opt.param_groups:      # capturable Adam keeps lr as a python float: eager mode only
                g["lr"] = lr
            if self.graph is not None:
                raise RuntimeError("the stock optimiser bakes lr into the captured graph; use the fused optimiser")

    def scheduler_step(self):
        """ExponentialLR.step(): lr = base_lr * gamma ** epochs, once per epoch."""
        self.sched_steps += 1
        self.epoch += 1
        self.set_lr(self.base_lr * self.lr_gamma ** self.sched_steps)

    def _snapshot(self):
        snap = {"buffers": [b.clone() for b in self.model.buffers()]}
        if self.fused_optimizer:
            snap["opt"] = self.opt.snapshot()
        else:
            import copy
            snap["params"] = [p.detach().clone() for p in self.model.parameters()]
            snap["opt"] = copy.deepcopy(self.opt.state_dict())
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            for b, v in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(v)
            if self.fused_optimizer:
                self.opt.restore(snap["opt"])
            else:
                for p, v in zip(self.model.parameters(), snap["params"]):
                    p.copy_(v)
                if snap["opt"]["state"]:
                    self.opt.load_state_dict(snap["opt"])
                else:       # fresh optimiser: keep the (capturable) state tensors the warm-up created, zeroed
                    for st in self.opt.state.values():
                        for v in st.values():
                            if isinstance(v, torch.Tensor):
                                v.zero_()

    def warmup_and_capture(self, eager_steps=3):
        """Eager warm-up on a side stream (allocator, library handles, lazily built kernels), then capture.  The
        warm-up steps are real training steps on whatever batch was loaded, so parameters, optimiser moments, step
        counter and BatchNorm running statistics are snapshotted before and restored after: warm-up followed by a
        checkpoint load, or the other way round, leaves the training state exactly as loaded."""
        if not self._loaded:
            raise RuntimeError("TrainStep.warmup_and_capture: load() a real batch first (all-zero inputs put NaNs "
                               "into the range-image maths)")
        snap = self._snapshot()
        s = streams.main_stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(eager_steps):
                before = _cabi.launch_count()
                self._step_body()
                self.launches_per_step = _cabi.launch_count() - before
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._restore(snap)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=s):
                self._step_body()
            torch.cuda.synchronize(self.device)

    def step(self):
        """One training step on whatever the static input buffers hold; asynchronous."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def scheduler_state_dict(self):
        """torch.optim.lr_scheduler.ExponentialLR.state_dict()'s keys (the trainer saves and reloads it, :220, 260)."""
        lr = self.current_lr()
        return {"gamma": self.lr_gamma, "base_lrs": [self.base_lr], "last_epoch": self.sched_steps,
                "_step_count": self.sched_steps + 1, "_get_lr_called_within_step": False, "_last_lr": [lr],
                "_is_initial": False}

    def state_dict(self):
        """Checkpoint with the keys the reference trainer writes and reads back unconditionally on resume
        (train20v2learn_wandb_proj.py:218-221, 255-260)."""
        return {"model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.opt.state_dict(),
                "scheduler_state_dict": self.scheduler_state_dict(), "epoch": self.epoch}

    def load_state_dict(self, ckpt):
        """Resume: parameters, buffers, moments, step count and learning rate are written IN PLACE (the flat buffers and
        a captured graph stay valid; the learning rate is device state, so replays pick it up)."""
        self.model.load_state_dict(ckpt["model_state_dict"])
        if ckpt.get("optimizer_state_dict") is not None:
            self.opt.load_state_dict(ckpt["optimizer_state_dict"])
        sch = ckpt.get("scheduler_state_dict")
        if sch is not None:
            self.lr_gamma, self.sched_steps = float(sch["gamma"]), int(sch["last_epoch"])
            self.base_lr = float(sch["base_lrs"][0])
            if sch.get("_last_lr"):
                self.set_lr(float(sch["_last_lr"][0]))
        if "epoch" in ckpt:
            self.epoch = int(ckpt["epoch"])

    def step_from_host(self, host_batch):
        """End-to-end form: pinned host batch -> device, one step, loss back on the host."""
        self.load(host_batch, non_blocking=True)
        self.step()
        return float(self.loss.item())


class HostPipeline:
    """Double-buffered end-to-end feed of a TrainStep (the reference trainer copies eight
    tensors synchronously and calls loss.item() every iteration, train20v2learn_wandb_proj.py:435-470).

    Batch i + 1 travels from pinned host memory to a device staging set on a copy stream while step i runs; the step
    stream then waits for that copy, moves the staging set into the graph's static inputs (device to device) and replays.
    Each step's loss and refined pose out_3 (the trainer reads both back every iteration: loss.item() :469,
    cal_rete_once(out3) :485) are copied to pinned host memory asynchronously and handed out one call later, so the
    host never blocks on the step it has just launched.

        pipe = HostPipeline(eng)
        pipe.submit(batches[0])
        for i in range(n):
            pipe.step()                          # step i on the batch submitted last
            if i + 1 < n: pipe.submit(batches[i + 1])
            prev = pipe.loss()                   # loss of step i - 1 (None at i = 0); pipe.drain() -> the last one
    """

    def __init__(self, eng):
        self.eng = eng
        dev = eng.device
        self.staging = {k: torch.empty_like(v) for k, v in eng.inputs.items()}
        self.copy_stream = torch.cuda.Stream(dev)
        self.copied, self.consumed = torch.cuda.Event(), torch.cuda.Event()
        self.consumed.record(torch.cuda.current_stream(dev))
        self.loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
        self.out3_host = [torch.zeros(eng.batch, 7).pin_memory(), torch.zeros(eng.batch, 7).pin_memory()]
        self.d2h_bytes_per_step = 4 + eng.batch * 7 * 4
        self.loss_ready = [None, None]
        self.n_steps = 0

    def submit(self, host_batch):
        """Start the host-to-device copy of the next batch (pinned tensors)."""
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed)        # the staging set has been moved into the static inputs
            for k in INPUT_KEYS:
                self.staging[k].copy_(host_batch[k], non_blocking=True)
            self.copied.record(self.copy_stream)

    def step(self):
        eng = self.eng
        main = torch.cuda.current_stream(eng.device)
        main.wait_event(self.copied)
        for k in INPUT_KEYS:
            eng.inputs[k].copy_(self.staging[k], non_blocking=True)
        self.consumed.record(main)
        eng.step()
        slot = self.n_steps & 1
        self.loss_host[slot].copy_(eng.loss, non_blocking=True)
        self.out3_host[slot].copy_(eng.out3, non_blocking=True)     # the trainer's cal_rete_once(out3, ...) read-back (:485)
        ev = torch.cuda.Event()
        ev.record(main)
        self.loss_ready[slot] = ev
        self.n_steps += 1

    def loss(self):
        """Loss of the step before the one launched last (its copy has long landed), or None before there is one."""
        if self.n_steps < 2:
            return None
        slot = (self.n_steps - 2) & 1
        self.loss_ready[slot].synchronize()
        return float(self.loss_host[slot])

    def out3(self):
        """out_3 (B,7) of the step before the one launched last (pinned host tensor), or None."""
        if self.n_steps < 2:
            return None
        slot = (self.n_steps - 2) & 1
        self.loss_ready[slot].synchronize()
        return self.out3_host[slot]

    def drain(self):
        """Wait for everything; -> loss of the last step."""
        slot = (self.n_steps - 1) & 1
        self.loss_ready[slot].synchronize()
        return float(self.loss_host[slot])
