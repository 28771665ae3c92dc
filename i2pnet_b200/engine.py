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

from . import _cabi, streams
from .compute_loss import Get_loss
from .config_proj_lidarcenter import I2PNetConfig
from .modellearn_proj_center import RegNet_v2

INPUT_KEYS = ("rgb", "lidar", "raw_point_xyz", "lidar_feats", "intrinsic", "q_gt", "t_gt")


class FlatGradBucket:
    """One contiguous gradient buffer for all parameters.

    Autograd accumulates into an existing `p.grad` with one `add_` launch per parameter (299 tiny
    kernels per step here) and the buffer would need a memset first.  Instead `release()` drops every
    `p.grad` before backward, so that autograd simply keeps the gradient tensors its functions produce,
    and `gather()` packs them into the flat buffer with a single batched concatenation; `p.grad` then
    alias slices of the flat buffer for the all-reduce, the clip and the optimiser."""

    def __init__(self, params, align=1):
        """align: every parameter's slice starts at a multiple of `align` elements (zero padding in between), so that a
        second flat buffer with the same layout can hold the parameters themselves at a vector-load alignment."""
        self.params = [p for p in params if p.requires_grad]
        ref = self.params[0]
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // align) * align
        self.flat = torch.zeros(off, dtype=ref.dtype, device=ref.device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self._pads = [torch.zeros(self.offsets[i + 1] - self.offsets[i] - p.numel() if i + 1 < len(self.params)
                                  else off - self.offsets[i] - p.numel(), dtype=ref.dtype, device=ref.device)
                      for i, p in enumerate(self.params)]
        self._alias()

    def _alias(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def release(self):
        for p in self.params:
            p.grad = None

    def gather(self):
        pieces = []
        for p, pad in zip(self.params, self._pads):
            pieces.append((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1))
            if pad.numel():
                pieces.append(pad)
        torch.cat(pieces, out=self.flat)
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


class FlatAdam:
    """clip_grad_norm_(max_norm) + torch.optim.Adam(lr, betas, eps, weight_decay).step() of the reference trainer
    (train20v2learn_wandb_proj.py:198-205, 481-483) as two launches of csrc/optim.cu on flat buffers.

    The parameters are re-seated as views into one flat f32 buffer laid out like the gradient bucket (`bucket.offsets`),
    both moments are flat buffers of the same layout, and the step count lives on the device, so the update is
    capturable into a CUDA graph.  The gradient buffer is expected to hold the SUM over `world` ranks."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=10.0):
        self.bucket, self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = bucket, lr, betas, eps, weight_decay, max_norm
        flat = bucket.flat
        self.param = torch.zeros_like(flat)
        with torch.no_grad():
            for p, o in zip(bucket.params, bucket.offsets):
                view = self.param[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        self.state = torch.zeros(_cabi.lib().i2p_optim_state_bytes(), dtype=torch.uint8, device=flat.device)

    # ---- checkpoint / resume in torch.optim.Adam's own format, so that the reference trainer's
    # `optimizer_state_dict` (train20v2learn_wandb_proj.py:220, 259) loads here and vice versa
    def _step_count(self):
        return self.state[:32].view(torch.float64)[1]

    def state_dict(self):
        b, step = self.bucket, self._step_count().to(torch.float32).cpu()
        state = {}
        if float(step) > 0:
            for i, (p, o) in enumerate(zip(b.params, b.offsets)):
                state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(b.params))))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        b = self.bucket
        group = sd["param_groups"][0]
        if len(sd["param_groups"]) != 1 or len(group["params"]) != len(b.params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group over %d tensors" % len(b.params))
        self.lr, self.betas, self.eps, self.weight_decay = group["lr"], tuple(group["betas"]), group["eps"], group["weight_decay"]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(b.params, b.offsets)):
                st = sd["state"].get(i)
                if st is None:
                    continue
                self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                steps.add(float(st["step"]))
            if len(steps) > 1:
                raise ValueError("FlatAdam.load_state_dict: parameters with different step counts %s" % sorted(steps))
            self._step_count().fill_(steps.pop() if steps else 0.0)

    def step(self, world=1):
        b = self.bucket
        _cabi.call("i2p_clip_adam_step", b.flat.device, b.flat.numel(), self.param.data_ptr(), b.flat.data_ptr(),
                   self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.state.data_ptr(), float(self.lr),
                   float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), float(self.max_norm),
                   int(world))


class TrainStep:
    def __init__(self, batch, n_points=20480, image_hw=(160, 512), cfg=I2PNetConfig, device="cuda:0", seed=0,
                 use_graph=True, lr=1e-3, weight_decay=1e-4, clip=10.0, group=None, channels_last_rgb=False,
                 cudnn_benchmark=False, fused_optimizer=True):
        self.device = torch.device(device)
        self.cfg, self.batch, self.clip, self.group, self.use_graph = cfg, batch, clip, group, use_graph
        if cudnn_benchmark:   # let cuDNN time its f32 algorithms for the 15 convolutions during the eager warm-up
            torch.backends.cudnn.benchmark = True
        torch.manual_seed(seed)
        self.model = RegNet_v2(cfg=cfg).to(self.device)
        self.model.train()
        self.channels_last_rgb = channels_last_rgb
        if channels_last_rgb:  # NHWC image branch: ATen's channels-last batch-norm / pooling kernels
            for name in ("RGB_net1", "RGB_net2", "RGB_net3"):
                getattr(self.model, name).to(memory_format=torch.channels_last)
        self.fused_optimizer = fused_optimizer and not channels_last_rgb
        if self.fused_optimizer:
            self.bucket = FlatGradBucket(self.model.parameters(), align=64)
            self.opt = FlatAdam(self.bucket, lr=lr, weight_decay=weight_decay, max_norm=clip)
        else:   # the stock optimiser (A/B measurements, parity tests of the fused one)
            self.bucket = FlatGradBucket(self.model.parameters())
            self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay, capturable=use_graph,
                                        foreach=True)
        h, w = image_hw
        shapes = dict(rgb=(batch, 3, h, w), lidar=(batch, n_points, 3), raw_point_xyz=(batch, n_points, 3),
                      lidar_feats=(batch, n_points, 1), intrinsic=(batch, 3, 3), q_gt=(batch, 4), t_gt=(batch, 3))
        self.inputs = {k: torch.zeros(s, device=self.device) for k, s in shapes.items()}
        if channels_last_rgb:
            self.inputs["rgb"] = self.inputs["rgb"].contiguous(memory_format=torch.channels_last)
        self.loss = torch.zeros(1, device=self.device)
        self.graph = None
        self.launches_per_step = None

    # ---- one eager step on the static buffers
    def _step_body(self):
        x = self.inputs
        self.bucket.release()
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

    def load(self, batch_dict, non_blocking=True):
        """Copy a batch (device or pinned-host tensors) into the static input buffers."""
        for k in INPUT_KEYS:
            self.inputs[k].copy_(batch_dict[k], non_blocking=non_blocking)

    def warmup_and_capture(self, eager_steps=3):
        """Eager warm-up on a side stream (allocator, cuBLAS handles, Adam state), then capture."""
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(eager_steps):
                before = _cabi.launch_count()
                self._step_body()
                self.launches_per_step = _cabi.launch_count() - before
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step_body()
            torch.cuda.synchronize(self.device)

    def step(self):
        """One training step on whatever the static input buffers hold; asynchronous."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def state_dict(self):
        """Checkpoint with the keys the reference trainer writes (train20v2learn_wandb_proj.py:255-260)."""
        return {"model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.opt.state_dict()}

    def load_state_dict(self, ckpt):
        """Resume: parameters and buffers are copied IN PLACE (the flat buffers and a captured graph stay valid)."""
        self.model.load_state_dict(ckpt["model_state_dict"])
        if ckpt.get("optimizer_state_dict") is not None:
            self.opt.load_state_dict(ckpt["optimizer_state_dict"])

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
    Each step's loss is copied to pinned host memory asynchronously and handed out one call later, so the host never
    blocks on the step it has just launched.

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

    def drain(self):
        """Wait for everything; -> loss of the last step."""
        slot = (self.n_steps - 1) & 1
        self.loss_ready[slot].synchronize()
        return float(self.loss_host[slot])
