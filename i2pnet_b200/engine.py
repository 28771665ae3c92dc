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

from . import _cabi
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

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._alias()

    def _alias(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def release(self):
        for p in self.params:
            p.grad = None

    def gather(self):
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
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

    def clip_(self, max_norm):
        """clip_grad_norm_ on the flat buffer, no host synchronisation."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


class TrainStep:
    def __init__(self, batch, n_points=20480, image_hw=(160, 512), cfg=I2PNetConfig, device="cuda:0", seed=0,
                 use_graph=True, lr=1e-3, weight_decay=1e-4, clip=10.0, group=None, channels_last_rgb=False,
                 cudnn_benchmark=False):
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
        loss.backward()
        self.bucket.gather()
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

    def step_from_host(self, host_batch):
        """End-to-end form: pinned host batch -> device, one step, loss back on the host."""
        self.load(host_batch, non_blocking=True)
        self.step()
        return float(self.loss.item())
