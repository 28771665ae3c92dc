"""Branch-level concurrency: independent sub-graphs of the forward are issued on side CUDA streams.

Most kernels of this network (everything at pyramid levels 2-4, the second cost-volume stage, the heads) occupy a
fraction of the 148 SMs, and the step is a long chain of them.  Where the data flow forks -- image pyramid vs LiDAR
pyramid, the two up-convolutions, the position encodings of a cost volume vs its feature MLP -- one side of the fork
is issued on another stream, forked from and joined back into the current one with events, so that it is captured
into the same CUDA graph (and, in eager mode, simply overlaps).  Autograd replays every backward function on the
stream its forward ran on and inserts the cross-stream waits itself, so the backward chains overlap as well.

Memory safety with the caching allocator: tensors produced on one stream and consumed on the other are
`record_stream`-ed on the consumer (inputs at the fork, outputs at the join).

    with Fork(x, y) as f:          # x, y: tensors the branch reads
        z = branch(x, y)           # issued on a side stream
    w = other_work()               # current stream, concurrent with the branch
    z = f.join(z)                  # the current stream waits for the branch
"""
import os

import torch

ENABLED = os.environ.get("I2P_STREAMS", "1") != "0"
_POOL_SIZE = 4
_pools = {}

# Stream priorities (lower number = dispatched first).  The block scheduler hands out the blocks of the kernel that was
# launched first; a 200 us kernel with thousands of blocks on one stream therefore starves every small kernel launched
# after it on another stream until its last block has been dispatched -- in the replayed step the image branch's backward
# (a chain of fifteen 10-70 us kernels, the tail of the step) sat idle for half of its 1.9 ms behind the LiDAR pyramid's
# wide weight-gradient kernels (CUPTI timeline, tools/profile_step.py timeline).  With priorities the scheduler prefers
# the blocks of the chain kernels as soon as slots free up: branches (consumed by the chain) first, the chain itself
# second, weight gradients (consumed by nobody before the optimiser) last.  Captured kernel nodes keep the priority
# of the stream they were captured on.  I2P_STREAM_PRIORITIES=0: all streams at the default priority.
PRIORITIES = os.environ.get("I2P_STREAM_PRIORITIES", "1") != "0"
_PRIORITY = {"branch": -2, "main": -1, "wgrad": 0, "cwgrad": 0}     # cwgrad: the image pyramid's weight gradients
if os.environ.get("I2P_STREAM_PRIORITY_LEVELS"):        # tuning override: "branch,main,wgrad[,cwgrad]", e.g. "-2,-1,-1"
    _lv = [int(v) for v in os.environ["I2P_STREAM_PRIORITY_LEVELS"].split(",")]
    _PRIORITY = dict(zip(("branch", "main", "wgrad", "cwgrad"), _lv + _lv[2:3] * (4 - len(_lv))))


def priority_of(kind):
    return _PRIORITY[kind] if PRIORITIES else 0


def main_stream(device):
    """A stream of the chain's priority for a step engine to run / capture the step on."""
    return torch.cuda.Stream(device, priority=priority_of("main"))


def _side_stream(device, avoid, kind="branch"):
    """Round-robin over a small pool of side streams.  Two disjoint pools: "branch" for sub-graphs whose results the
    forward / backward chain consumes (image pyramid, up-convolutions, position encodings), and "wgrad" for weight-gradient
    work that nobody but the optimiser reads.  They must not share streams: a weight-gradient fork makes its side stream
    wait for the forking stream, and if that side stream were also the stream a branch lives on (the image pyramid's
    backward, say), the branch would inherit a dependency on wherever the forking chain happened to be -- measured: the
    image branch's backward started 0.9 ms late, serialised behind the LiDAR pyramid's backward."""
    pool = _pools.setdefault((device.index, kind), {"streams": [torch.cuda.Stream(device, priority=priority_of(kind)) for _ in range(_POOL_SIZE)], "next": 0})
    for _ in range(_POOL_SIZE):
        s = pool["streams"][pool["next"]]
        pool["next"] = (pool["next"] + 1) % _POOL_SIZE
        if s != avoid:
            return s
    raise RuntimeError("no side stream available")


class Fork:
    def __init__(self, *inputs, enabled=True, kind="branch"):
        self.inputs = [t for t in inputs if isinstance(t, torch.Tensor) and t.is_cuda]
        self.on = ENABLED and enabled and len(self.inputs) > 0
        self.ctx = self.side = None
        self.kind = kind

    def __enter__(self):
        if self.on:
            dev = self.inputs[0].device
            self.main = torch.cuda.current_stream(dev)
            if self.side is None:       # a Fork may be re-entered: later sections continue on the same side stream
                self.side = _side_stream(dev, self.main, self.kind)
                for t in self.inputs:
                    t.record_stream(self.side)
            self.side.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False

    def join(self, *outputs, on_current=False):
        """Make the stream that was current at the fork (or, on_current, the stream current now) wait for the branch;
        -> the outputs (one, or a tuple)."""
        if self.on:
            target = torch.cuda.current_stream(self.inputs[0].device) if on_current else self.main
            target.wait_stream(self.side)
            for t in outputs:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(target)
        return outputs[0] if len(outputs) == 1 else outputs


# Branches whose results nobody needs until the end of the backward pass (weight gradients): a step engine that calls
# join_pending() after loss.backward() sets DEFER_JOINS, and such branches then stay un-joined until that call.
DEFER_JOINS = False
_pending = []


def defer_or_join(fork, *outputs):
    if DEFER_JOINS and fork.on:
        _pending.append((fork, outputs))
    else:
        fork.join(*outputs)


def join_pending():
    """The current stream waits for every deferred branch."""
    while _pending:
        fork, outputs = _pending.pop()
        fork.join(*outputs, on_current=True)
