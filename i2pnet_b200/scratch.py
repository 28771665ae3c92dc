"""Zero-initialised work space for one training step, cleared by ONE memset.

The backward pass needs many small zero-filled buffers -- the f64 batch-norm sums of every shared-MLP chain and every
image-pyramid block, the accumulators of the cost-volume and gather backward kernels -- and `torch.zeros` is a fill
launch each: ~70 launches of 1-2 us on the serial chain of a 6 ms step.  Under a step engine they are carved out of one
arena instead, which `begin_step()` clears with a single memset.  Outside `begin_step() .. end_step()` (plain module
use, tests of single operators) `zeros()` is `torch.zeros`.

The arena's tensors are valid until the next `begin_step()`: they may carry gradients through the backward pass of the
step, never state that outlives it.  Capacity follows demand: a request the arena cannot serve falls back to
`torch.zeros` and is remembered, and the next `begin_step()` outside a graph capture grows the arena (two eager warm-up
steps before a capture are enough).
"""
import torch

_ALIGN = 256
_arenas = {}
_active = None


class _Arena:
    def __init__(self, device):
        self.device, self.buf, self.cap, self.used, self.need, self.dirty = device, None, 0, 0, 0, 0

    def begin(self):
        want = max(self.need, self.used)
        if want > self.cap and not torch.cuda.is_current_stream_capturing():
            self.cap = int(want * 1.25) + _ALIGN
            self.buf = torch.zeros(self.cap, dtype=torch.uint8, device=self.device)
            self.dirty = 0
        if self.buf is not None and self.dirty:
            self.buf[:self.dirty].zero_()
        self.used = self.need = self.dirty = 0

    def take(self, nbytes):
        n = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.need += n
        if self.buf is None or self.used + n > self.cap:
            return None
        out = self.buf[self.used:self.used + nbytes]
        self.used += n
        self.dirty = self.used
        return out


def begin_step(device):
    """Clear the device's arena (one memset over what the previous step used) and route zeros() to it."""
    global _active
    device = torch.device(device)
    if device.type != "cuda":
        return
    a = _arenas.get(device.index)
    if a is None:
        a = _arenas[device.index] = _Arena(device)
    a.begin()
    _active = a


def end_step():
    global _active
    _active = None


def zeros(shape, dtype, device):
    """torch.zeros(shape, dtype, device), from the step's arena when one is open on that device."""
    device = torch.device(device)
    a = _active
    if a is not None and device.type == "cuda" and device.index == a.device.index:
        numel = 1
        for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
            numel *= int(s)
        raw = a.take(numel * torch.empty((), dtype=dtype).element_size())
        if raw is not None:
            return raw.view(dtype).view(shape)
    return torch.zeros(shape, dtype=dtype, device=device)
