"""RGB feature pyramid and small conv wrappers (mirror of src/modules/basicConv.py:
createCNNs :6-20, Conv1d :63-85).  Module indices inside the Sequential -- hence the
state_dict keys `RGB_net1.{0,1,4,5,...}` -- follow the reference: block i owns slots
4i (conv), 4i+1 (BatchNorm2d), 4i+2 (LeakyReLU 0.1), 4i+3 (MaxPool 3x3)."""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _cabi
from .. import streams as _streams
from .._cabi import _ptr, call, f32

USE_FUSED_RGB_TAIL = True   # False: nn.BatchNorm2d -> nn.LeakyReLU -> nn.MaxPool2d through ATen / cuDNN
SPLIT_CONV_BACKWARD = os.environ.get("I2P_SPLIT_WGRAD", "1") != "0"  # weight gradients of the 3x3 convolutions on a side stream (see _ConvSplitBackward)


class _BlockTail(Function):
    """BatchNorm2d -> LeakyReLU -> MaxPool2d(3, stride, 1) of one pyramid block on the kernels of
    csrc/rgb.cu: one statistics pass + one pooled write forward, one reduction + one write backward,
    nothing but the convolution output y kept for backward (the pooling arg-max is re-derived from it)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, bn, slope, stride):
        dev = y.device
        B, C, H, W = y.shape
        L = _cabi.lib()
        Ho, Wo = L.i2p_rgb_pool_out(H, stride), L.i2p_rgb_pool_out(W, stride)
        stats = torch.empty(4, C, dtype=f32, device=dev)
        s12 = torch.empty(L.i2p_rgb_s12_slots(), C, dtype=torch.float64, device=dev)
        batch_stats = bn.training or not bn.track_running_stats
        yp = _ptr(y, f32, "conv output", dev)
        if batch_stats:
            ntiles = B * L.i2p_rgb_num_chunks(H * W)
            tiles = torch.empty(C, ntiles, 3, dtype=f32, device=dev)
            call("i2p_rgb_bn_stats", dev, B, C, H, W, yp, tiles.data_ptr())
            track = bn.track_running_stats and bn.running_mean is not None
            call("i2p_rgb_bn_finalize", dev, C, ntiles, tiles.data_ptr(), _ptr(gamma, f32, "bn weight", dev),
                 _ptr(beta, f32, "bn bias", dev), float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1),
                 bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                 bn.num_batches_tracked.data_ptr() if track else None, stats.data_ptr(), s12.data_ptr())
        else:
            call("i2p_rgb_bn_from_running", dev, C, _ptr(gamma, f32, "bn weight", dev), _ptr(beta, f32, "bn bias", dev),
                 float(bn.eps), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), stats.data_ptr(), s12.data_ptr())
        out = torch.empty(B, C, Ho, Wo, dtype=f32, device=dev)
        call("i2p_rgb_bn_act_pool_fwd", dev, B, C, H, W, stride, yp, stats.data_ptr(), float(slope), out.data_ptr())
        ctx.save_for_backward(y, stats, s12)
        ctx.meta = (stride, float(slope), bool(batch_stats))
        ctx.used = False
        return out

    @staticmethod
    def backward(ctx, dout):
        y, stats, s12 = ctx.saved_tensors
        stride, slope, batch_stats = ctx.meta
        dev = y.device
        B, C, H, W = y.shape
        if ctx.used:          # a second backward through the same graph: the sums start from zero again
            s12.zero_()
        ctx.used = True
        dout = dout.contiguous()
        dy = torch.empty_like(y)
        dgb = torch.empty(2, C, dtype=f32, device=dev)
        call("i2p_rgb_bn_act_pool_bwd", dev, B, C, H, W, stride, int(batch_stats), y.data_ptr(), stats.data_ptr(), slope,
             _ptr(dout, f32, "grad_output", dev), s12.data_ptr(), dy.data_ptr(), dgb[0].data_ptr(),
             dgb[1].data_ptr())
        return dy, dgb[0], dgb[1], None, None, None


class _ConvSplitBackward(Function):
    """nn.Conv2d (library convolution, cuDNN) whose backward issues the weight gradient on a side stream: only the data
    gradient is on the dependent chain of the image branch, the weight gradients are needed by the optimiser alone
    (streams.defer_or_join: they stay un-joined until the end of the backward pass when a step engine asks for that)."""

    @staticmethod
    def forward(ctx, x, w, b, conf):
        ctx.save_for_backward(x, w)
        ctx.conf, ctx.has_bias = conf, b is not None
        ctx.params = (w, b)
        stride, padding, dilation, groups = conf
        return torch.nn.functional.conv2d(x, w, b, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        bias_sizes = [w.shape[0]] if ctx.has_bias else None
        back = torch.ops.aten.convolution_backward
        from ..engine import grad_sink
        sink_w = grad_sink(ctx.params[0])
        sink_b = grad_sink(ctx.params[1]) if ctx.has_bias else None
        with _streams.Fork(gy, x, w) as branch:
            _, gw, gb = back(gy, x, w, bias_sizes, stride, padding, dilation, False, [0, 0], groups,
                             [False, True, ctx.has_bias])
            if sink_w is not None:
                # A tensor produced on the side stream must not be handed to autograd: AccumulateGrad runs on THIS node's
                # stream and would read it without waiting for the branch.  The gradients go straight into the step
                # engine's flat buffer on the branch's own stream; the engine joins the branch before it reads the buffer.
                sink_w.add_(gw)
                if sink_b is not None:
                    sink_b.add_(gb)
                gw = gb = None
        gx = None
        if ctx.needs_input_grad[0]:
            gx = back(gy, x, w, bias_sizes, stride, padding, dilation, False, [0, 0], groups, [True, False, False])[0]
        if sink_w is not None:
            _streams.defer_or_join(branch)
        else:
            branch.join(gw, gb)          # no engine: autograd owns the gradients, so they are joined before it sees them
        return gx, gw, gb, None


def _conv(conv, x):
    if (SPLIT_CONV_BACKWARD and x.is_cuda and torch.is_grad_enabled() and conv.padding_mode == "zeros"
            and not isinstance(conv.padding, str)):
        return _ConvSplitBackward.apply(x, conv.weight, conv.bias, (list(conv.stride), list(conv.padding), list(conv.dilation),
                                                                    conv.groups))
    return conv(x)


def _fusable(y, bn, act, pool):
    def one(v):
        return v if isinstance(v, int) else (v[0] if v[0] == v[1] else None)
    return (USE_FUSED_RGB_TAIL and y.is_cuda and y.dtype == f32 and y.is_contiguous() and bn.affine
            and isinstance(act, nn.LeakyReLU) and one(pool.kernel_size) == 3 and one(pool.padding) == 1
            and one(pool.stride) in (1, 2) and one(pool.dilation) == 1 and not pool.ceil_mode
            and (bn.training or bn.track_running_stats))


class _Pyramid(nn.Sequential):
    """The Sequential the reference builds (same slots, same state_dict keys); on CUDA the three
    element-wise slots of every block run as one fused operator behind the library convolution."""

    def forward(self, x):
        mods = list(self)
        for i in range(0, len(mods), 4):
            conv, bn, act, pool = mods[i:i + 4]
            y = _conv(conv, x)
            if _fusable(y, bn, act, pool):
                stride = pool.stride if isinstance(pool.stride, int) else pool.stride[0]
                x = _BlockTail.apply(y, bn.weight, bn.bias, bn, act.negative_slope, stride)
            else:
                x = pool(act(bn(y)))
        return x


def createCNNs(in_channel, channels, strides):
    layers = _Pyramid()
    last = in_channel
    for i, (out_channel, stride) in enumerate(zip(channels, strides)):
        layers.add_module(str(4 * i), nn.Conv2d(last, out_channel, kernel_size=3, stride=1, padding=1, bias=True))
        layers.add_module(str(4 * i + 1), nn.BatchNorm2d(out_channel))
        layers.add_module(str(4 * i + 2), nn.LeakyReLU(negative_slope=0.1))
        layers.add_module(str(4 * i + 3), nn.MaxPool2d(3, stride=stride, padding=1))
        last = out_channel
    return layers


class Conv1d(nn.Module):
    """(B,N,C) -> (B,N,C'): Conv1d(k=1) [+ BatchNorm1d] [+ (Leaky)ReLU] under `composed_module`."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, use_activation=True,
                 use_leaky=True, bn=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        if use_activation:
            act = nn.LeakyReLU(0.1, inplace=True) if use_leaky else nn.ReLU(inplace=True)
        else:
            act = nn.Identity()
        self.composed_module = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.BatchNorm1d(out_channels) if bn else nn.Identity(),
            act)

    def forward(self, x):
        return self.composed_module(x.permute(0, 2, 1)).permute(0, 2, 1)
