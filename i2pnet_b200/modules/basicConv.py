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

USE_FUSED_RGB_TAIL = True   # False: nn.BatchNorm2d -> nn.LeakyReLU -> nn.MaxPool2d through ATen / cuDNN (A/B, with USE_OWN_CONV off)
USE_OWN_CONV = os.environ.get("I2P_OWN_CONV", "1") != "0"   # "0": library (cuDNN) convolutions, for A/B measurements only


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


class _ConvBlock(Function):
    """One pyramid block -- Conv2d(3x3, pad 1) -> BatchNorm2d -> LeakyReLU -> MaxPool2d(3, stride, 1) -- entirely on own
    kernels (csrc/conv.cu, csrc/rgb.cu):
      forward : weight pack, implicit-GEMM convolution on the tensor cores with the batch statistics of its output in
                the epilogue, statistics merge, pooled write;
      backward: pooling / activation / batch-norm backward -> dy; the data gradient is the same tensor-core kernel on
                the flipped weights; the weight gradient runs on a side stream (only the optimiser needs it) and
                accumulates straight into the step engine's flat gradient buffer when there is one (engine.grad_sink)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, bn, slope, stride):
        dev = x.device
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        L = _cabi.lib()
        xp = _ptr(x, f32, "conv input", dev)
        from ..projectPN.fused_mlp import conv_weight_pack
        pack = conv_weight_pack(w, Cin, Cout, 0, dev)       # once per step for all layers under a step engine
        batch_stats = bn.training or not bn.track_running_stats
        ntiles = L.i2p_conv3x3_stat_slots(B, Cout, H, W)     # one statistics slot per persistent CTA
        y = torch.empty(B, Cout, H, W, dtype=f32, device=dev)
        tiles = torch.empty(Cout, ntiles, 3, dtype=f32, device=dev) if batch_stats else None
        call("i2p_conv3x3_tc", dev, B, Cin, Cout, H, W, xp, pack.data_ptr(), _ptr(b, f32, "conv bias", dev) if b is not None else None,
             y.data_ptr(), tiles.data_ptr() if batch_stats else None)
        stats = torch.empty(4, Cout, dtype=f32, device=dev)
        s12 = torch.empty(L.i2p_rgb_s12_slots(), Cout, dtype=torch.float64, device=dev)
        if batch_stats:
            track = bn.track_running_stats and bn.running_mean is not None
            if track and bn.momentum is None:
                raise _cabi.I2PError("BatchNorm2d(momentum=None) is not covered by the fused RGB block")
            call("i2p_rgb_bn_finalize", dev, Cout, ntiles, tiles.data_ptr(), _ptr(gamma, f32, "bn weight", dev),
                 _ptr(beta, f32, "bn bias", dev), float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1),
                 bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                 bn.num_batches_tracked.data_ptr() if track else None, stats.data_ptr(), s12.data_ptr())
        else:
            call("i2p_rgb_bn_from_running", dev, Cout, _ptr(gamma, f32, "bn weight", dev), _ptr(beta, f32, "bn bias", dev),
                 float(bn.eps), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), stats.data_ptr(), s12.data_ptr())
        Ho, Wo = L.i2p_rgb_pool_out(H, stride), L.i2p_rgb_pool_out(W, stride)
        out = torch.empty(B, Cout, Ho, Wo, dtype=f32, device=dev)
        call("i2p_rgb_bn_act_pool_fwd", dev, B, Cout, H, W, stride, y.data_ptr(), stats.data_ptr(), float(slope), out.data_ptr())
        ctx.save_for_backward(x, w, y, stats, s12)
        ctx.meta = (stride, float(slope), bool(batch_stats), b is not None)
        ctx.params = (w, b)
        ctx.used = False
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..engine import grad_sink
        x, w, y, stats, s12 = ctx.saved_tensors
        stride, slope, batch_stats, has_bias = ctx.meta
        dev = y.device
        B, Cout, H, W = y.shape
        Cin = x.shape[1]
        L = _cabi.lib()
        if ctx.used:          # a second backward through the same graph: the sums start from zero again
            s12.zero_()
        ctx.used = True
        dout = dout.contiguous()
        dy = torch.empty_like(y)
        dgb = torch.empty(2, Cout, dtype=f32, device=dev)
        call("i2p_rgb_bn_act_pool_bwd", dev, B, Cout, H, W, stride, int(batch_stats), y.data_ptr(), stats.data_ptr(), slope,
             _ptr(dout, f32, "grad_output", dev), s12.data_ptr(), dy.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr())
        # weight gradient: nobody but the optimiser reads it -> side stream, joined at the end of the backward pass
        sink_w = grad_sink(ctx.params[0])
        gw = gb = None
        with _streams.Fork(dy, x, kind="cwgrad") as branch:
            if sink_w is None:
                gw = torch.zeros_like(w)
            call("i2p_conv3x3_wgrad", dev, B, Cin, Cout, H, W, x.data_ptr(), dy.data_ptr(),
                 (sink_w if sink_w is not None else gw).data_ptr())
        if sink_w is not None:
            _streams.defer_or_join(branch)      # no tensor of the branch is handed to autograd: the sink received it
        else:
            branch.join(gw)
            if has_bias:                          # a bias in front of a batch-norm: sum(dy) over (B, H, W), ~0 by construction
                gb = dy.sum(dim=(0, 2, 3)) if not batch_stats else torch.zeros(Cout, dtype=f32, device=dev)
        if sink_w is not None and has_bias and not batch_stats:
            grad_sink(ctx.params[1]).add_(dy.sum(dim=(0, 2, 3)))
        gx = None
        if ctx.needs_input_grad[0]:
            from ..projectPN.fused_mlp import conv_weight_pack
            pack = conv_weight_pack(ctx.params[0], Cin, Cout, 1, dev)
            gx = torch.empty_like(x)
            call("i2p_conv3x3_tc", dev, B, Cout, Cin, H, W, dy.data_ptr(), pack.data_ptr(), None, gx.data_ptr(), None)
        return gx, gw, gb, dgb[0], dgb[1], None, None, None


def _conv_block_supported(x, conv, bn, act, pool):
    def one(v):
        return v if isinstance(v, int) else (v[0] if v[0] == v[1] else None)
    return (x.dtype == f32 and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros" and bn.affine
            and isinstance(act, nn.LeakyReLU) and one(pool.kernel_size) == 3 and one(pool.padding) == 1
            and one(pool.stride) in (1, 2) and one(pool.dilation) == 1 and not pool.ceil_mode
            and (bn.training or bn.track_running_stats))


def _fusable(y, bn, act, pool):
    def one(v):
        return v if isinstance(v, int) else (v[0] if v[0] == v[1] else None)
    return (USE_FUSED_RGB_TAIL and y.is_cuda and y.dtype == f32 and y.is_contiguous() and bn.affine
            and isinstance(act, nn.LeakyReLU) and one(pool.kernel_size) == 3 and one(pool.padding) == 1
            and one(pool.stride) in (1, 2) and one(pool.dilation) == 1 and not pool.ceil_mode
            and (bn.training or bn.track_running_stats))


class _Pyramid(nn.Sequential):
    """The Sequential the reference builds (same slots, same state_dict keys).  On CUDA every block runs as ONE fused
    operator on own kernels (_ConvBlock); a block configuration those kernels do not cover raises -- there is no
    library fallback on the device.  CPU tensors (host-logic tests only) take the reference's module-by-module form."""

    def forward(self, x):
        mods = list(self)
        for i in range(0, len(mods), 4):
            conv, bn, act, pool = mods[i:i + 4]
            if x.is_cuda and USE_OWN_CONV:
                if not _conv_block_supported(x, conv, bn, act, pool):
                    raise _cabi.I2PError("RGB pyramid block %d: configuration not covered by the sm_100a kernels" % (i // 4))
                stride = pool.stride if isinstance(pool.stride, int) else pool.stride[0]
                x = _ConvBlock.apply(x.contiguous(), conv.weight, conv.bias, bn.weight, bn.bias, bn, act.negative_slope, stride)
                continue
            y = conv(x)          # A/B measurements (USE_OWN_CONV = False: library convolution) and CPU host logic
            if _fusable(y, bn, act, pool):
                stride = pool.stride if isinstance(pool.stride, int) else pool.stride[0]
                x = _BlockTail.apply(y, bn.weight, bn.bias, bn, act.negative_slope, stride)
            else:
                x = pool(act(bn(y)))
        return x


def pyramid_out_hw(nets, H, W):
    """Spatial size of the output of a chain of createCNNs pyramids for an (H, W) input: the 3x3 / stride-1 / pad-1
    convolutions keep it, MaxPool2d(3, s, 1) gives floor((n - 1) / s) + 1.  (Lets the caller prepare what depends on the
    feature map's SHAPE only -- the pixel rays of the cost volumes -- before the pyramid has run.)"""
    for net in nets:
        for m in net:
            if isinstance(m, nn.MaxPool2d):
                s = m.stride if isinstance(m.stride, int) else m.stride[0]
                H, W = (H - 1) // s + 1, (W - 1) // s + 1
    return H, W


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
