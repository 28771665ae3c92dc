"""Fused shared MLP: a chain of [1x1 conv -> batch-statistics BatchNorm -> (Leaky)ReLU] layers over a
channels-last (rows, C) tensor, optionally max-reduced over groups of K consecutive rows, as ONE
autograd.Function on the kernels of i2pnet_b200/csrc/mlp.cu.

The reference runs each layer as permute / nn.Conv2d / BatchNorm2d / activation / permute
(src/projectPN/PPBackbone_center.py:35-46) and the reduction as torch.max(dim=2) (:127, :198, :284);
autograd then replays every one of those passes backwards.  Here only the raw layer outputs y_l
touch HBM: one GEMM launch + one tiny finalize per layer forward, two GEMM launches per layer
backward (see mlp.cu for the algebra).  Numerically it is the same f32 computation with exact
(merged Welford) batch statistics; the bias feeding a BatchNorm gets an exactly zero gradient
instead of autograd's rounding noise.
"""
import os
import weakref

import torch
from torch.autograd import Function

from .. import _cabi
from .._cabi import _ptr, call, f32, i32

f64 = torch.float64


def _p(t):
    return None if t is None else t.data_ptr()


class _PackRegistry:
    """Packed (tf32 hi / lo, UMMA layout) copies of the weights of the tensor-core kernels under a step engine.  The weights
    change once per optimiser step, so the engine packs every registered layer in ONE launch per kernel family at the
    start of the step (`packs_begin_step`) instead of one launch per layer inside the forward / backward chains; layers are
    registered the first time a pass meets them (the eager warm-up steps).  Without an engine every pass packs its own
    copy, as before.  kind "mlp": i2p_pw_pack_weights (both orientations in one buffer); kind "conv": i2p_conv3x3_pack,
    variant = dgrad."""

    MULTI = {"mlp": ("i2p_pw_pack_weights_multi", lambda e, key: [e["cin"] | (e["cout"] << 32), e["ptr"], e["pack"].data_ptr()]),
             "conv": ("i2p_conv3x3_pack_multi", lambda e, key: [e["cin"] | (e["cout"] << 32), key[2], e["ptr"], e["pack"].data_ptr()])}

    def __init__(self, device):
        self.device, self.entries, self.tables, self.dirty, self.active, self.packed = device, {}, {}, False, False, False

    def begin_step(self):
        self.active, self.packed = True, False
        dead = [k for k, e in self.entries.items() if e["ref"]() is None]      # layers of models that are gone
        for k in dead:
            del self.entries[k]
            self.dirty = True
        if not self.entries:
            self.tables = {}
            return
        if (not self.tables or self.dirty) and not torch.cuda.is_current_stream_capturing():
            rows = {}
            for key, e in self.entries.items():
                e["ptr"] = e["ref"]().data_ptr()
                rows.setdefault(key[0], []).append(self.MULTI[key[0]][1](e, key))
                e["in_table"] = True
            self.tables = {kind: (len(r), torch.tensor(r, dtype=torch.int64).to(self.device)) for kind, r in rows.items()}
            self.dirty = False
        if self.tables and not self.dirty:
            for kind, (n, table) in self.tables.items():
                call(self.MULTI[kind][0], self.device, n, table.data_ptr())
            self.packed = True

    def get(self, kind, w, variant, cin, cout, floats):
        """-> (pack tensor, already packed this step)"""
        key = (kind, id(w), variant)
        e = self.entries.get(key)
        if e is not None and e["ref"]() is not w:       # the id of a dead parameter, reused
            e = None
        if e is None:
            pack = torch.empty(floats, dtype=f32, device=self.device)
            e = self.entries[key] = dict(ref=weakref.ref(w), cin=cin, cout=cout, pack=pack, ptr=None, in_table=False)
            self.dirty = True
        ready = self.packed and e["in_table"] and e["ptr"] == w.data_ptr()
        if not ready and e["in_table"] and e["ptr"] != w.data_ptr():
            self.dirty = True              # the parameter was re-seated: rebuild the table at the next step
        return e["pack"], ready


_REGISTRIES = {}
_PACK_ONCE = os.environ.get("I2P_PACK_ONCE", "mlp,conv").split(",")      # A/B switch: which kernel families pack once per step


def packs_begin_step(device):
    device = torch.device(device)
    if device.type != "cuda":
        return
    reg = _REGISTRIES.get(device.index)
    if reg is None:
        reg = _REGISTRIES[device.index] = _PackRegistry(device)
    reg.begin_step()


def packs_end_step():
    for reg in _REGISTRIES.values():
        reg.active = False


def _weight_pack(w, cin, cout, dev):
    floats = _cabi.lib().i2p_pw_pack_floats(cin, cout)
    reg = _REGISTRIES.get(dev.index)
    if reg is not None and reg.active and "mlp" in _PACK_ONCE:
        pack, ready = reg.get("mlp", w, 0, cin, cout, floats)
    else:
        pack, ready = torch.empty(floats, dtype=f32, device=dev), False
    if not ready:
        call("i2p_pw_pack_weights", dev, cin, cout, _ptr(w, f32, "weight", dev), pack.data_ptr())
    return pack


def conv_weight_pack(w, cin, cout, dgrad, dev):
    """Packed operand of the 3x3 convolution kernels (csrc/conv.cu) for weight w (cout, cin, 3, 3)."""
    floats = _cabi.lib().i2p_conv3x3_pack_floats(cin, cout, int(dgrad))
    reg = _REGISTRIES.get(dev.index)
    if reg is not None and reg.active and "conv" in _PACK_ONCE:
        pack, ready = reg.get("conv", w, int(dgrad), cin, cout, floats)
    else:
        pack, ready = torch.empty(floats, dtype=f32, device=dev), False
    if not ready:
        call("i2p_conv3x3_pack", dev, cin, cout, int(dgrad), _ptr(w, f32, "weight", dev), pack.data_ptr())
    return pack


class FusedMLPFunction(Function):
    @staticmethod
    def forward(ctx, x, reduce_k, slopes, eps, trackers, *params):
        """x (rows, cin) f32 contiguous; params = (w_1 (c1,cin), b_1, gamma_1, beta_1, w_2, ...);
        reduce_k: 0 -> out (rows, c_L); K > 0 -> out (rows / K, c_L) = max over each K consecutive rows.
        trackers[l]: the BatchNorm module whose running statistics layer l updates (training mode of a
        track_running_stats norm, as nn.BatchNorm2d does: momentum blend with the unbiased batch variance), or None."""
        dev = x.device
        rows = x.shape[0]
        L = len(params) // 4
        lib = _cabi.lib()
        ntiles = lib.i2p_pw_num_tiles(rows)
        tc_mask = lib.i2p_get_mlp_tensor_cores()
        ys, stats, packs = [], [], []
        inp, in_stats, in_slope = x, None, 1.0
        for l in range(L):
            w, b, gamma, beta = params[4 * l:4 * l + 4]
            cout, cin = w.shape[0], w.shape[1]       # (cout, cin, 1, 1): the nn.Conv2d parameter itself
            y = torch.empty(rows, cout, dtype=f32, device=dev)
            tiles = torch.empty(ntiles, cout, 2, dtype=f32, device=dev)
            fwd_tc = bool(tc_mask & 1) and lib.i2p_pw_tc_supported(0, rows, cin, cout)
            dx_tc = bool(tc_mask & 2) and lib.i2p_pw_tc_supported(1, rows, cin, cout)
            pack = None
            if fwd_tc or dx_tc:   # tf32 (hi, lo) halves of W in the UMMA layout, both orientations, once per step
                pack = _weight_pack(w, cin, cout, dev)
            scale_p = _p(in_stats[2]) if in_stats is not None else None
            shift_p = _p(in_stats[3]) if in_stats is not None else None
            if fwd_tc:
                call("i2p_pw_linear_fwd_tc", dev, rows, cin, cout, _ptr(inp, f32, "x", dev), scale_p, shift_p, float(in_slope),
                     pack.data_ptr(), _ptr(b, f32, "bias", dev), y.data_ptr(), tiles.data_ptr())
            else:
                call("i2p_pw_linear_fwd", dev, rows, cin, cout, _ptr(inp, f32, "x", dev), scale_p, shift_p, float(in_slope),
                     _ptr(w, f32, "weight", dev), _ptr(b, f32, "bias", dev), y.data_ptr(), tiles.data_ptr())
            st = torch.empty(4, cout, dtype=f32, device=dev)  # mean, rstd, scale, shift
            call("i2p_bn_finalize", dev, rows, cout, tiles.data_ptr(), _ptr(gamma, f32, "gamma", dev),
                 _ptr(beta, f32, "beta", dev), float(eps[l]), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(),
                 st[3].data_ptr())
            bnm = trackers[l]
            if bnm is not None:
                bnm.num_batches_tracked.add_(1)
                var = (1.0 / (st[1] * st[1]) - float(eps[l])).clamp_(min=0.0)       # biased batch variance back from rstd
                unbias = rows / max(rows - 1, 1)
                if bnm.momentum is None:    # nn.BatchNorm: cumulative moving average, factor 1 / num_batches_tracked
                    f = 1.0 / bnm.num_batches_tracked.to(f32)
                    bnm.running_mean.add_((st[0] - bnm.running_mean) * f)
                    bnm.running_var.add_((var * unbias - bnm.running_var) * f)
                else:
                    mom = bnm.momentum
                    bnm.running_mean.mul_(1.0 - mom).add_(st[0], alpha=mom)
                    bnm.running_var.mul_(1.0 - mom).add_(var, alpha=mom * unbias)
            ys.append(y)
            stats.append(st)
            packs.append(pack)
            inp, in_stats, in_slope = y, st, slopes[l]
        c_last = ys[-1].shape[1]
        st = stats[-1]
        arg = None
        if reduce_k:
            groups = rows // reduce_k
            out = torch.empty(groups, c_last, dtype=f32, device=dev)
            arg = torch.empty(groups, c_last, dtype=i32, device=dev)
            call("i2p_bn_act_maxk", dev, groups, reduce_k, c_last, ys[-1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(),
                 float(slopes[-1]), out.data_ptr(), arg.data_ptr())
        else:
            out = torch.empty(rows, c_last, dtype=f32, device=dev)
            call("i2p_bn_act", dev, rows, c_last, ys[-1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), float(slopes[-1]),
                 out.data_ptr())
        ctx.save_for_backward(x, *params, *ys, *stats, *([arg] if arg is not None else []))
        ctx.meta = (L, reduce_k, tuple(slopes))
        ctx.weights = [params[4 * l] for l in range(L)]      # the Parameter objects (their gradient sinks, engine.grad_sink)
        ctx.biases = [params[4 * l + 1] for l in range(L)]
        ctx.packs = packs     # the weights do not change between forward and backward
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L, reduce_k, slopes = ctx.meta
        saved = ctx.saved_tensors
        x, params = saved[0], saved[1:1 + 4 * L]
        ys, stats = saved[1 + 4 * L:1 + 5 * L], saved[1 + 5 * L:1 + 6 * L]
        arg = saved[1 + 6 * L] if reduce_k else None
        dev = x.device
        rows = x.shape[0]
        grad_out = grad_out.contiguous()
        # one zero-filled f64 buffer for every layer's batch-norm sums and one f32 buffer for every dW and
        # (identically zero) bias gradient: two fill launches per chain instead of three per layer
        couts = [ys[l].shape[1] for l in range(L)]
        from .. import scratch
        from ..engine import grad_sink
        s12_all = scratch.zeros(2 * sum(couts), f64, dev)
        # weights / biases with a gradient sink (step engine) need no buffer of their own: dW accumulates into the sink, the
        # bias gradient under a batch-norm is identically zero and the sink already holds zero
        w_sinks = [grad_sink(ctx.weights[l]) for l in range(L)]
        b_sunk = [grad_sink(ctx.biases[l]) is not None for l in range(L)]
        wsizes = [0 if w_sinks[l] is not None else params[4 * l].numel() for l in range(L)]
        bsizes = [0 if b_sunk[l] else couts[l] for l in range(L)]
        wz = scratch.zeros(sum(wsizes) + sum(bsizes), f32, dev) if sum(wsizes) + sum(bsizes) else None
        s12, dws, dbs = [], [], []
        o12 = ow = 0
        for l in range(L):
            s12.append(s12_all[o12:o12 + 2 * couts[l]].view(2, couts[l]))
            o12 += 2 * couts[l]
            dws.append(wz[ow:ow + wsizes[l]].view(params[4 * l].shape) if wsizes[l] else None)
            ow += wsizes[l]
        for l in range(L):
            dbs.append(wz[ow:ow + bsizes[l]] if bsizes[l] else None)
            ow += bsizes[l]

        def src(l, g):  # (g_dense, dout, arg, k) of layer l
            if l == L - 1 and reduce_k:
                return None, grad_out.data_ptr(), arg.data_ptr(), reduce_k
            return g.data_ptr(), None, None, 1

        def bn(l):
            st = stats[l]
            return (ys[l].data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(),
                    float(slopes[l]))

        lib = _cabi.lib()
        tc_mask = lib.i2p_get_mlp_tensor_cores()
        packs = ctx.packs
        if grad_out.data_ptr() % 16:
            grad_out = grad_out.clone()        # the tensor-core kernels read gradients with 128-bit loads
        g = None if reduce_k else grad_out
        c_last = ys[-1].shape[1]
        call("i2p_bn_bwd_reduce", dev, rows, c_last, *src(L - 1, g), *bn(L - 1), s12[L - 1].data_ptr())
        grads = [None] * (4 * L)
        dx = None
        from .. import streams as _streams
        for l in range(L - 1, -1, -1):
            w = params[4 * l]
            cout, cin = w.shape[0], w.shape[1]
            inp = ys[l - 1] if l > 0 else x
            pst = stats[l - 1] if l > 0 else None
            pscale = _p(pst[2]) if pst is not None else None
            pshift = _p(pst[3]) if pst is not None else None
            pslope = float(slopes[l - 1]) if l > 0 else 1.0
            on_tc = (tc_mask & 32) or not (l == L - 1 and reduce_k)      # bit 32: max-over-K sources on the tensor cores too
            # The weight gradient is needed by the optimiser alone, the data gradient by the rest of the backward pass:
            # dW goes to a side stream (it reads what dX reads), and under a step engine it accumulates straight into
            # the flat gradient buffer and stays un-joined until the whole backward has been issued (streams.defer_or_join).
            sink = w_sinks[l]
            dw_target = sink if sink is not None else dws[l]
            with _streams.Fork(grad_out, g if g is not None else grad_out, ys[l], stats[l], s12_all, inp,
                               *([pst] if pst is not None else []), *([arg] if arg is not None else []), kind="wgrad") as branch:
                if on_tc and (tc_mask & 4) and lib.i2p_pw_tc_supported(2, rows, cin, cout):
                    call("i2p_pw_linear_bwd_dw_tc", dev, rows, cin, cout, *src(l, g), *bn(l), s12[l].data_ptr(), inp.data_ptr(),
                         pscale, pshift, pslope, dw_target.data_ptr())
                else:
                    call("i2p_pw_linear_bwd_dw", dev, rows, cin, cout, *src(l, g), *bn(l), s12[l].data_ptr(), inp.data_ptr(),
                         pscale, pshift, pslope, dw_target.data_ptr())
            if sink is not None:
                _streams.defer_or_join(branch)
            else:
                branch.join(dws[l])
            grads[4 * l] = None if sink is not None else dws[l]
            grads[4 * l + 1] = dbs[l]                                       # bias under BN: exactly zero
            if l > 0 or ctx.needs_input_grad[0]:
                dx = torch.empty(rows, cin, dtype=f32, device=dev)
                prev = bn(l - 1) if l > 0 else (None, None, None, None, None, 1.0)
                prev_s12 = s12[l - 1].data_ptr() if l > 0 else None
                if on_tc and (tc_mask & 2) and packs[l] is not None and lib.i2p_pw_tc_supported(1, rows, cin, cout):
                    call("i2p_pw_linear_bwd_dx_tc", dev, rows, cin, cout, *src(l, g), *bn(l), s12[l].data_ptr(),
                         packs[l].data_ptr(), dx.data_ptr(), *prev, prev_s12)
                else:
                    call("i2p_pw_linear_bwd_dx", dev, rows, cin, cout, *src(l, g), *bn(l), s12[l].data_ptr(), w.data_ptr(),
                         dx.data_ptr(), *prev, prev_s12)
                g = dx
        gb = s12_all.to(f32)          # every layer's (d beta = sum dz, d gamma = sum dz * yhat), one conversion
        o12 = 0
        for l in range(L):
            grads[4 * l + 3] = gb[o12:o12 + couts[l]]
            grads[4 * l + 2] = gb[o12 + couts[l]:o12 + 2 * couts[l]]
            o12 += 2 * couts[l]
        return (dx if ctx.needs_input_grad[0] else None, None, None, None, None, *grads)


def fusable(convs):
    """The fused path covers what the models in the reference use: an affine BatchNorm on batch statistics after
    every 1x1 conv -- the large-range model's use_bn_input norms (never tracking), and the small-range model's
    tracking norms while training (their running statistics are updated from the kernels' batch statistics).
    A tracking norm in eval mode normalises with its running statistics: layer-by-layer path."""
    return all(c.bn and c.bn_linear.affine and (c.bn_linear.training or not c.bn_linear.track_running_stats)
               and c.out_channels % 16 == 0 and c.out_channels <= 512 for c in convs)


def fused_mlp(x, convs, reduce_k=False):
    """x (..., [K,] cin) -> (..., [K,] c_L), or (..., c_L) with the max over the K axis when reduce_k."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    params, slopes, eps, trackers = [], [], [], []
    for c in convs:
        trackers.append(c.bn_linear if c.bn_linear.track_running_stats and c.bn_linear.running_mean is not None else None)
        params += [c.conv.weight, c.conv.bias, c.bn_linear.weight, c.bn_linear.bias]
        slopes.append(1.0 if not c.activation_fn else (0.1 if c.leaky_relu else 0.0))
        eps.append(c.bn_linear.eps)
    k = int(lead[-1]) if reduce_k else 0
    out = FusedMLPFunction.apply(x2, k, tuple(slopes), tuple(eps), tuple(trackers), *params)
    return out.view(*(lead[:-1] if reduce_k else lead), out.shape[-1])
