"""Projection-aware building blocks of the large-range model (mirror of the reference's
src/projectPN/PPBackbone_center.py: Conv2d :10, ProjectPointNet :54, ProjSetUpconvModule :208,
CostVolume :305, PoseHead :503, FlowPredictor :566).

Constructor arguments, forward signatures, return tuples and parameter names (hence
state_dict keys: `mlp_convs.0.conv.weight`, `...bn_linear.bias`, `hidden_layer.
composed_module.0.weight`, ...) are the reference's, so checkpoints and the unchanged model
assembly interchange.  The data flow underneath is re-designed:

* tensors stay channels-last (B, N, K, C) end to end: the shared MLP is a row-major GEMM
  + batch-statistics normalisation on that layout, not the reference's permute -> NCHW 1x1
  conv -> BatchNorm2d -> permute sandwich (:35-46);
* neighbourhoods come from the warp-per-centre select kernel as one int32 flat index (no
  int64 (b,h,w) triples, no zero-filled `valid_idx` buffers) and are fetched with the
  vectorised row gather, whose backward is a vector red.add scatter;
* centres of a regular stride grid are a strided view, not a gather;
* the cost volume never materialises the reference's `repeat`ed (B,N,K,C) operands.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..modules.basicConv import Conv1d
from . import fused_mlp as _fm
from ..streams import Fork
from .utils import (FLAG_COPY, FLAG_SHIFT, StrideGrid, check_valid, gather_rows, gather_torch, knn_point,
                    select_flat)


USE_FUSED_MLP = True   # False: the layer-by-layer ATen formulation (kept for A/B parity tests)


LIBRARY_FALLBACKS = 0   # how often a CUDA tensor went through the layer-by-layer ATen / cuBLAS formulation (see run_mlp)


def run_mlp(convs, x, reduce_k=False):
    """Apply a chain of Conv2d blocks to channels-last x (b, n, k, c); reduce_k: max over the k axis.
    CUDA tensors take the fused kernels (fused_mlp.py).  The layer-by-layer formulation is the CPU host-logic path and
    the explicit A/B switch (USE_FUSED_MLP = False).  A CUDA tensor whose layer configuration the fused kernels do not
    cover (only: a running-statistics BatchNorm in eval mode, i.e. the small-range model at inference) is COUNTED in
    LIBRARY_FALLBACKS and warned about once -- the large-range model never gets there (tests/test_model_gpu.py asserts
    the counter stays 0) -- and raises under I2P_STRICT=1."""
    global LIBRARY_FALLBACKS
    convs = list(convs)
    if USE_FUSED_MLP and x.is_cuda and len(convs) > 0:
        if _fm.fusable(convs):
            return _fm.fused_mlp(x, convs, reduce_k)
        LIBRARY_FALLBACKS += 1
        import os
        import warnings
        if os.environ.get("I2P_STRICT", "0") == "1":
            from .. import _cabi
            raise _cabi.I2PError("shared MLP configuration not covered by the sm_100a kernels (I2P_STRICT=1)")
        warnings.warn("i2pnet_b200: a shared-MLP chain ran through the ATen / cuBLAS formulation (layer configuration "
                      "outside the fused kernels' coverage)", RuntimeWarning, stacklevel=2)
    for conv in convs:
        x = conv._forward_layer(x)
    return torch.max(x, dim=2)[0] if reduce_k else x


class Conv2d(nn.Module):
    """1x1 conv (+ batch-statistics BN) (+ ReLU / LeakyReLU(0.1)) over the last axis of a
    channels-last (b, n, s, c) tensor.  `conv` / `bn_linear` hold the parameters under the
    reference's names; with use_bn_input the norm always uses the statistics of the current
    batch, in eval() too (track_running_stats=False, PPBackbone_center.py:30)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=None, bn=False, activation_fn=True,
                 leaky_relu=True, use_bn_input=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride if stride is not None else [1, 1]
        self.bn, self.activation_fn, self.use_bn_input, self.leaky_relu = bn, activation_fn, use_bn_input, leaky_relu
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, self.stride)
        if bn:
            self.bn_linear = nn.BatchNorm2d(out_channels, track_running_stats=not use_bn_input)
        if activation_fn:
            self.relu = nn.ReLU(inplace=True) if not leaky_relu else nn.LeakyReLU(0.1, inplace=True)

    def forward(self, x):
        return run_mlp([self], x)

    def _forward_layer(self, x):
        lead = x.shape[:-1]
        y = F.linear(x.reshape(-1, self.in_channels), self.conv.weight.view(self.out_channels, self.in_channels),
                     self.conv.bias)
        if self.bn:
            y = _batch_norm_rows(y, self.bn_linear)
        if self.activation_fn:
            y = self.relu(y)
        return y.view(*lead, self.out_channels)

    def set_bn(self):
        if self.bn:
            self.bn_linear.track_running_stats = not self.use_bn_input
            self.bn_linear.training = True


def _batch_norm_rows_stable(y, m):
    """BatchNorm over the rows of y (rows, C) with the parameters / buffers of the BatchNorm2d `m`.
    Batch statistics (biased variance, as nn.BatchNorm2d normalises) unless m tracks running
    statistics and is in eval mode.  Statistics come from torch.var_mean (a stable pairwise /
    Welford reduction on both devices); F.batch_norm's 2-D CPU path accumulates in f32 and is
    off by 1e-3 on 2e5 rows."""
    if m.training or not m.track_running_stats:
        var, mean = torch.var_mean(y, dim=0, unbiased=False)
        if m.track_running_stats and m.training:
            with torch.no_grad():
                mom = m.momentum if m.momentum is not None else 0.1
                n = y.shape[0]
                m.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                m.running_var.mul_(1 - mom).add_(var * (n / max(n - 1, 1)), alpha=mom)
                m.num_batches_tracked += 1
    else:
        mean, var = m.running_mean, m.running_var
    scale = torch.rsqrt(var + m.eps)
    if m.weight is not None:
        return (y - mean) * (scale * m.weight) + m.bias
    return (y - mean) * scale


def _batch_norm_rows(y, m):
    """The product path: ATen's fused channels-last batch-norm kernels (Welford statistics, one
    normalise pass, fused backward).  _batch_norm_rows_stable is the same function spelled with
    torch.var_mean; the CPU host-logic tests substitute it because ATen's 2-D CPU kernel
    accumulates in f32."""
    use_batch = m.training or not m.track_running_stats
    return F.batch_norm(y, m.running_mean if m.track_running_stats else None,
                        m.running_var if m.track_running_stats else None, m.weight, m.bias, use_batch,
                        m.momentum if m.momentum is not None else 0.1, m.eps)


def _centres(sample_idx, B, out_h, out_w, stride_H, stride_W, device):
    return sample_idx if sample_idx is not None else StrideGrid(B, out_h, out_w, stride_H, stride_W, device)


def _take_centres(image, sample_idx, B, H, W):
    if isinstance(sample_idx, StrideGrid):
        return sample_idx.take(image).contiguous()
    return gather_torch(image, *sample_idx, B, H, W)


def _centre_index(sample_idx, B, out_h, out_w):
    """What the select kernel needs: the grid itself, or (B,N,2) int32 for an explicit triple."""
    if isinstance(sample_idx, StrideGrid):
        return sample_idx
    _, h, w = sample_idx
    return torch.stack([h, w], dim=-1).reshape(B, out_h * out_w, 2).to(torch.int32)


class ProjectPointNet(nn.Module):
    """Set abstraction on the range image: stride-grid centres, K nearest inside a window,
    shared MLP, max over K."""

    def __init__(self, H, W, out_h, out_w, stride_H, stride_W, kernel_size, nsample, distance, in_channel, mlp,
                 use_trans=False, use_bn_p=True, use_bn_input=True):
        super().__init__()
        self.H, self.W, self.out_h, self.out_w = H, W, out_h, out_w
        self.stride_H, self.stride_W = stride_H, stride_W
        self.kernel_size, self.distance, self.nsample, self.usetrans = kernel_size, distance, nsample, use_trans
        self.mlp_convs = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(Conv2d(last_channel, out_channel, kernel_size=(1, 1), bn=use_bn_p,
                                         leaky_relu=False, use_bn_input=use_bn_input))
            last_channel = out_channel

    def _group(self, xyz_proj_raw, xyz_proj, sample_idx, raw_feat_point):
        B = xyz_proj.shape[0]
        n = self.out_h * self.out_w
        sample_idx = _centres(sample_idx, B, self.out_h, self.out_w, self.stride_H, self.stride_W, xyz_proj.device)
        new_xyz_proj = _take_centres(xyz_proj, sample_idx, B, self.H, self.W)
        new_xyz_proj_raw = _take_centres(xyz_proj_raw, sample_idx, B, self.H, self.W)
        xyz_pr = xyz_proj if self.usetrans else xyz_proj_raw
        flat, _ = select_flat(xyz_pr, xyz_pr, _centre_index(sample_idx, B, self.out_h, self.out_w),
                              self.kernel_size, self.nsample, FLAG_SHIFT | FLAG_COPY, self.distance)
        src, ctr = (xyz_proj_raw, new_xyz_proj_raw) if raw_feat_point else (xyz_proj, new_xyz_proj)
        grouped_xyz = gather_rows(src, flat)                        # B,N,K,3
        grouped_xyz_norm = grouped_xyz - ctr.reshape(B, n, 1, 3)
        return sample_idx, new_xyz_proj_raw, new_xyz_proj, flat, grouped_xyz, grouped_xyz_norm

    def _mlp_max(self, new_points, B):
        return run_mlp(self.mlp_convs, new_points, reduce_k=True).view(B, self.out_h, self.out_w, -1)

    def forward(self, xyz_proj_raw, xyz_proj, feature_proj, sample_idx=None, cfg=None, raw_feat_point=False):
        """xyz_proj_raw / xyz_proj (B,H,W,3), feature_proj (B,H,W,C) ->
        (new_xyz_proj_raw, new_xyz_proj (B,h,w,3), new_points (B,h,w,mlp[-1]), grouped_xyz, sample_idx)"""
        B = xyz_proj.shape[0]
        sample_idx, new_raw, new_xyz, flat, grouped_xyz, norm = self._group(xyz_proj_raw, xyz_proj, sample_idx,
                                                                            raw_feat_point)
        grouped_points = gather_rows(feature_proj, flat)            # B,N,K,C
        new_points = self._mlp_max(torch.cat([norm, grouped_points], -1), B)
        return new_raw, new_xyz, new_points, grouped_xyz, sample_idx

    def forward_center(self, xyz_proj_raw, xyz_proj, feature_proj, sample_idx=None, cfg=None, using_intens=False,
                       raw_feat_point=False):
        """First level: the feature is geometric only -- (offset, centre, neighbour, |offset|) = 10
        channels (+ feature_proj when using_intens).  The reference also gathers feature_proj when
        it is not used (:157); this does not."""
        B = xyz_proj.shape[0]
        needs_grad = torch.is_grad_enabled() and (xyz_proj.requires_grad or xyz_proj_raw.requires_grad)
        if USE_FUSED_SA_GEOMETRY and xyz_proj.is_cuda and xyz_proj.dtype == torch.float32 and not using_intens and not needs_grad:
            return self._forward_center_fused(xyz_proj_raw, xyz_proj, sample_idx, raw_feat_point)
        sample_idx, new_raw, new_xyz, flat, grouped_xyz, norm = self._group(xyz_proj_raw, xyz_proj, sample_idx,
                                                                            raw_feat_point)
        centre = new_xyz.reshape(B, self.out_h * self.out_w, 1, 3).expand(-1, -1, norm.shape[2], -1)
        dist = torch.norm(norm, p=2, dim=3, keepdim=True)
        parts = [norm, centre, grouped_xyz, dist]
        if using_intens:
            parts.append(gather_rows(feature_proj, flat))
        new_points = self._mlp_max(torch.cat(parts, -1), B)
        return new_raw, new_xyz, new_points, grouped_xyz, sample_idx

    def _forward_center_fused(self, xyz_proj_raw, xyz_proj, sample_idx, raw_feat_point):
        """forward_center with the ten geometric channels written by one kernel (csrc/gather.cu sa_geometry) instead of
        gather + subtract + broadcast + norm + concatenate; the coordinates carry no gradient at the first level."""
        from .. import _cabi
        B, n, K = xyz_proj.shape[0], self.out_h * self.out_w, self.nsample
        dev, f32 = xyz_proj.device, torch.float32
        sample_idx = _centres(sample_idx, B, self.out_h, self.out_w, self.stride_H, self.stride_W, dev)
        new_xyz = _take_centres(xyz_proj, sample_idx, B, self.H, self.W)
        new_raw = _take_centres(xyz_proj_raw, sample_idx, B, self.H, self.W)
        xyz_pr = xyz_proj if self.usetrans else xyz_proj_raw
        flat, _ = select_flat(xyz_pr, xyz_pr, _centre_index(sample_idx, B, self.out_h, self.out_w),
                              self.kernel_size, self.nsample, FLAG_SHIFT | FLAG_COPY, self.distance)
        src, ctr = (xyz_proj_raw, new_raw) if raw_feat_point else (xyz_proj, new_xyz)
        src, ctr, cen = src.detach().contiguous(), ctr.detach().contiguous(), new_xyz.detach().contiguous()
        flat = flat.contiguous()
        operand = torch.empty(B, n, K, 10, dtype=f32, device=dev)
        _cabi.call("i2p_sa_geometry", dev, B, self.H * self.W, n, K, _cabi._ptr(src, f32, "xyz", dev), _cabi._ptr(ctr, f32, "centres", dev),
                   _cabi._ptr(cen, f32, "centres", dev), _cabi._ptr(flat, torch.int32, "neighbour index", dev), operand.data_ptr())
        new_points = self._mlp_max(operand, B)
        return new_raw, new_xyz, new_points, operand[..., 6:9], sample_idx

    def set_bn(self):
        for conv in self.mlp_convs:
            conv.set_bn()


class ProjSetUpconvModule(nn.Module):
    """Feature propagation from a coarse level (xyz2, feat2) to a finer one (xyz1, feat1)."""

    def __init__(self, H, W, out_h, out_w, stride_H, stride_W, kernel_size, nsample, distance, in_channels, mlp,
                 mlp2, use_trans=False, use_bn_p=True, use_bn_input=True):
        super().__init__()
        self.nsample, self.mlp, self.mlp2 = nsample, mlp, mlp2
        self.H, self.W, self.out_h, self.out_w = H, W, out_h, out_w
        self.stride_H, self.stride_W = stride_H, stride_W
        self.kernel_size, self.distance, self.use_trans = kernel_size, distance, use_trans
        self.last_channel = in_channels[-1] + 3
        self.mlp_conv = nn.ModuleList()
        self.mlp2_conv = nn.ModuleList()
        if mlp is not None:
            for c in mlp:
                self.mlp_conv.append(Conv2d(self.last_channel, c, [1, 1], stride=[1, 1], bn=use_bn_p,
                                            use_bn_input=use_bn_input))
                self.last_channel = c
        self.last_channel = (mlp[-1] if len(mlp) > 0 else self.last_channel) + in_channels[0]
        if mlp2 is not None:
            for c in mlp2:
                self.mlp2_conv.append(Conv2d(self.last_channel, c, [1, 1], stride=[1, 1], bn=use_bn_p,
                                             use_bn_input=use_bn_input))
                self.last_channel = c

    def forward(self, xyz1_raw, xyz2_raw, xyz1, xyz2, idx_n2, feat1, feat2, cfg=None, raw_feat_point=False):
        """xyz1* (B,out_h,out_w,3), xyz2* (B,H,W,3), idx_n2 (B,out_h*out_w,2) int32 (or a StrideGrid),
        feat1 (B,out_h,out_w,c1), feat2 (B,H,W,c2) -> (B, out_h*out_w, mlp2[-1])"""
        B = xyz1.shape[0]
        n = self.out_h * self.out_w
        xyz1_pr, xyz2_pr = (xyz1, xyz2) if self.use_trans else (xyz1_raw, xyz2_raw)
        flat, _ = select_flat(xyz1_pr, xyz2_pr, idx_n2, self.kernel_size, self.nsample, FLAG_SHIFT | FLAG_COPY,
                              self.distance, self.stride_H, self.stride_W)
        src2, src1 = (xyz2_raw, xyz1_raw) if raw_feat_point else (xyz2, xyz1)
        xyz_diff = gather_rows(src2, flat) - src1.reshape(B, n, 1, 3)
        upfeats = torch.cat([gather_rows(feat2, flat), xyz_diff], dim=3)   # B,N,K,C+3
        if self.mlp is not None and len(self.mlp_conv) > 0:
            feat1_new = run_mlp(self.mlp_conv, upfeats, reduce_k=True)
        else:
            feat1_new = torch.max(upfeats, dim=2)[0]
        feat1_new = feat1_new.view(B, self.out_h, self.out_w, -1)
        if feat1 is not None:
            feat1_new = torch.cat([feat1_new, feat1], dim=3)
        if self.mlp2 is not None:
            feat1_new = run_mlp(self.mlp2_conv, feat1_new)
        return feat1_new.reshape(B, n, -1)

    def set_bn(self):
        for conv in list(self.mlp_conv) + list(self.mlp2_conv):
            conv.set_bn()


USE_FUSED_SA_GEOMETRY = True   # False: the first level's geometric operand through gather / subtract / norm / concatenate
USE_FUSED_CV = True   # False: the reference's broadcast / mask / concatenate / softmax formulation through ATen


class _CvBuild(torch.autograd.Function):
    """First-layer operand of the cost volume (csrc/cv.cu cv_build): X (B,N,K,6+C[+C]) and the 6 coordinate
    channels alone, from xyz1 (B,N,3), xyz2 (B,N2,3), pi (B,N,C), qi (B,N2,C), maxc (B,N2,C) | None,
    idx (B,N,K) int32 | None (None: K == N2, every point sees every pixel)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, pi, qi, maxc, idx):
        from .. import _cabi
        f32, dev = torch.float32, pi.device
        xyz1, xyz2, pi, qi = xyz1.contiguous(), xyz2.contiguous(), pi.contiguous(), qi.contiguous()
        maxc = maxc.contiguous() if maxc is not None else None
        B, N, C = pi.shape
        N2 = qi.shape[1]
        K = idx.shape[2] if idx is not None else N2
        Cx = 6 + C + (C if maxc is not None else 0)
        X = torch.empty(B, N, K, Cx, dtype=f32, device=dev)
        xyz6 = torch.empty(B, N, K, 6, dtype=f32, device=dev)
        _cabi.call("i2p_cv_build", dev, B, N, K, N2, C, _cabi._ptr(xyz1, f32, "xyz1", dev), _cabi._ptr(xyz2, f32, "xyz2", dev),
                   _cabi._ptr(pi, f32, "pi", dev), _cabi._ptr(qi, f32, "qi", dev),
                   _cabi._ptr(maxc, f32, "maxc", dev) if maxc is not None else None,
                   _cabi._ptr(idx, torch.int32, "idx", dev) if idx is not None else None, X.data_ptr(), xyz6.data_ptr())
        ctx.save_for_backward(pi, qi, idx if idx is not None else torch.empty(0, device=dev))
        ctx.meta = (B, N, K, N2, C, maxc is not None, idx is not None)
        return X, xyz6

    @staticmethod
    def backward(ctx, dX, dxyz6):
        from .. import _cabi
        pi, qi, idx = ctx.saved_tensors
        B, N, K, N2, C, has_max, has_idx = ctx.meta
        f32, dev = torch.float32, pi.device
        dX = dX.contiguous()
        dxyz6 = dxyz6.contiguous() if dxyz6 is not None else None
        # five accumulators carved out of one zero-filled buffer (a single fill launch)
        sizes = [B * N * 3, B * N * C, B * N2 * 3, B * N2 * C, B * N2 * C if has_max else 0]
        from .. import scratch
        flat = scratch.zeros(sum(sizes), f32, dev)
        offs = [0]
        for n_ in sizes:
            offs.append(offs[-1] + n_)
        dxyz1 = flat[offs[0]:offs[1]].view(B, N, 3)
        dpi = flat[offs[1]:offs[2]].view(B, N, C)
        dxyz2 = flat[offs[2]:offs[3]].view(B, N2, 3)
        dqi = flat[offs[3]:offs[4]].view(B, N2, C)
        dmaxc = flat[offs[4]:offs[5]].view(B, N2, C) if has_max else None
        _cabi.call("i2p_cv_build_bwd", dev, B, N, K, N2, C, int(has_max), _cabi._ptr(dX, f32, "dX", dev),
                   _cabi._ptr(dxyz6, f32, "dxyz6", dev) if dxyz6 is not None else None, pi.data_ptr(), qi.data_ptr(),
                   idx.data_ptr() if has_idx else None, dxyz1.data_ptr(), dxyz2.data_ptr(), dpi.data_ptr(), dqi.data_ptr(),
                   dmaxc.data_ptr() if has_max else None)
        return dxyz1, dxyz2, dpi, dqi, dmaxc, None


_PREP_SCRATCH = {}


def _prep_scratch(dev, B, N, N2, C):
    """Work space of cv_prep's cross-block reduction: per-block extrema and one ticket per cloud that the kernel itself
    leaves at zero -- so one zero-filled buffer per (stream, sizes) serves every launch (launches on one stream are
    ordered; different streams get different buffers)."""
    from .. import _cabi
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, B, N, N2, C)
    buf = _PREP_SCRATCH.get(key)
    if buf is None:
        buf = _PREP_SCRATCH[key] = torch.zeros(_cabi.lib().i2p_cv_prep_scratch_floats(B, N, N2, C), dtype=torch.float32, device=dev)
    return buf


class _CvPrep(torch.autograd.Function):
    """Operand preparation of the cost volume as one kernel per direction (csrc/cv.cu cv_prep): depth restoration
    xyz = uv * z, the row-wise standardisation of the point and pixel features and, for the backward-validation channel,
    maxc[k,c] = max over the valid points n of pi[n,c] * qi[k,c] = qi[k,c] times the largest (qi > 0) or smallest valid
    pi[:,c] -- multiplication by a constant is monotonic also after rounding (:379-397).
    uv (B,N,3), z (B,N,1), pf (B,N,C), qf (B,N2,C) -> xyz (B,N,3), pi (B,N,C), qi (B,N2,C), maxc (B,N2,C) | None"""

    @staticmethod
    def forward(ctx, uv, z, pf, qf, has_max):
        from .. import _cabi
        f32, dev = torch.float32, pf.device
        uv, z, pf, qf = uv.contiguous(), z.contiguous(), pf.contiguous(), qf.contiguous()
        B, N, C = pf.shape
        N2 = qf.shape[1]
        xyz, pi, qi = torch.empty_like(uv), torch.empty_like(pf), torch.empty_like(qf)
        den = torch.empty(B * (N + N2), dtype=f32, device=dev)
        den_p, den_q = den[:B * N], den[B * N:]
        if has_max:
            maxc = torch.empty_like(qf)
            ext = torch.empty(2, B, C, dtype=f32, device=dev)
            arg = torch.empty(2, B, C, dtype=torch.int32, device=dev)
        else:
            maxc = ext = arg = None
        P = lambda x, n: _cabi._ptr(x, f32, n, dev)
        scratch = _prep_scratch(dev, B, N, N2, C) if has_max else None
        _cabi.call("i2p_cv_prep_fwd", dev, B, N, N2, C, int(has_max), P(uv, "warped_xyz"), P(z, "lidar_z"), P(pf, "warped_points"),
                   P(qf, "f2_points"), xyz.data_ptr(), pi.data_ptr(), qi.data_ptr(), den_p.data_ptr(), den_q.data_ptr(),
                   maxc.data_ptr() if has_max else None, ext[0].data_ptr() if has_max else None,
                   ext[1].data_ptr() if has_max else None, arg[0].data_ptr() if has_max else None,
                   arg[1].data_ptr() if has_max else None, scratch.data_ptr() if has_max else None)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(uv, z, pi, qi, den, ext if has_max else empty, arg if has_max else empty)
        ctx.meta = (B, N, N2, C, bool(has_max))
        ctx.set_materialize_grads(False)
        return xyz, pi, qi, maxc

    @staticmethod
    def backward(ctx, d_xyz, d_pi, d_qi, d_maxc):
        from .. import _cabi
        uv, z, pi, qi, den, ext, arg = ctx.saved_tensors
        B, N, N2, C, has_max = ctx.meta
        f32, dev = torch.float32, pi.device
        d_uv, d_z, d_pf, d_qf = torch.empty_like(uv), torch.empty_like(z), torch.empty_like(pi), torch.empty_like(qi)
        G = lambda g, n: _cabi._ptr(g.contiguous(), f32, n, dev) if g is not None else None
        # the contiguous copies must outlive the launch: keep them referenced until the call returns
        keep = [g.contiguous() if g is not None else None for g in (d_xyz, d_pi, d_qi, d_maxc)]
        _cabi.call("i2p_cv_prep_bwd", dev, B, N, N2, C, int(has_max), uv.data_ptr(), z.data_ptr(), pi.data_ptr(), qi.data_ptr(),
                   den[:B * N].data_ptr(), den[B * N:].data_ptr(), ext[0].data_ptr() if has_max else None,
                   ext[1].data_ptr() if has_max else None, arg[0].data_ptr() if has_max else None,
                   arg[1].data_ptr() if has_max else None, *[G(g, "gradient") for g in keep], d_uv.data_ptr(), d_z.data_ptr(),
                   d_pf.data_ptr(), d_qf.data_ptr())
        return d_uv, d_z, d_pf, d_qf, None


class _SoftmaxWSum(torch.autograd.Function):
    """sum_k softmax_k(logit [masked]) * value over axis 2 of (B,N,K,C) tensors -> (B,N,C)  (csrc/cv.cu)."""

    @staticmethod
    def forward(ctx, logit, value, mask):
        from .. import _cabi
        f32, dev = torch.float32, logit.device
        logit, value = logit.contiguous(), value.contiguous()
        mask = mask.contiguous() if mask is not None else None
        B, N, K, C = logit.shape
        out = torch.empty(B, N, C, dtype=f32, device=dev)
        stat = torch.empty(B, N, C, 2, dtype=f32, device=dev) if C % 8 == 0 else None      # (max, 1 / sum) per (group, channel)
        _cabi.call("i2p_softmax_wsum", dev, B * N, K, C, _cabi._ptr(logit, f32, "logit", dev), _cabi._ptr(value, f32, "value", dev),
                   _cabi._ptr(mask, f32, "mask", dev) if mask is not None else None, out.data_ptr(),
                   stat.data_ptr() if stat is not None else None)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(logit, value, out, mask if mask is not None else empty, stat if stat is not None else empty)
        ctx.has_mask, ctx.has_stat = mask is not None, stat is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        from .. import _cabi
        logit, value, out, mask, stat = ctx.saved_tensors
        f32, dev = torch.float32, logit.device
        B, N, K, C = logit.shape
        gout = gout.contiguous()
        dlogit, dvalue = torch.empty_like(logit), torch.empty_like(value)
        _cabi.call("i2p_softmax_wsum_bwd", dev, B * N, K, C, logit.data_ptr(), value.data_ptr(),
                   mask.data_ptr() if ctx.has_mask else None, out.data_ptr(), _cabi._ptr(gout, f32, "grad", dev),
                   stat.data_ptr() if ctx.has_stat else None, dlogit.data_ptr(), dvalue.data_ptr())
        return dlogit, dvalue, None


def _softmax_wsum(logit, value, mask=None):
    """(B,N,K,C) logits / values [+ (B,N,K,1) validity mask] -> (B,N,C)"""
    if USE_FUSED_CV and logit.is_cuda:
        return _SoftmaxWSum.apply(logit, value, mask.reshape(mask.shape[:3]) if mask is not None else None)
    if mask is not None:
        logit = logit * mask + -1e10 * (1 - mask)
    return torch.sum(F.softmax(logit, dim=2) * value, dim=2)


def _standardise(x):
    """(x - mean) / max(std, 1e-12) over the channel axis, unbiased std (:384-389)."""
    return (x - x.mean(-1, keepdim=True)) / torch.clip(x.std(-1, keepdim=True), min=1e-12)


class CostVolume(nn.Module):
    """2D-3D cost volume: every LiDAR point attends over image pixels (all of them, or its
    nsample_q nearest on the normalised plane), then over its nsample 3-D neighbours."""

    def __init__(self, H, W, kernel_size, distance, nsample, nsample_q, rgb_in_channels, lidar_in_channels, mlp1,
                 mlp2, backward_validation=False, use_trans=False, use_bn_p=True, use_bn_input=True):
        super().__init__()
        self.H, self.W, self.nsample, self.nsample_q, self.distance = H, W, nsample, nsample_q, distance
        self.mlp1, self.mlp2, self.kernel_size = mlp1, mlp2, kernel_size
        self.backward_validation, self.use_trans = backward_validation, use_trans
        kw = dict(stride=[1, 1], bn=use_bn_p, use_bn_input=use_bn_input)
        self.in_channels = rgb_in_channels + (lidar_in_channels if backward_validation else 0) + 6
        self.mlp1_convs, self.mlp2_convs, self.mlp2_convs_2 = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for c in mlp1:
            self.mlp1_convs.append(Conv2d(self.in_channels, c, [1, 1], **kw))
            self.in_channels = c
        self.pi_encoding = Conv2d(6, mlp1[-1], [1, 1], **kw)
        self.in_channels = 2 * mlp1[-1]
        for c in mlp2:
            self.mlp2_convs.append(Conv2d(self.in_channels, c, [1, 1], **kw))
            self.in_channels = c
        self.pc_encoding = Conv2d(10, mlp1[-1], [1, 1], **kw)
        self.in_channels = 2 * mlp1[-1] + lidar_in_channels
        for c in mlp2:
            self.mlp2_convs_2.append(Conv2d(self.in_channels, c, [1, 1], **kw))
            self.in_channels = c

    def forward(self, xyz_proj_raw, warped_xyz, warped_points, idx_n2, f2_xyz, f2_points, lidar_z, cfg=None):
        """warped_xyz (B,HW,3) on the normalised plane, warped_points (B,HW,C), idx_n2 (B,HW,2),
        f2_xyz (B,N2,3), f2_points (B,N2,C), lidar_z (B,HW,1) -> (B,H,W,mlp2[-1])"""
        B, N, C = warped_points.shape
        if USE_FUSED_CV and warped_points.is_cuda:
            pi_feat1_new, pi_xyz_diff_concat, warped_xyz = self._first_layer_operand_fused(warped_xyz, warped_points, f2_xyz,
                                                                                       f2_points, lidar_z)
        else:
            pi_feat1_new, pi_xyz_diff_concat, warped_xyz = self._first_layer_operand(warped_xyz, warped_points, f2_xyz,
                                                                                 f2_points, lidar_z)
        # The neighbourhoods and the position encoding of the second stage depend on the coordinates alone, the
        # pixel-pair encoding on xyz6 alone: both run beside the feature MLP of the first stage (streams.py).
        warped_xyz_bhw = warped_xyz.view(B, self.H, self.W, 3)
        xyz_pr = warped_xyz_bhw if self.use_trans else xyz_proj_raw
        with Fork(warped_xyz, xyz_pr) as pc_branch:
            flat, valid_mask = select_flat(xyz_pr, xyz_pr, idx_n2, self.kernel_size, self.nsample, FLAG_SHIFT,
                                           self.distance)
            pc_xyz_grouped = gather_rows(warped_xyz_bhw, flat)           # B,N,K,3
            pc_xyz_new = warped_xyz[:, :, None, :].expand(-1, -1, self.nsample, -1)
            pc_xyz_diff = pc_xyz_grouped - pc_xyz_new
            pc_euc_diff = torch.sqrt(torch.sum(pc_xyz_diff * pc_xyz_diff, dim=3, keepdim=True) + 1e-20)
            pc_xyz_encoding = self.pc_encoding(torch.cat([pc_xyz_new, pc_xyz_grouped, pc_xyz_diff, pc_euc_diff], dim=3))
        with Fork(pi_xyz_diff_concat) as pi_branch:
            pi_xyz_encoding = self.pi_encoding(pi_xyz_diff_concat)
        pi_feat1_new = run_mlp(self.mlp1_convs, pi_feat1_new)
        pi_concat = torch.cat([pi_branch.join(pi_xyz_encoding), pi_feat1_new], dim=3)
        pi_concat = run_mlp(self.mlp2_convs, pi_concat)
        pi_feat1_new = _softmax_wsum(pi_concat, pi_feat1_new)        # B,N,mlp1[-1]

        # second stage: re-weight over the nsample 3-D neighbours of every point
        flat, valid_mask, pc_xyz_encoding = pc_branch.join(flat, valid_mask, pc_xyz_encoding)
        pc_points_grouped = gather_rows(pi_feat1_new, flat)          # B,N,K,mlp1[-1]
        pc_points_new = warped_points[:, :, None, :].expand(-1, -1, self.nsample, -1)
        pc_concat = torch.cat([pc_xyz_encoding, pc_points_new, pc_points_grouped], dim=-1)
        pc_concat = run_mlp(self.mlp2_convs_2, pc_concat)
        pc_feat1_new = _softmax_wsum(pc_concat, pc_points_grouped, valid_mask)
        return pc_feat1_new.view(B, self.H, self.W, -1)

    def _first_layer_operand(self, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
        """The reference's formulation (:366-400): broadcast operands, mask, max, concatenate.
        -> (B,N,K,6+C[+C]) operand of mlp1, (B,N,K,6) coordinate pairs, depth-restored xyz (B,N,3)"""
        B, N, C = warped_points.shape
        if self.nsample_q > 0:
            idx = knn_point(self.nsample_q, f2_xyz.contiguous(), warped_xyz.contiguous()).to(torch.int32)
            qi_xyz = gather_rows(f2_xyz, idx)                       # B,N,K,3
            qi_points = gather_rows(f2_points, idx)                 # B,N,K,C
        else:
            qi_xyz = f2_xyz.unsqueeze(1).expand(-1, N, -1, -1)      # B,N,N2,3
            qi_points = f2_points.unsqueeze(1)                      # B,1,N2,C (broadcast over points)
        K = qi_xyz.shape[2]
        warped_xyz = warped_xyz.mul(lidar_z)                        # restore depth (:379)
        pi_xyz_diff_concat = torch.cat([warped_xyz[:, :, None, :].expand(-1, -1, K, -1), qi_xyz], dim=3)
        pi_n = _standardise(warped_points)[:, :, None, :]           # B,N,1,C
        qi_n = _standardise(qi_points)                              # B,(1|N),K,C
        corr = pi_n * qi_n                                          # B,N,K,C
        parts = [pi_xyz_diff_concat, corr]
        if self.backward_validation:
            valid = check_valid(warped_xyz).unsqueeze(-1)           # B,N,1,1
            masked = corr * valid + -1e10 * (1 - valid)
            parts.append(torch.max(masked, 1, keepdim=True)[0].expand(-1, N, -1, -1))
        return torch.cat(parts, dim=3), pi_xyz_diff_concat, warped_xyz

    def _first_layer_operand_fused(self, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
        """Same values from one build kernel.  The per-row standardisation commutes with the pixel gather, so the
        pixels are standardised once (B,N2,C) instead of per (point, neighbour); the backward-validation maximum
        over the points, max_n(valid_n ? pi[n,c] * qi[k,c] : -1e10), is qi[k,c] times the largest (qi > 0) or
        smallest (qi < 0) valid pi[:,c] -- multiplication by a constant is monotonic also after rounding -- so
        it needs two (B,C) reductions instead of passes over the (B,N,K,C) product."""
        B, N, C = warped_points.shape
        idx = None
        if self.nsample_q > 0:
            idx = knn_point(self.nsample_q, f2_xyz.contiguous(), warped_xyz.contiguous()).to(torch.int32)
        warped_xyz, pi_n, qi_n, maxc = _CvPrep.apply(warped_xyz, lidar_z, warped_points, f2_points, self.backward_validation)
        X, xyz6 = _CvBuild.apply(warped_xyz, f2_xyz, pi_n, qi_n, maxc, idx)
        return X, xyz6, warped_xyz

    def set_bn(self):
        for conv in list(self.mlp2_convs) + list(self.mlp1_convs) + list(self.mlp2_convs_2):
            conv.set_bn()
        self.pc_encoding.set_bn()
        self.pi_encoding.set_bn()


class _PoseHeadFn(torch.autograd.Function):
    """PoseHead.forward as one kernel per direction (csrc/head.cu): softmax of the mask over the points, weighted pooling,
    hidden layer, dropout, the two linear heads, quaternion normalisation.  Parameter gradients accumulate straight into
    the step engine's flat gradient buffer when there is one."""

    @staticmethod
    def forward(ctx, pred, mask, w1, b1, wq, bq, wt, bt, drop):
        from .. import _cabi
        f32, dev = torch.float32, pred.device
        pred, mask = pred.contiguous(), mask.contiguous()
        B, N, C = pred.shape
        Hd = w1.shape[0]
        mask_p = torch.empty(B, N, C, dtype=f32, device=dev)
        pooled = torch.empty(B, C, dtype=f32, device=dev)
        hidden = torch.empty(B, Hd, dtype=f32, device=dev)
        q_raw, q, t = (torch.empty(B, 4, dtype=f32, device=dev), torch.empty(B, 4, dtype=f32, device=dev),
                       torch.empty(B, 3, dtype=f32, device=dev))
        P = lambda x, n: _cabi._ptr(x, f32, n, dev)
        _cabi.call("i2p_pose_head_fwd", dev, B, N, C, Hd, P(pred, "prediction"), P(mask, "mask"), P(w1, "hidden weight"),
                   P(b1, "hidden bias"), P(wq, "quat weight"), P(bq, "quat bias"), P(wt, "trans weight"), P(bt, "trans bias"),
                   P(drop, "dropout mask") if drop is not None else None, mask_p.data_ptr(), pooled.data_ptr(),
                   hidden.data_ptr(), q_raw.data_ptr(), q.data_ptr(), t.data_ptr())
        ctx.save_for_backward(pred, mask_p, pooled, hidden, q_raw, w1, wq, wt, *([drop] if drop is not None else []))
        ctx.params = (w1, b1, wq, bq, wt, bt)
        ctx.mark_non_differentiable(mask_p)
        ctx.set_materialize_grads(False)
        return q, t, mask_p

    @staticmethod
    def backward(ctx, dq, dt, _):
        from .. import _cabi
        from ..engine import grad_sink
        saved = ctx.saved_tensors
        pred, mask_p, pooled, hidden, q_raw, w1, wq, wt = saved[:8]
        drop = saved[8] if len(saved) > 8 else None
        f32, dev = torch.float32, pred.device
        B, N, C = pred.shape
        Hd = w1.shape[0]
        dpred, dmask = torch.empty_like(pred), torch.empty_like(pred)
        sinks = [grad_sink(p) for p in ctx.params]
        grads = [None if s is not None else torch.zeros_like(p) for s, p in zip(sinks, ctx.params)]
        tgt = [s if s is not None else g for s, g in zip(sinks, grads)]
        _cabi.call("i2p_pose_head_bwd", dev, B, N, C, Hd, pred.data_ptr(), mask_p.data_ptr(), pooled.data_ptr(), hidden.data_ptr(),
                   q_raw.data_ptr(), drop.data_ptr() if drop is not None else None, w1.data_ptr(), wq.data_ptr(), wt.data_ptr(),
                   _cabi._ptr(dq.contiguous(), f32, "dq", dev) if dq is not None else None,
                   _cabi._ptr(dt.contiguous(), f32, "dt", dev) if dt is not None else None, dpred.data_ptr(), dmask.data_ptr(),
                   tgt[0].data_ptr(), tgt[1].data_ptr(), tgt[2].data_ptr(), tgt[3].data_ptr(), tgt[4].data_ptr(), tgt[5].data_ptr())
        return (dpred, dmask, *grads, None)


class PoseHead(nn.Module):
    """Mask-weighted pooling over points -> hidden -> (unit quaternion, translation)."""

    def __init__(self, in_channels, mlp1, mlp2, hidden, q_dim, t_dim, dropout_rate=0.5, split_dp=False,
                 pos_embed=False, sigmoid=False, maxhead=False):
        super().__init__()
        self.sigmoid, self.maxhead, self.pos_embed = sigmoid, maxhead, pos_embed
        in_channel, _ = in_channels
        self.DP1 = nn.Identity() if split_dp else nn.Dropout(dropout_rate)
        self.DP2 = nn.Dropout(dropout_rate) if split_dp else nn.Identity()
        self.hidden_layer = Conv1d(in_channel, hidden, use_activation=False)
        self.quat_head = Conv1d(hidden, q_dim, use_activation=False)
        self.trans_head = Conv1d(hidden, t_dim, use_activation=False)

    def forward(self, prediction, mask, xyz, feature, projection_mask):
        """prediction, mask (B,N,C) -> q (B,4), t (B,3), mask_p (B,N,C)"""
        if prediction.is_cuda and prediction.dtype == torch.float32:
            return self._forward_fused(prediction, mask, projection_mask)
        if not self.sigmoid:
            if projection_mask is not None:
                projection_mask = torch.argmax(projection_mask.detach(), dim=-1, keepdim=True).float()
                mask = mask * projection_mask + -1e10 * (1. - projection_mask)
        else:
            prediction = prediction * projection_mask
        if self.maxhead:
            mask = torch.max(mask, dim=-1, keepdim=True)[0]
        mask_p = F.softmax(mask, dim=1)
        pooled = torch.sum(prediction * mask_p, dim=1, keepdim=True)      # B,1,C
        hidden = self.DP1(self.hidden_layer(pooled))
        q = self.quat_head(self.DP2(hidden)).squeeze(1)
        t = self.trans_head(self.DP2(hidden)).squeeze(1)
        q = q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10)
        return q, t, mask_p


def _pose_head_fused(self, prediction, mask, projection_mask):
    """The configuration every shipped config uses (no sigmoid / max head / projection mask / split dropout, plain linear
    layers); anything else has no kernel and raises -- there is no library fallback on the device."""
    lin = [m.composed_module for m in (self.hidden_layer, self.quat_head, self.trans_head)]
    plain = all(isinstance(m[1], nn.Identity) and isinstance(m[2], nn.Identity) and m[0].kernel_size == (1,) for m in lin)
    if self.sigmoid or self.maxhead or projection_mask is not None or not isinstance(self.DP2, nn.Identity) or not plain:
        from .. import _cabi
        raise _cabi.I2PError("PoseHead: this configuration is not covered by the sm_100a pose-head kernel")
    drop = None
    if isinstance(self.DP1, nn.Dropout) and self.training and self.DP1.p > 0:
        B, Hd = prediction.shape[0], lin[0][0].out_channels
        drop = F.dropout(torch.ones(B, Hd, device=prediction.device), self.DP1.p, True)   # 0 or 1 / (1 - p), graph-safe RNG
    w = lambda m: m[0].weight      # (out, in, 1): the Parameter itself, so that its gradient sink is found
    return _PoseHeadFn.apply(prediction, mask, w(lin[0]), lin[0][0].bias, w(lin[1]), lin[1][0].bias, w(lin[2]), lin[2][0].bias, drop)


PoseHead._forward_fused = _pose_head_fused


class FlowPredictor(nn.Module):
    """Per-point MLP over the concatenation [points_f1, cost_volume, upsampled_feat]."""

    def __init__(self, in_channels, mlp, is_training, bn_decay, bn=True, use_bn_input=True):
        super().__init__()
        self.in_channels, self.mlp, self.is_training, self.bn_decay, self.bn = in_channels, mlp, is_training, bn_decay, bn
        self.mlp_conv = nn.ModuleList()
        for c in mlp:
            self.mlp_conv.append(Conv2d(self.in_channels, c, [1, 1], stride=[1, 1], bn=bn, use_bn_input=use_bn_input))
            self.in_channels = c

    def forward(self, points_f1, upsampled_feat, cost_volume):
        parts = [points_f1, cost_volume] + ([upsampled_feat] if upsampled_feat is not None else [])
        return run_mlp(self.mlp_conv, torch.cat(parts, -1).unsqueeze(2)).squeeze(2)

    def set_bn(self):
        for conv in self.mlp_conv:
            conv.set_bn()
