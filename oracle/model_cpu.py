"""CPU restatement of the reference's large-range forward pass (RegNet_v2), functional style.

TEST INFRASTRUCTURE ONLY (see oracle/i2p_oracle.c): the checker of tests/, the `cpu_baseline`
leg and the `--impl reference` arm of bench.py.  Nothing in i2pnet_b200/ imports it.

The reference has no CPU path (its extensions are CUDA-only, `.cuda()` is hard-coded in
src/modules/warp_utils.py:5,18-19), so "the reference on host cores" is this restatement: naive
PyTorch on CPU tensors for everything the reference does in PyTorch, written the way the
reference writes it (materialised `repeat`s, NCHW 1x1 convs between permutes, matmul + topk kNN,
torch.gather with an expanded index, per-sample index_put_ loop), and the C oracle for its two
CUDA extensions.  It is a pure function of a reference-layout state_dict, so it shares no
module code with the product.

Pinned against tests/golden/ref_model_kitti_b2.npz, which was recorded from the real reference
Python (tests/golden/make_golden.py); see tests/test_oracle_golden_model.py.

Citations: src/modellearn_proj_center.py (M), src/projectPN/PPBackbone_center.py (P),
src/projectPN/utils.py (U), src/modules/basicConv.py (C), src/modules/warp_utils.py (W).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


class KittiShape:
    """The shape constants of src/config_proj_lidarcenter.py that the forward reads."""
    init_H, init_W = 64, 1800
    stride_Hs, stride_Ws = [4, 2, 2, 1], [8, 2, 2, 2]
    kernel_sizes = [[9, 15], [9, 15], [5, 9], [5, 9]]
    down_conv_dis = [0.75, 3.0, 6.0, 12.0]
    lidar_group_samples = [32, 16, 16, 16, 16]
    cost_volume_dis, cost_volume_kernel_size = [4.5, 4.5], [[3, 5], [3, 5]]
    cost_volume_nsamples = [4, [-1, 32]]
    backward_validation = [True, False]
    up_conv_dis, up_conv_kernel_size, setupconv_nsamples = [9.0, 9.0], [[5, 9], [5, 9]], [8, 8]
    rgb_pool_strides = [[2, 1, 1, 1, 2], [2, 1, 1, 1, 2], [1, 1, 1, 1, 2]]
    fup, fdown = 2.0, -24.8


def _conv(sd, p, x, leaky=True, act=True):
    """P:35-46 -- permute to NCHW, 1x1 conv, BatchNorm2d on batch statistics, activation, permute back."""
    x = x.permute(0, 3, 2, 1)
    y = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"])
    if p + ".bn_linear.weight" in sd:
        y = F.batch_norm(y, None, None, sd[p + ".bn_linear.weight"], sd[p + ".bn_linear.bias"], True, 0.1, 1e-5)
    if act:
        y = F.leaky_relu(y, 0.1) if leaky else F.relu(y)
    return y.permute(0, 3, 2, 1)


def _mlp(sd, prefix, n, x, leaky=True):
    for i in range(n):
        x = _conv(sd, "%s.%d" % (prefix, i), x, leaky)
    return x


def _count(sd, prefix):
    return len({k[len(prefix) + 1:].split(".")[0] for k in sd if k.startswith(prefix + ".")})


def _rgb(sd, name, x, strides):
    """C:6-20 -- [conv3x3, BatchNorm2d (training statistics), LeakyReLU(0.1), MaxPool3x3] x 5"""
    for i, s in enumerate(strides):
        x = F.conv2d(x, sd["%s.%d.weight" % (name, 4 * i)], sd["%s.%d.bias" % (name, 4 * i)], padding=1)
        x = F.batch_norm(x, None, None, sd["%s.%d.weight" % (name, 4 * i + 1)], sd["%s.%d.bias" % (name, 4 * i + 1)],
                         True, 0.1, 1e-5)
        x = F.max_pool2d(F.leaky_relu(x, 0.1), 3, stride=s, padding=1)
    return x


def _stride_grid(B, oh, ow, sh, sw):
    hh, ww = np.meshgrid(np.arange(oh) * sh, np.arange(ow) * sw, indexing="ij")
    return np.broadcast_to(np.stack([hh, ww], -1).reshape(1, -1, 2), (B, oh * ow, 2)).astype(np.int32).copy()


def _select(x1, x2, idx_n2, kernel, K, flag, dist, sh=1, sw=1):
    """U:63-103 / U:253-293 -> flat index (B,N,K) int64 into x2's pixels, mask (B,N,K,1)"""
    _, h, w, m = orc.fused_conv_select_k(x1.detach().cpu().numpy(), x2.detach().cpu().numpy(), idx_n2,
                                         np.arange(kernel[0] * kernel[1], dtype=np.int32), kernel[0], kernel[1], K,
                                         flag, dist, sh, sw)
    return torch.from_numpy(h * x2.shape[2] + w).to(x1.device), torch.from_numpy(m)[..., None].to(x1.device, x1.dtype)


def _gather(feature, flat):
    """U:36-60 -- torch.gather on (B,HW,C) with the index repeated over C"""
    B, C = feature.shape[0], feature.shape[-1]
    f = feature.reshape(B, -1, C)
    out = torch.gather(f, 1, flat.reshape(B, -1, 1).repeat(1, 1, C))
    return out.view(*flat.shape, C)


def _check_valid(xyz):
    return torch.any(torch.ne(xyz, 0), dim=-1, keepdim=True).float()


def _set_abstraction(sd, name, raw, xyz, feat, lv_in, oh, ow, sh, sw, kernel, K, dist, centre_form):
    """P:77-131 (forward) / P:133-200 (forward_center), use_trans=True, raw_feat_point=True"""
    B, H, W, _ = xyz.shape
    n = oh * ow
    new_xyz = xyz[:, ::sh, ::sw][:, :oh, :ow].reshape(B, n, 3)
    new_raw = raw[:, ::sh, ::sw][:, :oh, :ow].reshape(B, n, 3)
    flat, _ = _select(xyz, xyz, _stride_grid(B, oh, ow, sh, sw), kernel, K, 3, dist)
    grouped_xyz = _gather(raw, flat)
    norm = grouped_xyz - new_raw.view(B, n, 1, 3)
    if centre_form:
        pts = torch.cat([norm, new_xyz.view(B, n, 1, 3).repeat(1, 1, K, 1), grouped_xyz,
                         torch.norm(norm, p=2, dim=3).unsqueeze(3)], -1)
    else:
        pts = torch.cat([norm, _gather(feat, flat)], -1)
    pts = _mlp(sd, name + ".mlp_convs", _count(sd, name + ".mlp_convs"), pts, leaky=False)
    return new_raw.view(B, oh, ow, 3), new_xyz.view(B, oh, ow, 3), torch.max(pts, dim=2)[0].view(B, oh, ow, -1)


def _standardise(x):
    return (x - torch.mean(x, -1, keepdim=True)) / torch.clip(torch.std(x, -1, keepdim=True), min=1e-12)


def _knn(k, xyz, new_xyz):
    """U:344-380 -- the full distance matrix, then topk"""
    d = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    d += torch.sum(new_xyz ** 2, -1).view(*new_xyz.shape[:2], 1)
    d += torch.sum(xyz ** 2, -1).view(xyz.shape[0], 1, xyz.shape[1])
    return torch.topk(d, k, dim=-1, largest=False, sorted=False)[1]


def _cost_volume(sd, name, cfg, which, H, W, warped_xyz, warped_points, f2_xyz, f2_points, lidar_z):
    """P:354-490"""
    B, N, _ = warped_xyz.shape
    nq, K2 = cfg.cost_volume_nsamples[1][which], cfg.cost_volume_nsamples[0]
    if nq > 0:
        idx = _knn(nq, f2_xyz, warped_xyz)
        qi_xyz, qi_pts = _gather(f2_xyz, idx), _gather(f2_points, idx)
    else:
        qi_xyz = f2_xyz.unsqueeze(1).repeat(1, N, 1, 1)
        qi_pts = f2_points.unsqueeze(1).repeat(1, N, 1, 1)
    warped_xyz = warped_xyz.mul(lidar_z)
    K = qi_xyz.shape[2]
    pi_xyz = warped_xyz[:, :, None, :].repeat(1, 1, K, 1)
    pi_pts = _standardise(warped_points[:, :, None, :].repeat(1, 1, K, 1))
    qi_pts = _standardise(qi_pts)
    xyz6 = torch.cat([pi_xyz, qi_xyz], dim=3)
    corr = pi_pts * qi_pts
    feat = torch.cat([xyz6, corr], dim=3)
    if cfg.backward_validation[which]:
        valid = _check_valid(warped_xyz).unsqueeze(-1)
        masked = corr * valid + -1e10 * (1 - valid)
        feat = torch.cat([feat, torch.max(masked, 1, keepdim=True)[0].repeat(1, N, 1, 1)], dim=-1)
    feat = _mlp(sd, name + ".mlp1_convs", 3, feat)
    w = _mlp(sd, name + ".mlp2_convs", 2, torch.cat([_conv(sd, name + ".pi_encoding", xyz6), feat], dim=3))
    feat = torch.sum(F.softmax(w, dim=2) * feat, dim=2)                      # B,N,64

    bhw = warped_xyz.view(B, H, W, 3)
    flat, vmask = _select(bhw, bhw, _stride_grid(B, H, W, 1, 1), cfg.cost_volume_kernel_size[which], K2, 2,
                          cfg.cost_volume_dis[which])
    g_xyz, g_pts = _gather(bhw, flat), _gather(feat, flat)
    c_xyz = warped_xyz[:, :, None, :].repeat(1, 1, K2, 1)
    c_pts = warped_points[:, :, None, :].repeat(1, 1, K2, 1)
    diff = g_xyz - c_xyz
    euc = torch.sqrt(torch.sum(diff * diff, dim=3, keepdim=True) + 1e-20)
    enc = _conv(sd, name + ".pc_encoding", torch.cat([c_xyz, g_xyz, diff, euc], dim=3))
    w = _mlp(sd, name + ".mlp2_convs_2", 2, torch.cat([enc, c_pts, g_pts], dim=-1))
    w = w * vmask + -1e10 * (1 - vmask)
    return torch.sum(F.softmax(w, dim=2) * g_pts, dim=2).view(B, H, W, -1)


def _upconv(sd, name, which, cfg, xyz1, xyz2, raw1, raw2, feat1, feat2):
    """P:241-295"""
    B, oh, ow, _ = xyz1.shape
    flat, _ = _select(xyz1, xyz2, _stride_grid(B, oh, ow, 1, 1), cfg.up_conv_kernel_size[which],
                      cfg.setupconv_nsamples[which], 3, cfg.up_conv_dis[which], cfg.stride_Hs[-1], cfg.stride_Ws[-1])
    diff = _gather(raw2, flat) - raw1.view(B, oh * ow, 1, 3)
    up = _mlp(sd, name + ".mlp_conv", 2, torch.cat([_gather(feat2, flat), diff], dim=3))
    new = torch.cat([torch.max(up, dim=2)[0].view(B, oh, ow, -1), feat1], dim=3)
    return _mlp(sd, name + ".mlp2_conv", 1, new).reshape(B, oh * ow, -1)


def _flow(sd, name, parts):
    return _mlp(sd, name + ".mlp_conv", 2, torch.cat(parts, -1).unsqueeze(2)).squeeze(2)


def _head(sd, name, prediction, mask):
    """P:528-563 with the dropout silenced (p = 0 in the parity runs)"""
    mask_p = F.softmax(mask, dim=1)
    g = torch.sum(prediction * mask_p, dim=1, keepdim=True).permute(0, 2, 1)
    lin = lambda p, x: F.conv1d(x, sd["%s.%s.composed_module.0.weight" % (name, p)],
                                sd["%s.%s.composed_module.0.bias" % (name, p)])
    hid = lin("hidden_layer", g)
    q, t = lin("quat_head", hid).squeeze(-1), lin("trans_head", hid).squeeze(-1)
    return q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10), t


def _mul_q(a, b):
    """W:25-57"""
    a = a.unsqueeze(1) if a.ndim == 2 else a
    b = b.unsqueeze(1) if b.ndim == 2 else b
    return torch.stack([
        a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
        a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
        a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
        a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]], dim=-1)


def _inv_q(q):
    q = q.reshape(q.shape[0], 4)
    return torch.cat([q[:, :1], -q[:, 1:]], dim=-1) / (torch.sum(q * q, dim=-1, keepdim=True) + 1e-10)


def forward(sd, rgb, lidar, lidar_raw, intrinsic, lidar_feature=None, cfg=KittiShape, intermediates=None):
    """M:216-424.  sd: reference-layout state_dict of CPU tensors -> out_3 (B,7), result_4 (B,7)"""
    B, N = rgb.shape[0], lidar.shape[1]
    keep = (lambda k, v: intermediates.__setitem__(k, v.detach())) if intermediates is not None else (lambda k, v: None)
    RF = rgb
    for i in range(3):
        RF = _rgb(sd, "RGB_net%d" % (i + 1), RF, cfg.rgb_pool_strides[i])
    dev = rgb.device   # CPU in every timed use; the GPU parity test runs the same code on cuda tensors
    feat0 = torch.zeros(B, N, 3) if lidar_feature is None else lidar_feature.cpu()
    raw_img, (_, cam_img) = orc.project_seq(lidar_raw.cpu().numpy(), [feat0.numpy(), lidar.cpu().numpy()], cfg.init_H,
                                            cfg.init_W, cfg.fup, cfg.fdown)                # M:247, U:111-187
    dt = rgb.dtype     # f32 everywhere except the f64 "truth" run of the GPU gradient test
    raw_l, xyz_l, feat_l = torch.from_numpy(raw_img).to(dev, dt), torch.from_numpy(cam_img).to(dev, dt), None
    Hs = [int(np.ceil(cfg.init_H / s)) for s in np.cumprod(cfg.stride_Hs)]
    Ws = [int(np.ceil(cfg.init_W / s)) for s in np.cumprod(cfg.stride_Ws)]
    levels = []
    for lv in range(4):                                                                    # M:256-259
        raw_l, xyz_l, feat_l = _set_abstraction(sd, "LiDAR_lv%d" % (lv + 1), raw_l, xyz_l, feat_l, lv, Hs[lv], Ws[lv],
                                                cfg.stride_Hs[lv], cfg.stride_Ws[lv], cfg.kernel_sizes[lv],
                                                cfg.lidar_group_samples[lv], cfg.down_conv_dis[lv], lv == 0)
        levels.append((raw_l, xyz_l, feat_l))
    (P3_raw, P3, LF3), (P4_raw, P4, LF4) = levels[2], levels[3]
    keep("LiDAR_lv2", levels[1][2]), keep("LiDAR_lv3", LF3)
    H3, W3, H4, W4 = Hs[2], Ws[2], Hs[3], Ws[3]

    K3 = intrinsic.to(dt).clone()                                                          # M:443-449
    h, w = RF.shape[2:]
    K3[:, 0, 0] *= w / rgb.shape[3]; K3[:, 0, 2] *= w / rgb.shape[3]
    K3[:, 1, 1] *= h / rgb.shape[2]; K3[:, 1, 2] *= h / rgb.shape[2]
    jj, ii = torch.meshgrid(torch.arange(w, dtype=dt, device=dev), torch.arange(h, dtype=dt, device=dev),
                            indexing="xy")
    pix = torch.stack([jj, ii, torch.ones_like(jj)], -1).reshape(1, -1, 3).repeat(B, 1, 1)
    RF_index = torch.bmm(torch.inverse(K3), pix.permute(0, 2, 1)).permute(0, 2, 1)         # M:282-284
    RF3 = RF.reshape(B, RF.shape[1], -1).permute(0, 2, 1)

    P3_l4, LF3_cv = P3.reshape(B, H3 * W3, 3), LF3.reshape(B, H3 * W3, -1)
    z = P3_l4[:, :, 2:]
    cv1 = _cost_volume(sd, "cost_volume1", cfg, 0, H3, W3, P3_l4 / (z + 1e-10), LF3_cv, RF_index, RF3, z)
    keep("cost_volume1", cv1)
    _, _, l4_pred = _set_abstraction(sd, "layer_idx", P3_raw, P3, cv1, 3, H4, W4, cfg.stride_Hs[3], cfg.stride_Ws[3],
                                     cfg.kernel_sizes[3], cfg.lidar_group_samples[4], cfg.down_conv_dis[3], False)
    keep("layer_idx", l4_pred)
    l4_pred = l4_pred.view(B, H4 * W4, -1)
    v4 = _check_valid(P4_raw).view(B, -1, 1)
    l4_w = _flow(sd, "flow_predictor0", [LF4.view(B, H4 * W4, -1), l4_pred])
    l4_w = l4_w * v4 + -1e10 * (1 - v4)
    q4, t4 = _head(sd, "l4_head", l4_pred, l4_w)

    zero = torch.zeros(B, 1, device=dev, dtype=dt)
    homo = torch.cat([torch.zeros(B, H3 * W3, 1, device=dev, dtype=dt), P3_l4], -1)        # W:78-94
    homo = _mul_q(_mul_q(q4, homo), _inv_q(q4)) + torch.cat([zero, t4], -1).reshape(B, 1, 4)
    P3_w = homo[:, :, 1:4] * _check_valid(P3_l4)
    w_up = _upconv(sd, "set_upconv0_w_upsample", 0, cfg, P3, P4, P3_raw, P4_raw, LF3, l4_w.view(B, H4, W4, -1))
    e_up = _upconv(sd, "set_upconv0_upsample", 1, cfg, P3, P4, P3_raw, P4_raw, LF3, l4_pred.view(B, H4, W4, -1))
    keep("set_upconv0_upsample", e_up)
    z = P3_w[:, :, 2:]
    cv2 = _cost_volume(sd, "cost_volume2", cfg, 1, H3, W3, P3_w / (z + 1e-10), LF3_cv, RF_index, RF3, z)
    keep("cost_volume2", cv2)
    l3_pred = _flow(sd, "flow_predictor0_predict", [LF3_cv, cv2.view(B, H3 * W3, -1), e_up])
    l3_w = _flow(sd, "flow_predictor0_w", [LF3_cv, l3_pred, w_up])
    v3 = _check_valid(P3_raw).view(B, -1, 1)
    q3, t3 = _head(sd, "l3_head", l3_pred, l3_w * v3 + -1e10 * (1 - v3))

    q = _mul_q(q3.view(B, 1, 4), q4.view(B, 1, 4)).squeeze(1)                              # M:388-404
    t = (_mul_q(_mul_q(q3, torch.cat([zero, t4], 1).view(B, 1, 4)), _inv_q(q3)) +
         torch.cat([zero, t3], 1).view(B, 1, 4)).squeeze(1)
    return torch.cat([q, t[:, 1:]], 1), torch.cat([q4, t4], 1)


def loss_fn(out3, out4, q_gt, t_gt, sx, sq):
    """compute_loss.py:102-133 with l1_trans_loss"""
    def terms(o):
        lq = torch.mean(torch.sqrt(torch.sum((q_gt - o[:, :4]) ** 2, dim=-1, keepdim=True) + 1e-10))
        return lq, F.l1_loss(o[:, 4:], t_gt)
    (q3, x3), (q4, x4) = terms(out3), terms(out4)
    l3 = x3 * torch.exp(-sx) + sx + q3 * torch.exp(-sq) + sq
    l4 = x4 * torch.exp(-sx) + sx + q4 * torch.exp(-sq) + sq
    return 1.6 * l4 + 0.8 * l3


def random_state(seed=0):
    """A reference-layout state_dict with default-style random initialisation (no checkpoints
    exist offline): shapes from the KITTI config."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cin, cout, k, bn, dims):
        bound = 1.0 / np.sqrt(cin * k ** dims)
        shape = (cout, cin) + (k,) * dims
        sd[name[0]] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name[1]] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        if bn:
            sd[bn + ".weight"], sd[bn + ".bias"] = torch.ones(cout), torch.zeros(cout)

    def pw(prefix, cin, cout):
        conv((prefix + ".conv.weight", prefix + ".conv.bias"), cin, cout, 1, prefix + ".bn_linear", 2)

    def mlp(prefix, cin, chans):
        for i, c in enumerate(chans):
            pw("%s.%d" % (prefix, i), cin, c)
            cin = c

    enc = [[16, 16, 32], [32, 32, 64], [64, 64, 128], [128, 128, 256], [128, 64, 64]]
    for lv, cin in enumerate([10, 35, 67, 131]):
        mlp("LiDAR_lv%d.mlp_convs" % (lv + 1), cin, enc[lv])
    mlp("layer_idx.mlp_convs", 67, enc[4])
    for i, (cin, chans) in enumerate([(3, [16, 16, 16, 16, 32]), (32, [32, 32, 32, 32, 64]), (64, [64, 64, 64, 64, 128])]):
        for j, c in enumerate(chans):
            n = "RGB_net%d.%d" % (i + 1, 4 * j)
            conv((n + ".weight", n + ".bias"), cin, c, 3, "RGB_net%d.%d" % (i + 1, 4 * j + 1), 2)
            cin = c
    for i, cin in enumerate([262, 134]):
        n = "cost_volume%d" % (i + 1)
        mlp(n + ".mlp1_convs", cin, [128, 64, 64])
        pw(n + ".pi_encoding", 6, 64)
        mlp(n + ".mlp2_convs", 128, [128, 64])
        pw(n + ".pc_encoding", 10, 64)
        mlp(n + ".mlp2_convs_2", 256, [128, 64])
    mlp("flow_predictor0.mlp_conv", 320, [128, 64])
    for n in ("set_upconv0_w_upsample", "set_upconv0_upsample"):
        mlp(n + ".mlp_conv", 67, [128, 64])
        mlp(n + ".mlp2_conv", 192, [64])
    mlp("flow_predictor0_predict.mlp_conv", 256, [128, 64])
    mlp("flow_predictor0_w.mlp_conv", 256, [128, 64])
    for h in ("l4_head", "l3_head"):
        for part, (cin, cout) in dict(hidden_layer=(64, 256), quat_head=(256, 4), trans_head=(256, 3)).items():
            n = "%s.%s.composed_module.0" % (h, part)
            conv((n + ".weight", n + ".bias"), cin, cout, 1, None, 1)
    sd["sq"], sd["sx"] = torch.tensor([-2.5]), torch.tensor([0.0])
    return sd
