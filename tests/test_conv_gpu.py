"""GPU: the 3x3 convolution kernels of csrc/conv.cu through the C ABI -- implicit GEMM on tcgen05 (3xTF32) for the
forward and the data gradient, f32 FMA weight gradient -- against torch's convolution evaluated in f64
(src/modules/basicConv.py:11: nn.Conv2d(k = 3, stride 1, padding 1)).  Bars: forward / data gradient 8e-6 of the
largest output (measured: <= 4.2e-6 at K = 9 * 64 = 576 with non-centred inputs, 1-2e-6 at K = 144; the tensor core
accumulates with truncation, 216 accumulation steps at K = 576; north_star's feature tolerance is 1e-4), weight
gradient 1e-5 (sums over up to 1.6e5 positions), batch statistics from the epilogue 1e-5."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (B, cin, cout, H, W): the fifteen layers of the KITTI pyramid (batch 2), the nuScenes first layer, odd shapes
SHAPES = [(2, 3, 16, 160, 512), (2, 16, 16, 80, 256), (2, 16, 32, 80, 256), (2, 32, 32, 40, 128), (2, 32, 32, 20, 64),
          (2, 32, 64, 20, 64), (2, 64, 64, 10, 32), (2, 64, 128, 10, 32), (1, 3, 16, 320, 640),
          (1, 5, 20, 7, 13), (3, 17, 7, 9, 130), (2, 40, 33, 3, 5), (1, 1, 1, 1, 1), (1, 8, 16, 2, 127)]


def _ids(s):
    return "b%d_%dto%d_%dx%d" % s


def _conv_tc(x, w, bias, dgrad=False, stats=False):
    from i2pnet_b200 import _cabi
    L = _cabi.lib()
    B, ki, H, W = x.shape
    cout, cin = w.shape[:2]
    no = cin if dgrad else cout
    pack = torch.empty(L.i2p_conv3x3_pack_floats(cin, cout, int(dgrad)), device=x.device)
    _cabi.call("i2p_conv3x3_pack", x.device, cin, cout, int(dgrad), w.data_ptr(), pack.data_ptr())
    y = torch.full((B, no, H, W), float("nan"), device=x.device)
    tiles = torch.full((no, L.i2p_conv3x3_stat_slots(B, no, H, W), 3), float("nan"), device=x.device) if stats else None
    _cabi.call("i2p_conv3x3_tc", x.device, B, ki, no, H, W, x.data_ptr(), pack.data_ptr(),
               bias.data_ptr() if bias is not None else None, y.data_ptr(), tiles.data_ptr() if stats else None)
    return y, tiles


@pytest.mark.parametrize("shape", SHAPES, ids=_ids)
def test_conv3x3_forward_and_statistics(shape):
    B, cin, cout, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = torch.randn(B, cin, H, W, device=DEV, generator=g) * 2 + 0.3
    w = torch.randn(cout, cin, 3, 3, device=DEV, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, device=DEV, generator=g)
    y, tiles = _conv_tc(x, w, bias, stats=True)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    assert torch.isfinite(y).all()
    err = float((y.double() - want).abs().max() / want.abs().max())
    lib = float((F.conv2d(x, w, bias, padding=1).double() - want).abs().max() / want.abs().max())
    print("forward %s: error vs f64 %.2e (library f32 convolution: %.2e)" % (shape, err, lib))
    assert err < 8e-6, err
    # epilogue statistics: Chan-merge the tiles like rgb_bn_finalize does
    n, mu, m2 = tiles[..., 0].double(), tiles[..., 1].double(), tiles[..., 2].double()
    N = n.sum(1)
    assert torch.all(N == B * H * W)
    mean = (n * mu).sum(1) / N
    var = (m2.sum(1) + (n * (mu - mean[:, None]) ** 2).sum(1)) / N
    wm, wv = want.mean(dim=(0, 2, 3)), want.var(dim=(0, 2, 3), unbiased=False)
    scale = want.abs().max()
    assert float((mean - wm).abs().max() / scale) < 1e-5
    assert float((var - wv).abs().max() / wv.max().clamp_min(1e-30)) < 1e-5 or B * H * W == 1


@pytest.mark.parametrize("shape", SHAPES, ids=_ids)
def test_conv3x3_data_and_weight_gradients(shape):
    from i2pnet_b200 import _cabi
    B, cin, cout, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(sum(shape) + 1)
    x = torch.randn(B, cin, H, W, device=DEV, generator=g)
    w = torch.randn(cout, cin, 3, 3, device=DEV, generator=g) * (2.0 / (9 * cin)) ** 0.5
    dy = torch.randn(B, cout, H, W, device=DEV, generator=g)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xd, wd, None, padding=1).backward(dy.double())
    dx, _ = _conv_tc(dy, w, None, dgrad=True)
    err = float((dx.double() - xd.grad).abs().max() / xd.grad.abs().max())
    assert err < 8e-6, err
    dw = torch.zeros_like(w)
    _cabi.call("i2p_conv3x3_wgrad", DEV if isinstance(DEV, torch.device) else torch.device(DEV), B, cin, cout, H, W,
               x.data_ptr(), dy.data_ptr(), dw.data_ptr())
    err = float((dw.double() - wd.grad).abs().max() / wd.grad.abs().max())
    assert err < 1e-5, err
    # it ACCUMULATES (the step engine points it at a zeroed slice of the flat gradient buffer)
    _cabi.call("i2p_conv3x3_wgrad", torch.device(DEV), B, cin, cout, H, W, x.data_ptr(), dy.data_ptr(), dw.data_ptr())
    assert float((dw.double() - 2 * wd.grad).abs().max() / wd.grad.abs().max()) < 2e-5


def test_image_branch_launches_no_library_convolution():
    """The three pyramids forward + backward under the profiler: every kernel on the device is one of ours or an ATen
    element-wise / copy kernel -- no cuDNN / implicit-GEMM / Winograd / FFT convolution (a12)."""
    from torch.profiler import ProfilerActivity, profile
    from i2pnet_b200.config_proj_lidarcenter import I2PNetConfig as cfg
    from i2pnet_b200.modules.basicConv import createCNNs
    torch.manual_seed(0)
    nets = [createCNNs(cin, ch, st).to(DEV) for cin, ch, st in cfg.rgb_encoder_channels]
    x = torch.rand(2, 3, 160, 512, device=DEV, requires_grad=True)

    def run():
        h = x
        for n in nets:
            h = n(h)
        h.square().mean().backward()
    run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    bad = [n for n in names if any(t in n.lower() for t in ("cudnn", "implicit_convolve", "wgrad_alg", "dgrad_engine", "fft2d",
                                                             "winograd", "convolve_common", "sgemm", "cutlass"))]
    assert not bad, bad
    assert any("conv3x3_tc_kernel" in n for n in names) and any("wgrad_kernel" in n for n in names), names
