"""GPU: the fused RGB pyramid block (csrc/conv.cu: tensor-core 3x3 convolution with the batch statistics in its
epilogue, data / weight gradients; csrc/rgb.cu: BatchNorm2d -> LeakyReLU(0.1) -> MaxPool2d(3, s, 1)) against the
reference formulation -- the four nn modules of src/modules/basicConv.py:11-17 through ATen / the library convolution --
in f32 and, as ground truth, in f64: outputs, running statistics, input / weight / bias gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, cin, channels, strides, H, W)
CASES = [
    # (full-resolution planes are covered per kernel in tests/test_conv_gpu.py; here the planes are kept small enough
    # that an input without near-tied pooling windows exists, see the comment in the test)
    (2, 3, [16, 16, 32], [2, 1, 2], 24, 64),       # RGB_net1 head
    (3, 32, [32, 64], [2, 2], 20, 64),             # RGB_net2
    (2, 64, [64, 128], [1, 2], 5, 16),             # RGB_net3 tail (plane of 80 elements)
    (1, 5, [7], [2], 13, 11),                      # odd sizes, odd channel count
    (2, 4, [8], [1], 1, 3),                        # degenerate height
]


def _run(net, x, dtype, fused):
    from i2pnet_b200.modules import basicConv
    net = copy.deepcopy(net).to(dtype)
    x = x.detach().clone().to(dtype).requires_grad_(True)
    basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = fused
    try:
        out = net(x)
        torch.manual_seed(1)
        g = torch.randn(out.shape, device=out.device, dtype=torch.float64).to(dtype)
        out.backward(g)
    finally:
        basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = True
    return dict(out=out.detach().double(), dx=x.grad.double(),
                grads={n: p.grad.double() for n, p in net.named_parameters()},
                buffers={n: b.double() for n, b in net.named_buffers()})


def _decision_margin(net, x):
    """Smallest margin, relative to the layer's largest activation, of the discrete decisions the backward pass routes
    gradients through, evaluated in f64: the gap between the two largest entries of every pooling window (when they are
    different positions' values) and the distance of every pre-activation from the LeakyReLU kink."""
    import torch.nn.functional as F
    net = copy.deepcopy(net).double()
    margins = []

    def pool_hook(mod, args):
        z = args[0]
        s = mod.stride if isinstance(mod.stride, int) else mod.stride[0]
        win = F.unfold(F.pad(z, (1, 1, 1, 1), value=float("-inf")), 3, stride=s).view(z.shape[0], z.shape[1], 9, -1)
        top = win.topk(2, dim=2).values
        margins.append(float((top[:, :, 0] - top[:, :, 1]).min() / z.abs().max()))

    def act_hook(mod, args):
        margins.append(float(args[0].abs().min() / args[0].abs().max()))
    hooks = [m.register_forward_pre_hook(pool_hook) for m in net if isinstance(m, torch.nn.MaxPool2d)]
    hooks += [m.register_forward_pre_hook(act_hook) for m in net if isinstance(m, torch.nn.LeakyReLU)]
    from i2pnet_b200.modules import basicConv
    basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = False
    try:
        with torch.no_grad():
            net(x.double())
    finally:
        basicConv.USE_FUSED_RGB_TAIL = basicConv.USE_OWN_CONV = True
        for h in hooks:
            h.remove()
    return min(margins)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "b%d_c%s_%dx%d" % (c[0], "x".join(map(str, c[2])), c[4], c[5]))
@pytest.mark.parametrize("training", [True, False], ids=["train", "eval"])
def test_block_tail_matches_aten_and_f64(case, training):
    from i2pnet_b200 import _cabi
    from i2pnet_b200.modules.basicConv import createCNNs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, cin, chans, strides, H, W = case
    dev = torch.device("cuda:0")
    torch.manual_seed(sum(chans) + H)
    net = createCNNs(cin, chans, strides).to(dev)
    with torch.no_grad():
        for m in net:
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.5, 0.5)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 2.0)
    net.train(training)
    # Inputs whose f64 evaluation keeps every discrete decision of the backward pass (pooling arg-max, LeakyReLU side)
    # at least 1e-6 of the layer's range away from a tie: two correct f32 evaluations (forward error 2e-7 ... 1e-6 here)
    # then route gradients identically, and the comparison below measures arithmetic.  A single flipped arg-max moves
    # dx by 5e-3 and every weight gradient by 1e-3 in relative L2 (one term of a random-walk sum), and which of two
    # f32 implementations flips against f64 is luck (measured both ways round, tools/debug_conv_chain.py).
    for attempt in range(64):
        x = torch.rand(B, cin, H, W, device=dev) * 4 - 1
        if _decision_margin(net, x) > 1e-6:
            break
    else:
        pytest.fail("no tie-free input found in 64 draws")
    before = _cabi.launch_count()
    fused = _run(net, x, torch.float32, True)
    assert _cabi.launch_count() - before >= 7 * len(chans), "the fused kernels did not run"
    aten = _run(net, x, torch.float32, False)
    truth = _run(net, x, torch.float64, False)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    l2 = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(fused["out"], truth["out"]) < max(1e-5, 3 * rel(aten["out"], truth["out"]))
    for n, b in truth["buffers"].items():
        assert rel(fused["buffers"][n], b) < 1e-5, n
    # gradients route through arg-max positions and LeakyReLU slopes that flip on 1e-7 forward differences:
    # compare in relative L2 against what the ATen formulation itself achieves in f32
    report = []
    for name, g, ga, gt in [("dx", fused["dx"], aten["dx"], truth["dx"])] + [
            (n, fused["grads"][n], aten["grads"][n], truth["grads"][n]) for n in truth["grads"]]:
        e, ea = l2(g, gt), l2(ga, gt)
        report.append("%-12s fused %.2e  aten %.2e" % (name, e, ea))
        assert e < max(1e-4, 3 * ea), "\n".join(report)
    print("\n".join(report))


def test_pyramid_keeps_reference_state_dict_keys():
    from i2pnet_b200.modules.basicConv import createCNNs
    net = createCNNs(3, [16, 32], [2, 2])
    keys = list(net.state_dict().keys())
    assert keys[:7] == ["0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var",
                        "1.num_batches_tracked"]
    assert "4.weight" in keys and "5.running_var" in keys
