"""GPU: the tensor-core shared-MLP GEMMs (csrc/mlp_tc.cu: forward, dX, dW as tcgen05 3xTF32) called through
the C ABI, kernel by kernel, against an f64 evaluation of the same layer algebra and against the f32
FMA kernels (csrc/mlp.cu) on the same inputs.  Shapes are the model's layers plus ragged ones (partial
row tiles, channel counts that are not multiples of the tile, 4- / 8- / 16-byte aligned rows)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# rows, cin, cout, input is a normalised layer (fused transform), previous-layer sums wanted
SHAPES = [
    (36480, 262, 128, False, False),    # cost volume 1, layer 1 (cin % 4 == 2: 8-byte rows)
    (36480, 128, 64, True, True),       # cost volume 1, layer 2
    (36480, 64, 64, True, True),
    (14592, 134, 128, False, False),    # cost volume 2, layer 1
    (3648, 67, 128, False, False),      # up-conv (odd cin: 4-byte rows)
    (3712, 131, 128, False, False),
    (3712, 128, 256, True, True),       # two output-channel tiles
    (300, 192, 64, True, True),         # partial row tile
    (1000, 320, 128, False, False),     # flow predictor
    (5000, 6, 64, False, False),        # position encoding: forward only on the tensor cores
]


@pytest.fixture(autouse=True, params=[7, 23], ids=["noswizzle", "sw128"])
def operand_layout(request):
    """Every kernel-level test runs with both operand layouts of the forward / dX kernels."""
    from i2pnet_b200 import _cabi
    L = _cabi.lib()
    before = L.i2p_get_mlp_tensor_cores()
    L.i2p_set_mlp_tensor_cores(request.param)
    yield request.param
    L.i2p_set_mlp_tensor_cores(before)


def _ids(s):
    return "r%d_%dto%d%s" % (s[0], s[1], s[2], "_bn" if s[3] else "")


def _act(z, slope):
    return torch.where(z > 0, z, z * slope)


class _Layer:
    """One layer's tensors in f32 on the device and its algebra in f64."""

    def __init__(self, rows, cin, cout, has_tf, seed, slope=0.1, in_slope=0.1):
        dev = torch.device("cuda:0")
        g = torch.Generator(device=dev).manual_seed(seed)
        r = lambda *s: torch.randn(*s, generator=g, device=dev)
        self.rows, self.cin, self.cout, self.has_tf, self.slope, self.in_slope = rows, cin, cout, has_tf, slope, in_slope
        self.x = r(rows, cin) * 2 + 0.5
        self.w = r(cout, cin) * 0.2
        self.b = r(cout)
        self.gamma = torch.rand(cout, generator=g, device=dev) + 0.5
        self.beta = r(cout) * 0.3
        if has_tf:     # the input is a raw previous-layer output with its own batch-norm constants
            xd = self.x.double()
            mean, var = xd.mean(0), xd.var(0, unbiased=False)
            pg, pb = torch.rand(cin, generator=g, device=dev).double() + 0.5, r(cin).double() * 0.3
            rstd = 1 / torch.sqrt(var + 1e-5)
            self.prev = torch.stack([mean, rstd, pg * rstd, pb - mean * pg * rstd]).float().contiguous()
            self.xin = _act(xd * self.prev[2].double() + self.prev[3].double(), in_slope)
        else:
            self.prev = None
            self.xin = self.x.double()
        self.y64 = self.xin @ self.w.double().t() + self.b.double()
        self.y = self.y64.float().contiguous()
        yd = self.y.double()
        mean, var = yd.mean(0), yd.var(0, unbiased=False)
        rstd = 1 / torch.sqrt(var + 1e-5)
        self.st = torch.stack([mean, rstd, self.gamma.double() * rstd, self.beta.double() - mean * self.gamma.double() * rstd]
                              ).float().contiguous()
        self.g = r(rows, cout)
        # backward algebra in f64 from the f32 tensors the kernels see
        st = self.st.double()
        z = yd * st[2] + st[3]
        dz = self.g.double() * torch.where(z > 0, 1.0, slope)
        yhat = (yd - st[0]) * st[1]
        self.s12 = torch.stack([dz.sum(0), (dz * yhat).sum(0)]).contiguous()
        self.dy = st[2] * (dz - self.s12[0] / rows - yhat * self.s12[1] / rows)
        self.dx64 = self.dy @ self.w.double()
        self.dw64 = self.dy.t() @ self.xin
        if has_tf:
            p = self.prev.double()
            pz = self.x.double() * p[2] + p[3]
            pdz = self.dx64 * torch.where(pz > 0, 1.0, in_slope)
            self.prev_s12 = torch.stack([pdz.sum(0), (pdz * ((self.x.double() - p[0]) * p[1])).sum(0)])


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _bn_args(st, slope):
    return (st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), float(slope))


@pytest.mark.parametrize("shape", SHAPES, ids=_ids)
def test_forward_tc_matches_f64(shape):
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    rows, cin, cout, has_tf, _ = shape
    L = _cabi.lib()
    assert L.i2p_pw_tc_supported(0, rows, cin, cout)
    lay = _Layer(rows, cin, cout, has_tf, seed=rows + cin)
    dev = lay.x.device
    pack = torch.empty(L.i2p_pw_pack_floats(cin, cout), device=dev)
    call("i2p_pw_pack_weights", dev, cin, cout, lay.w.data_ptr(), pack.data_ptr())
    ntiles = L.i2p_pw_num_tiles(rows)
    res = {}
    for kind in ("tc", "fma"):
        y = torch.full((rows, cout), float("nan"), device=dev)
        tiles = torch.zeros(ntiles, cout, 2, device=dev)
        sc = lay.prev[2].data_ptr() if has_tf else None
        sh = lay.prev[3].data_ptr() if has_tf else None
        if kind == "tc":
            call("i2p_pw_linear_fwd_tc", dev, rows, cin, cout, lay.x.data_ptr(), sc, sh, lay.in_slope, pack.data_ptr(),
                 lay.b.data_ptr(), y.data_ptr(), tiles.data_ptr())
        else:
            before = L.i2p_get_mlp_tensor_cores()
            L.i2p_set_mlp_tensor_cores(0)        # (the FMA kernel inside i2p_pw_linear_fwd; mask bit 8 would pick tcgen05 v1)
            call("i2p_pw_linear_fwd", dev, rows, cin, cout, lay.x.data_ptr(), sc, sh, lay.in_slope, lay.w.data_ptr(),
                 lay.b.data_ptr(), y.data_ptr(), tiles.data_ptr())
            L.i2p_set_mlp_tensor_cores(before)
        st = torch.empty(4, cout, device=dev)
        call("i2p_bn_finalize", dev, rows, cout, tiles.data_ptr(), lay.gamma.data_ptr(), lay.beta.data_ptr(), 1e-5,
             st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr())
        res[kind] = (y, st)
    e_tc, e_fma = _rel(res["tc"][0], lay.y64), _rel(res["fma"][0], lay.y64)
    print("forward  max-rel error vs f64: tcgen05 %.2e   fma %.2e" % (e_tc, e_fma))
    # the tensor core adds with truncation: ~1e-6 of the largest output at K = 262, four times an FMA chain
    assert e_tc < max(4e-6, 4 * e_fma), (e_tc, e_fma)
    yd = res["tc"][0].double()
    assert _rel(res["tc"][1][0], yd.mean(0)) < 1e-5
    assert _rel(res["tc"][1][1], 1 / torch.sqrt(yd.var(0, unbiased=False) + 1e-5)) < 1e-5


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[1] >= 32], ids=_ids)
def test_backward_tc_matches_f64(shape):
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    rows, cin, cout, has_tf, want_prev = shape
    L = _cabi.lib()
    assert L.i2p_pw_tc_supported(1, rows, cin, cout) and L.i2p_pw_tc_supported(2, rows, cin, cout)
    lay = _Layer(rows, cin, cout, has_tf, seed=rows + cin + 1)
    dev = lay.x.device
    pack = torch.empty(L.i2p_pw_pack_floats(cin, cout), device=dev)
    call("i2p_pw_pack_weights", dev, cin, cout, lay.w.data_ptr(), pack.data_ptr())
    s12 = lay.s12.clone()
    bn = _bn_args(lay.st, lay.slope)
    prev = (lay.x.data_ptr(), *_bn_args(lay.prev, lay.in_slope)) if has_tf else (None, None, None, None, None, 1.0)
    psc = lay.prev[2].data_ptr() if has_tf else None
    psh = lay.prev[3].data_ptr() if has_tf else None
    out = {}
    for kind in ("tc", "fma"):
        dx = torch.full((rows, cin), float("nan"), device=dev)
        dw = torch.zeros(cout, cin, device=dev)
        ps12 = torch.zeros(2, cin, dtype=torch.float64, device=dev)
        if kind == "tc":
            call("i2p_pw_linear_bwd_dx_tc", dev, rows, cin, cout, lay.g.data_ptr(), None, None, 1, lay.y.data_ptr(), *bn, s12.data_ptr(),
                 pack.data_ptr(), dx.data_ptr(), *prev, ps12.data_ptr() if has_tf else None)
            call("i2p_pw_linear_bwd_dw_tc", dev, rows, cin, cout, lay.g.data_ptr(), None, None, 1, lay.y.data_ptr(), *bn, s12.data_ptr(),
                 lay.x.data_ptr(), psc, psh, lay.in_slope if has_tf else 1.0, dw.data_ptr())
        else:
            call("i2p_pw_linear_bwd_dx", dev, rows, cin, cout, lay.g.data_ptr(), None, None, 1, lay.y.data_ptr(), *bn,
                 s12.data_ptr(), lay.w.data_ptr(), dx.data_ptr(), *prev, ps12.data_ptr() if has_tf else None)
            call("i2p_pw_linear_bwd_dw", dev, rows, cin, cout, lay.g.data_ptr(), None, None, 1, lay.y.data_ptr(), *bn,
                 s12.data_ptr(), lay.x.data_ptr(), psc, psh, lay.in_slope if has_tf else 1.0, dw.data_ptr())
        out[kind] = (dx, dw, ps12)
    rep = []
    ok = True
    for name, i, truth in (("dx", 0, lay.dx64), ("dw", 1, lay.dw64)) + ((("prev_s12", 2, lay.prev_s12),) if has_tf and want_prev else ()):
        e_tc, e_fma = _rel(out["tc"][i], truth), _rel(out["fma"][i], truth)
        good = e_tc < max(4e-6 if name != "prev_s12" else 1e-5, 4 * e_fma)
        ok &= good
        rep.append("%s %-8s max-rel error vs f64: tcgen05 %.2e   fma %.2e" % ("ok  " if good else "FAIL", name, e_tc, e_fma))
    print("\n".join(rep))
    assert ok, "\n" + "\n".join(rep)


def test_pack_layout_round_trips(operand_layout):
    """The pack kernel against a host restatement of the UMMA canonical K-major layouts (no swizzle / SWIZZLE_128B)."""
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    L = _cabi.lib()
    swz = bool(operand_layout & 16)
    dev = torch.device("cuda:0")
    cin, cout = 70, 96
    w = torch.randn(cout, cin, device=dev)
    n = L.i2p_pw_pack_floats(cin, cout)
    pack = torch.full((n,), float("nan"), device=dev)
    call("i2p_pw_pack_weights", dev, cin, cout, w.data_ptr(), pack.data_ptr())
    p = pack.cpu()
    assert torch.isfinite(p).all()
    wc = w.cpu()
    bn, nc = 128, 3            # forward section: cout 96 -> one 128-wide tile, cin 70 -> 3 chunks of 32
    for (nn, k) in [(0, 0), (5, 3), (95, 69), (17, 33), (96, 0), (0, 70)]:
        c, kl = divmod(k, 32)
        inner = nn * 32 + (((kl // 4) ^ (nn % 8)) * 4) + kl % 4 if swz else (kl // 4) * bn * 4 + nn * 4 + kl % 4
        off = (0 * nc + c) * 2 * bn * 32 + inner
        want = float(wc[nn, k]) if nn < cout and k < cin else 0.0
        hi, lo = float(p[off]), float(p[off + bn * 32])
        assert abs(hi + lo - want) <= 5e-7 * abs(want) + 1e-12, (nn, k, hi, lo, want)


@pytest.mark.parametrize("shape", [(29184, 64, 128, 16), (3712, 128, 256, 16), (14592, 128, 64, 8), (2400, 67, 64, 8)],
                         ids=lambda s: "r%d_%dto%d_k%d" % s)
def test_backward_tc_max_over_k_source_matches_fma(shape):
    """Gradient arriving through a max-over-K output (dout + arg-max): tensor-core dX / dW against the FMA kernels."""
    from i2pnet_b200 import _cabi
    from i2pnet_b200._cabi import call
    rows, cin, cout, K = shape
    L = _cabi.lib()
    lay = _Layer(rows, cin, cout, False, seed=rows + K)
    dev = lay.x.device
    gen = torch.Generator(device=dev).manual_seed(7)
    dout = torch.randn(rows // K, cout, device=dev, generator=gen)
    arg = torch.randint(0, K, (rows // K, cout), device=dev, generator=gen, dtype=torch.int32)
    pack = torch.empty(L.i2p_pw_pack_floats(cin, cout), device=dev)
    call("i2p_pw_pack_weights", dev, cin, cout, lay.w.data_ptr(), pack.data_ptr())
    bn = _bn_args(lay.st, lay.slope)
    s12 = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    call("i2p_bn_bwd_reduce", dev, rows, cout, None, dout.data_ptr(), arg.data_ptr(), K, lay.y.data_ptr(), *bn, s12.data_ptr())
    none_prev = (None, None, None, None, None, 1.0, None)
    out = {}
    for kind in ("tc", "fma"):
        dx = torch.full((rows, cin), float("nan"), device=dev)
        dw = torch.zeros(cout, cin, device=dev)
        sfx = "_tc" if kind == "tc" else ""
        src = (None, dout.data_ptr(), arg.data_ptr(), K)
        call("i2p_pw_linear_bwd_dx" + sfx, dev, rows, cin, cout, *src, lay.y.data_ptr(), *bn, s12.data_ptr(),
             pack.data_ptr() if kind == "tc" else lay.w.data_ptr(), dx.data_ptr(), *none_prev)
        call("i2p_pw_linear_bwd_dw" + sfx, dev, rows, cin, cout, *src, lay.y.data_ptr(), *bn, s12.data_ptr(), lay.x.data_ptr(),
             None, None, 1.0, dw.data_ptr())
        out[kind] = (dx, dw)
    for name, i in (("dx", 0), ("dw", 1)):
        e = _rel(out["tc"][i], out["fma"][i].double())
        assert e < 1e-5, (name, e)
