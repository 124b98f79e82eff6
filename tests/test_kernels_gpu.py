"""Per-kernel ("teacher-forced") parity on the GPU: every kernel gets bf16-rounded oracle inputs and is compared
with the oracle result rounded at the same storage point (SURVEY 8c / App. F protocol).
Tolerances: rel-L2 <= 1e-3 for bf16 tensor-core kernels (fp32 accumulate), bit-exact for indices."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from mimo_unet_b200 import _lib
from oracle import mimo_oracle as O
from tests.util import act_of, bf16r, get_nchw, make_buffer, p8, put_nchw, rel_l2, stream

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def lib():
    l = _lib.lib()
    _lib.check(l.mimo_check_device(), "check_device")
    return l


def _pack_weights(lib, w):
    cout, cin = w.shape[0], w.shape[1]
    wf = torch.zeros(9, cout, p8(cin), dtype=torch.bfloat16, device="cuda")
    wd = torch.zeros(9, cin, p8(cout), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.mimo_weight_pack(w.data_ptr(), cout, cin, wf.data_ptr(), p8(cin), wd.data_ptr(), p8(cout), stream()))
    return wf, wd


CONV_CASES = [
    # N, Cin, Cout, H, W
    (2, 3, 21, 32, 32),
    (2, 21, 21, 37, 45),
    (3, 63, 31, 16, 24),
    (2, 84, 168, 16, 20),
    (1, 168, 336, 8, 10),
    (2, 672, 336, 8, 10),
    (4, 64, 64, 64, 64),
    (2, 30, 45, 2, 3),
    # first-layer shapes (<= 4 input, <= 32 output channels): the CUDA-core kernels of conv_thin.cu (fprop, wgrad)
    (3, 3, 21, 33, 47),
    (1, 4, 32, 9, 11),
    (2, 1, 8, 5, 7),
]


@pytest.mark.parametrize("N,Cin,Cout,H,W", CONV_CASES)
def test_weight_pack_and_conv_fprop(lib, N, Cin, Cout, H, W):
    torch.manual_seed(1)
    x = bf16r(torch.randn(N, Cin, H, W, device="cuda"))
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)
    wf, wd = _pack_weights(lib, w)
    wq = bf16r(w)
    # packing parity (bit exact)
    assert torch.equal(wf[:, :, :Cin].float(), wq.permute(2, 3, 0, 1).reshape(9, Cout, Cin))
    assert torch.equal(wd[:, :, :Cout].float(), wq.flip(2, 3).permute(2, 3, 1, 0).reshape(9, Cin, Cout))
    xb = make_buffer(N, H, W, 1, p8(Cin))
    put_nchw(xb, x, 1)
    yb = make_buffer(N, H, W, 0, p8(Cout))
    tiles = lib.mimo_conv3x3_m_tiles(N, H, W)
    ssum = torch.full((tiles, p8(Cout)), float("nan"), device="cuda")
    ssq = torch.full((tiles, p8(Cout)), float("nan"), device="cuda")
    _lib.check(lib.mimo_conv3x3(act_of(xb, 1, 0, Cin), 0, wf.data_ptr(), Cout, p8(Cin), yb.data_ptr(), p8(Cout), ssum.data_ptr(),
                                ssq.data_ptr(), None, 0, stream()), "conv3x3")
    torch.cuda.synchronize()
    ref = O.conv3x3_reflect(x, wq, None)
    got = get_nchw(yb, 0, 0, Cout)
    err = rel_l2(got, bf16r(ref))
    assert err <= TOL, f"fprop rel-L2 {err}"
    # pad channels are written as exact zeros
    if p8(Cout) > Cout:
        assert torch.all(yb[..., Cout:] == 0)
    # BatchNorm partial statistics of the stored values
    s = ssum.sum(0)[:Cout]
    q = ssq.sum(0)[:Cout]
    assert rel_l2(s, got.sum(dim=(0, 2, 3))) <= 1e-4 or float((s - got.sum(dim=(0, 2, 3))).abs().max()) < 1e-2
    assert rel_l2(q, (got * got).sum(dim=(0, 2, 3))) <= 1e-4


def _dy_buffer(dy, dy_pad):
    """dY in the dense (pad 0) or zero-tail (pad 2: buffer [N][H+2][W+2], data at the top-left) layout."""
    N, Cout, H, W = dy.shape
    e = 2 if dy_pad else 0
    t = torch.zeros(N, H + e, W + e, p8(Cout), dtype=torch.bfloat16, device="cuda")
    t[:, :H, :W, :Cout] = dy.permute(0, 2, 3, 1).to(torch.bfloat16)
    return t, _lib.Act(t.data_ptr(), N, H, W, dy_pad, p8(Cout), 0, Cout)


@pytest.mark.parametrize("dy_pad", [0, 2])
@pytest.mark.parametrize("N,Cin,Cout,H,W", CONV_CASES)
def test_conv_dgrad(lib, N, Cin, Cout, H, W, dy_pad):
    torch.manual_seed(2)
    dy = bf16r(torch.randn(N, Cout, H, W, device="cuda"))
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)
    wf, wd = _pack_weights(lib, w)
    dyb, dya = _dy_buffer(dy, dy_pad)
    dpad = make_buffer(N, H + 2, W + 2, 0, p8(Cin))
    _lib.check(lib.mimo_conv3x3(dya, 1, wd.data_ptr(), Cin, p8(Cout), dpad.data_ptr(), p8(Cin), None, None, None, 0,
                                stream()), "dgrad")
    # gradient w.r.t. the PADDED input of a valid conv == full correlation with the flipped kernel
    ref_pad = F.conv_transpose2d(dy, bf16r(w))
    got_pad = get_nchw(dpad, 0, 0, Cin)
    assert rel_l2(got_pad, bf16r(ref_pad)) <= TOL
    # fold the reflect halo back (adjoint of F.pad reflect) and compare with autograd of the oracle conv
    g = make_buffer(N, H, W, 0, p8(Cin))
    a = act_of(dpad, 0, 0, Cin)
    _lib.check(lib.mimo_grad_gather(C.byref(a), None, None, act_of(g, 0, 0, Cin), 0, stream()), "grad_gather")
    xg = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    O.conv3x3_reflect(xg, bf16r(w), None).backward(dy)
    # the fold sums up to 4 bf16-rounded values, so compare against the fold of the rounded padded gradient
    ref_fold = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    F.pad(ref_fold, (1, 1, 1, 1), mode="reflect").backward(got_pad)
    assert rel_l2(get_nchw(g, 0, 0, Cin), bf16r(ref_fold.grad)) <= TOL
    assert rel_l2(get_nchw(g, 0, 0, Cin), xg.grad) <= 4e-3  # vs exact autograd: two bf16 roundings


@pytest.mark.parametrize("dy_pad", [0, 2])
@pytest.mark.parametrize("N,Cin,Cout,H,W", CONV_CASES)
def test_conv_wgrad(lib, N, Cin, Cout, H, W, dy_pad):
    torch.manual_seed(3)
    x = bf16r(torch.randn(N, Cin, H, W, device="cuda"))
    dy = bf16r(torch.randn(N, Cout, H, W, device="cuda"))
    xb = make_buffer(N, H, W, 1, p8(Cin))
    put_nchw(xb, x, 1)
    dyb, dya = _dy_buffer(dy, dy_pad)
    scratch = torch.empty(9 * Cout * p8(Cin), device="cuda")
    grad = torch.full((Cout, Cin, 3, 3), float("nan"), device="cuda")
    _lib.check(lib.mimo_conv3x3_wgrad(dya, act_of(xb, 1, 0, Cin), scratch.data_ptr(), p8(Cin), grad.data_ptr(), 0,
                                      stream()), "wgrad")
    w = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    O.conv3x3_reflect(x, w, None).backward(dy)
    assert rel_l2(grad, w.grad) <= TOL
    # accumulate mode adds onto the existing gradient
    _lib.check(lib.mimo_conv3x3_wgrad(dya, act_of(xb, 1, 0, Cin), scratch.data_ptr(), p8(Cin), grad.data_ptr(), 1,
                                      stream()), "wgrad")
    assert rel_l2(grad, 2 * w.grad) <= TOL


WGRAD_FLATK_CASES = [
    # N, Cin, Cout, H, W   (more than 64 input or output channels: the stream-K flat wgrad kernel, conv_wgrad_flatk.cu)
    (2, 100, 70, 7, 9),       # one ci pair, one co tile, ragged channels
    (1, 200, 150, 5, 6),      # pair + single ci chunk, two co tiles (second one 22 channels)
    (2, 130, 40, 9, 11),      # pair + single chunk with 2 channels
    (3, 40, 130, 6, 10),      # single ci chunk only (no two-chunk items), two co tiles
    (1, 336, 336, 8, 10),     # three pairs, three co tiles
    (2, 64, 65, 16, 20),      # the smallest shapes that leave the <= 64-channel kernel
]


@pytest.mark.parametrize("N,Cin,Cout,H,W", WGRAD_FLATK_CASES)
def test_conv_wgrad_flatk(lib, N, Cin, Cout, H, W):
    """Stream-K flat wgrad (zero-tail dY layout) against autograd of the oracle convolution, incl. accumulate mode."""
    torch.manual_seed(7)
    x = bf16r(torch.randn(N, Cin, H, W, device="cuda"))
    dy = bf16r(torch.randn(N, Cout, H, W, device="cuda"))
    xb = make_buffer(N, H, W, 1, p8(Cin))
    put_nchw(xb, x, 1)
    dyb, dya = _dy_buffer(dy, 2)
    scratch = torch.empty(9 * Cout * p8(Cin), device="cuda")
    grad = torch.full((Cout, Cin, 3, 3), float("nan"), device="cuda")
    _lib.check(lib.mimo_conv3x3_wgrad(dya, act_of(xb, 1, 0, Cin), scratch.data_ptr(), p8(Cin), grad.data_ptr(), 0, stream()), "wgrad")
    w = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    O.conv3x3_reflect(x, w, None).backward(dy)
    assert rel_l2(grad, w.grad) <= TOL
    _lib.check(lib.mimo_conv3x3_wgrad(dya, act_of(xb, 1, 0, Cin), scratch.data_ptr(), p8(Cin), grad.data_ptr(), 1, stream()), "wgrad")
    assert rel_l2(grad, 2 * w.grad) <= TOL


FLATK_CASES = [
    # N, Cin, Cout, H, W   (65..256 input channels, <= 192 output channels: the K-chunked flat kernel)
    (2, 84, 42, 16, 20),
    (2, 168, 84, 12, 18),
    (3, 84, 168, 9, 13),
    (1, 200, 100, 7, 9),
    (2, 70, 192, 6, 6),
]


@pytest.mark.parametrize("N,Cin,Cout,H,W", FLATK_CASES)
def test_conv_flatk_fprop_dgrad(lib, monkeypatch, N, Cin, Cout, H, W):
    """Forces the K-chunked flat kernel (it is normally reserved for large feature maps) and checks fprop + BatchNorm
    statistics and dgrad against the oracle, like test_weight_pack_and_conv_fprop / test_conv_dgrad."""
    monkeypatch.setenv("MIMO_FLATK_MIN_ITEMS", "1")
    torch.manual_seed(5)
    x = bf16r(torch.randn(N, Cin, H, W, device="cuda"))
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)
    wf, wd = _pack_weights(lib, w)
    wq = bf16r(w)
    xb = make_buffer(N, H, W, 1, p8(Cin))
    put_nchw(xb, x, 1)
    yb = make_buffer(N, H, W, 0, p8(Cout))
    rows = lib.mimo_conv3x3_m_tiles(N, H, W)
    ssum = torch.full((rows, p8(Cout)), float("nan"), device="cuda")
    ssq = torch.full((rows, p8(Cout)), float("nan"), device="cuda")
    _lib.check(lib.mimo_conv3x3(act_of(xb, 1, 0, Cin), 0, wf.data_ptr(), Cout, p8(Cin), yb.data_ptr(), p8(Cout), ssum.data_ptr(),
                                ssq.data_ptr(), None, 0, stream()), "conv3x3")
    got = get_nchw(yb, 0, 0, Cout)
    assert rel_l2(got, bf16r(O.conv3x3_reflect(x, wq, None))) <= TOL
    if p8(Cout) > Cout:
        assert torch.all(yb[..., Cout:] == 0)
    s, q = ssum.sum(0)[:Cout], ssq.sum(0)[:Cout]
    assert rel_l2(s, got.sum(dim=(0, 2, 3))) <= 1e-4 or float((s - got.sum(dim=(0, 2, 3))).abs().max()) < 1e-2
    assert rel_l2(q, (got * got).sum(dim=(0, 2, 3))) <= 1e-4
    # dgrad over the zero-tail dY layout
    dy = bf16r(torch.randn(N, Cout, H, W, device="cuda"))
    dyb, dya = _dy_buffer(dy, 2)
    dpad = make_buffer(N, H + 2, W + 2, 0, p8(Cin))
    _lib.check(lib.mimo_conv3x3(dya, 1, wd.data_ptr(), Cin, p8(Cout), dpad.data_ptr(), p8(Cin), None, None, None, 0, stream()), "dgrad")
    assert rel_l2(get_nchw(dpad, 0, 0, Cin), bf16r(F.conv_transpose2d(dy, wq))) <= TOL


def test_conv_eval_epilogue_bias_relu(lib):
    torch.manual_seed(4)
    N, Cin, Cout, H, W = 2, 16, 24, 12, 20
    x = bf16r(torch.randn(N, Cin, H, W, device="cuda"))
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device="cuda")
    wf, _ = _pack_weights(lib, w)
    xb = make_buffer(N, H, W, 1, p8(Cin))
    put_nchw(xb, x, 1)
    yb = make_buffer(N, H, W, 0, p8(Cout))
    _lib.check(lib.mimo_conv3x3(act_of(xb, 1, 0, Cin), 0, wf.data_ptr(), Cout, p8(Cin), yb.data_ptr(), p8(Cout), None, None,
                                b.data_ptr(), 1, stream()))
    ref = F.relu(O.conv3x3_reflect(x, bf16r(w), b))
    assert rel_l2(get_nchw(yb, 0, 0, Cout), bf16r(ref)) <= TOL


def test_pack_input_gather_and_halo(lib):
    torch.manual_seed(5)
    B, S, Cc, H, W = 5, 2, 3, 9, 11
    x = torch.rand(B, S, Cc, H, W, device="cuda")
    gather = torch.tensor([3, 1, 4, 0, 2], device="cuda")
    for s in range(S):
        buf = make_buffer(B, H, W, 1, 8)
        _lib.check(lib.mimo_pack_input(x[:, s].data_ptr(), S * Cc * H * W, H * W, gather.data_ptr(), act_of(buf, 1, 0, Cc), stream()))
        ref = F.pad(bf16r(x[gather, s]), (1, 1, 1, 1), mode="reflect")
        assert torch.equal(get_nchw(buf, 1, 0, Cc, with_halo=True), ref)


@pytest.mark.parametrize("H,W,CH,c_off,cp", [(8, 10, 21, 0, 24), (9, 11, 42, 42, 88), (3, 3, 5, 8, 16), (2, 2, 8, 0, 8), (37, 45, 21, 0, 64)])
def test_bn_finalize_apply_pool(lib, H, W, CH, c_off, cp):
    torch.manual_seed(6)
    N = 3
    y = bf16r(torch.randn(N, CH, H, W, device="cuda") * 2 + 0.5)
    yb = make_buffer(N, H, W, 0, p8(CH), fill=0.0)
    put_nchw(yb, y, 0)
    gamma, beta, bias = torch.rand(CH, device="cuda") + 0.5, torch.randn(CH, device="cuda") * 0.2, torch.randn(CH, device="cuda") * 0.1
    rm, rv = torch.randn(CH, device="cuda") * 0.2, torch.rand(CH, device="cuda") + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    nbt = torch.tensor(3, device="cuda")
    # partial statistics as the conv epilogue would produce them (2 fake tiles)
    ssum = torch.zeros(2, p8(CH), device="cuda")
    ssq = torch.zeros(2, p8(CH), device="cuda")
    ssum[0, :CH] = y[:1].sum(dim=(0, 2, 3)); ssum[1, :CH] = y[1:].sum(dim=(0, 2, 3))
    ssq[0, :CH] = (y[:1] ** 2).sum(dim=(0, 2, 3)); ssq[1, :CH] = (y[1:] ** 2).sum(dim=(0, 2, 3))
    vec = torch.empty(4, p8(CH), device="cuda")
    _lib.check(lib.mimo_bn_finalize(ssum.data_ptr(), ssq.data_ptr(), 2, p8(CH), CH, float(N * H * W), gamma.data_ptr(), beta.data_ptr(),
                                    bias.data_ptr(), rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), 0.1, 1e-5, vec[0].data_ptr(),
                                    vec[1].data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), stream()))
    out_ref, mean, var = O.batchnorm_train(y, gamma, beta)
    erm, erv = O.updated_running_stats(rm0, rv0, mean + bias, var, N * H * W)
    assert torch.allclose(vec[2, :CH], mean, atol=1e-5) and torch.allclose(vec[3, :CH], torch.rsqrt(var + 1e-5), rtol=1e-4)
    assert torch.allclose(rm, erm, atol=1e-5) and torch.allclose(rv, erv, rtol=1e-5) and int(nbt) == 4
    drop = (torch.rand(N, CH, device="cuda") > 0.3).float() / 0.7
    ob = make_buffer(N, H, W, 1, cp)
    pb = make_buffer(N, H // 2, W // 2, 1, cp) if H // 2 >= 2 and W // 2 >= 2 else None
    pool_act = act_of(pb, 1, c_off, CH) if pb is not None else None
    _lib.check(lib.mimo_bn_relu_apply(yb.data_ptr(), p8(CH), vec[0].data_ptr(), vec[1].data_ptr(), drop.data_ptr(), act_of(ob, 1, c_off, CH),
                                      C.byref(pool_act) if pool_act is not None else None, stream()))
    act_ref = bf16r(F.relu(out_ref) * drop[:, :, None, None])
    got = get_nchw(ob, 1, c_off, CH, with_halo=True)
    assert rel_l2(got[:, :, 1:-1, 1:-1], act_ref) <= TOL
    # halo == reflect padding of the stored interior (bit exact)
    assert torch.equal(got, F.pad(got[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode="reflect"))
    if pb is not None:
        gp = get_nchw(pb, 1, c_off, CH, with_halo=True)
        assert torch.equal(gp[:, :, 1:-1, 1:-1], F.max_pool2d(got[:, :, 1:-1, 1:-1], 2))
        assert torch.equal(gp, F.pad(gp[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode="reflect"))
    # channels outside the slice are untouched (still NaN)
    if c_off > 0:
        assert torch.isnan(ob[..., :c_off].float()).all()


def test_bn_eval_affine(lib):
    C_ = 21
    gamma, beta, bias = torch.rand(C_, device="cuda") + 0.5, torch.randn(C_, device="cuda"), torch.randn(C_, device="cuda")
    rm, rv = torch.randn(C_, device="cuda"), torch.rand(C_, device="cuda") + 0.5
    vec = torch.empty(4, 24, device="cuda")
    _lib.check(lib.mimo_bn_eval_affine(C_, gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(), rm.data_ptr(), rv.data_ptr(), 1e-5,
                                       vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), stream()))
    y = torch.randn(2, C_, 4, 4, device="cuda")
    ref = O.batchnorm_eval(y + bias[None, :, None, None], gamma, beta, rm, rv)
    got = y * vec[0, :C_][None, :, None, None] + vec[1, :C_][None, :, None, None]
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)


def test_maxpool_indices_golden(lib, golden_dir):
    g = torch.load(f"{golden_dir}/op_cases.pt")["components"]["maxpool"]
    x, pooled, idx = g["x"].cuda(), g["pooled"].cuda(), g["idx"].cuda()
    xq = bf16r(x)
    N, Cc, H, W = x.shape
    xb = make_buffer(N, H, W, 1, p8(Cc))
    put_nchw(xb, xq, 1)
    ob = make_buffer(N, H // 2, W // 2, 0, p8(Cc))
    ib = torch.full((N, Cc, H // 2, W // 2), -1, dtype=torch.int64, device="cuda")
    _lib.check(lib.mimo_maxpool2x2(act_of(xb, 1, 0, Cc), act_of(ob, 0, 0, Cc), ib.data_ptr(), stream()))
    # values are bf16-rounded before pooling -> compare with the oracle fed the same rounded input
    pr, ir = F.max_pool2d(xq, 2, return_indices=True)
    assert torch.equal(get_nchw(ob, 0, 0, Cc), pr)
    assert torch.equal(ib, ir)
    # rounding is monotone, so the indices also equal the reference's fp32 indices wherever there is no new tie
    same = (ib == idx)
    assert same.float().mean() >= 0.97
    assert ib[0, 0, 0, 0] == 0  # tie -> first element wins (reference fixture plants the tie)


def test_upsample_fwd_bwd(lib, golden_dir):
    g = torch.load(f"{golden_dir}/op_cases.pt")["components"]["bilinear"]
    x = bf16r(g["x"].cuda())
    N, Cc, H, W = x.shape
    xb = make_buffer(N, H, W, 1, p8(Cc))
    put_nchw(xb, x, 1)
    for (OH, OW) in [(2 * H, 2 * W), (2 * H + 1, 2 * W + 1)]:
        ob = make_buffer(N, OH, OW, 1, 24)
        _lib.check(lib.mimo_upsample_bilinear2x(act_of(xb, 1, 0, Cc), act_of(ob, 1, 10, Cc), stream()))
        ref = O.pad_to(O.upsample_bilinear2x_ac(x), OH, OW)
        got = get_nchw(ob, 1, 10, Cc, with_halo=True)
        assert rel_l2(got[:, :, 1:-1, 1:-1], bf16r(ref)) <= TOL
        assert torch.equal(got, F.pad(got[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode="reflect"))
        # backward (gather form) against autograd
        gdst = bf16r(torch.randn(N, Cc, OH, OW, device="cuda"))
        gb = make_buffer(N, OH, OW, 0, p8(Cc), fill=0.0)
        put_nchw(gb, gdst, 0)
        gs = make_buffer(N, H, W, 0, p8(Cc))
        _lib.check(lib.mimo_upsample_bilinear2x_bwd(act_of(gb, 0, 0, Cc), act_of(gs, 0, 0, Cc), 0, stream()))
        xg = x.clone().requires_grad_(True)
        O.pad_to(O.upsample_bilinear2x_ac(xg), OH, OW).backward(gdst)
        assert rel_l2(get_nchw(gs, 0, 0, Cc), bf16r(xg.grad)) <= TOL
    # the reference's golden forward (fp32) within bf16 storage error
    ob = make_buffer(N, 2 * H, 2 * W, 1, p8(Cc))
    _lib.check(lib.mimo_upsample_bilinear2x(act_of(xb, 1, 0, Cc), act_of(ob, 1, 0, Cc), stream()))
    assert rel_l2(get_nchw(ob, 1, 0, Cc), g["y"].cuda()) <= 6e-3


@pytest.mark.parametrize("N,C_,H,W,OH,OW", [(2, 42, 16, 20, 32, 40), (1, 84, 8, 10, 17, 21), (2, 21, 5, 7, 10, 15), (1, 336, 4, 5, 8, 10),
                                            (2, 13, 33, 4, 66, 9), (1, 8, 2, 3, 4, 6)])
def test_upsample_bwd_shapes(lib, N, C_, H, W, OH, OW):
    """Bilinear x2 (align_corners) backward in gather form against autograd, incl. the zero-padded odd skip sizes; maps of
    4+ pixels take the windowed kernel, smaller ones the generic one."""
    torch.manual_seed(12)
    x = bf16r(torch.randn(N, C_, H, W, device="cuda"))
    gdst = bf16r(torch.randn(N, C_, OH, OW, device="cuda"))
    gb = make_buffer(N, OH, OW, 0, p8(C_), fill=0.0)
    put_nchw(gb, gdst, 0)
    gs = make_buffer(N, H, W, 0, p8(C_))
    _lib.check(lib.mimo_upsample_bilinear2x_bwd(act_of(gb, 0, 0, C_), act_of(gs, 0, 0, C_), 0, stream()))
    xg = x.clone().requires_grad_(True)
    O.pad_to(O.upsample_bilinear2x_ac(xg), OH, OW).backward(gdst)
    assert rel_l2(get_nchw(gs, 0, 0, C_), bf16r(xg.grad)) <= TOL
    # accumulate mode adds onto what is there (the stored bf16 value)
    before = get_nchw(gs, 0, 0, C_).clone()
    _lib.check(lib.mimo_upsample_bilinear2x_bwd(act_of(gb, 0, 0, C_), act_of(gs, 0, 0, C_), 1, stream()))
    assert rel_l2(get_nchw(gs, 0, 0, C_), bf16r(before + xg.grad)) <= 3e-3


@pytest.mark.parametrize("C_,c_off,cpitch", [(42, 21, 64), (42, 0, 48), (84, 84, 168), (5, 10, 24), (21, 3, 24), (168, 168, 336)])
def test_upsample_into_concat_slice(lib, C_, c_off, cpitch):
    """Decoder / core concat geometry: the up-sampled map lands in a channel slice whose offset need not be a multiple
    of 8 (21 + 42 channels in the decoder); neighbouring slices and the halo must stay intact."""
    torch.manual_seed(11)
    N, H, W = 2, 5, 7
    x = bf16r(torch.randn(N, C_, H, W, device="cuda"))
    xb = make_buffer(N, H, W, 1, p8(C_))
    put_nchw(xb, x, 1)
    OH, OW = 2 * H, 2 * W + 1
    ob = make_buffer(N, OH, OW, 1, cpitch, fill=0.0)
    sentinel = 7.0
    ob[...] = sentinel
    _lib.check(lib.mimo_upsample_bilinear2x(act_of(xb, 1, 0, C_), act_of(ob, 1, c_off, C_), stream()))
    ref = O.pad_to(O.upsample_bilinear2x_ac(x), OH, OW)
    got = get_nchw(ob, 1, c_off, C_, with_halo=True)
    assert rel_l2(got[:, :, 1:-1, 1:-1], bf16r(ref)) <= TOL
    assert torch.equal(got, F.pad(got[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode="reflect"))
    # channels before the slice are untouched; channels after it are untouched unless they are pad channels of the
    # buffer's last 16-byte group (those may be zeroed)
    assert torch.all(ob[..., :c_off] == sentinel)
    rest = ob[..., c_off + C_:]
    assert torch.all((rest == sentinel) | (rest == 0))
    if c_off + C_ <= cpitch - 8:
        assert torch.all(rest == sentinel)


@pytest.mark.parametrize("f,cup,cpitch", [(21, 42, 64), (30, 60, 96), (8, 16, 24), (16, 16, 32), (5, 3, 8), (21, 84, 112)])
@pytest.mark.parametrize("H,W,extra", [(5, 7, 1), (3, 3, 0), (8, 6, 0)])
def test_upsample_concat_whole_lines(lib, f, cup, cpitch, H, W, extra):
    """Decoder concat in one pass (components.py:110-119): skip channels copied from a dense buffer, up-sampled channels funnel-shifted
    behind them (f % 8 = 5, 6, 0 ... covers the shifted and the aligned word layouts), pad channels zero, reflect halo written."""
    torch.manual_seed(13)
    N = 2
    x = bf16r(torch.randn(N, cup, H, W, device="cuda"))
    xb = make_buffer(N, H, W, 1, p8(cup))
    put_nchw(xb, x, 1)
    OH, OW = 2 * H, 2 * W + extra
    skip = bf16r(torch.randn(N, f, OH, OW, device="cuda"))
    sb = make_buffer(N, OH, OW, 0, p8(f))          # pad channels keep the NaN fill: they must not leak into the output
    put_nchw(sb, skip, 0)
    ob = make_buffer(N, OH, OW, 1, cpitch)
    _lib.check(lib.mimo_upsample_concat(act_of(xb, 1, 0, cup), act_of(sb, 0, 0, f), act_of(ob, 1, 0, f + cup), stream()))
    ref = torch.cat([skip, bf16r(O.pad_to(O.upsample_bilinear2x_ac(x), OH, OW))], 1)
    got = get_nchw(ob, 1, 0, f + cup, with_halo=True)
    assert torch.equal(got[:, :f, 1:-1, 1:-1], skip)                       # copied bit for bit
    assert rel_l2(got[:, f:, 1:-1, 1:-1], ref[:, f:]) <= TOL
    assert torch.equal(got, F.pad(got[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode="reflect"))
    assert torch.all(ob[..., f + cup:] == 0)                                # pad channels of the pitch
    # the same values as the two-step path (up-sampling into the slice)
    ob2 = make_buffer(N, OH, OW, 1, cpitch, fill=0.0)
    _lib.check(lib.mimo_upsample_bilinear2x(act_of(xb, 1, 0, cup), act_of(ob2, 1, f, cup), stream()))
    assert torch.equal(get_nchw(ob2, 1, f, cup), got[:, f:, 1:-1, 1:-1])


@pytest.mark.parametrize("H,W", [(8, 10), (9, 11), (3, 3)])
def test_grad_gather_fold_and_pool_bwd(lib, H, W):
    torch.manual_seed(7)
    N, Cc = 2, 13
    act = bf16r(torch.randn(N, Cc, H, W, device="cuda"))
    act[0, 0, 0, 0] = act[0, 0, 0, 1] = 3.0  # tie inside the first window
    ab = make_buffer(N, H, W, 1, 16)
    put_nchw(ab, act, 1)
    dpad = bf16r(torch.randn(N, Cc, H + 2, W + 2, device="cuda"))
    db = make_buffer(N, H + 2, W + 2, 0, 16, fill=0.0)
    put_nchw(db, dpad, 0)
    gp = bf16r(torch.randn(N, Cc, H // 2, W // 2, device="cuda"))
    gpb = make_buffer(N, max(H // 2, 1), max(W // 2, 1), 0, 16, fill=0.0)
    put_nchw(gpb, gp, 0)
    gout = make_buffer(N, H, W, 0, 16)
    a, b, c = act_of(db, 0, 0, Cc), act_of(gpb, 0, 0, Cc), act_of(ab, 1, 0, Cc)
    _lib.check(lib.mimo_grad_gather(C.byref(a), C.byref(b), C.byref(c), act_of(gout, 0, 0, Cc), 0, stream()))
    xa = act.clone().requires_grad_(True)
    (F.pad(xa, (1, 1, 1, 1), mode="reflect") * dpad).sum().backward()
    fold = xa.grad.clone()
    xa.grad = None
    (F.max_pool2d(xa, 2) * gp).sum().backward()
    ref = fold + xa.grad
    assert rel_l2(get_nchw(gout, 0, 0, Cc), bf16r(ref)) <= TOL


@pytest.mark.parametrize("training", [1, 0])
@pytest.mark.parametrize("dy_pad", [0, 2])
@pytest.mark.parametrize("C_,c_off,gcp", [(21, 0, 24), (42, 42, 88), (336, 0, 336)])
def test_bn_relu_bwd(lib, training, C_, c_off, gcp, dy_pad):
    _bn_relu_bwd_case(lib, training, C_, c_off, gcp, dy_pad, 3, 6, 7)


@pytest.mark.parametrize("training", [1, 0])
@pytest.mark.parametrize("N,H,W,C_", [(4, 96, 160, 21), (2, 8, 400, 21), (2, 33, 70, 168), (5, 10, 12, 336), (1, 300, 37, 31)])
def test_bn_relu_bwd_bulk_shapes(lib, training, N, H, W, C_):
    """Dense operands take the cp.async.bulk pipelined kernels: many chunks per block, several rows per chunk, rows split into
    segments (W = 400), ragged last chunks."""
    _bn_relu_bwd_case(lib, training, C_, 0, p8(C_), 2, N, H, W)


@pytest.mark.parametrize("training", [1, 0])
@pytest.mark.parametrize("N,H,W,C_", [(2, 16, 20, 21), (3, 4, 3, 31), (2, 9, 13, 168), (1, 64, 160, 21), (2, 5, 20, 336), (2, 5, 40, 336)])
def test_bn_relu_bwd_folded(lib, training, N, H, W, C_):
    """BN/ReLU backward whose upstream gradient is fold_reflect(dpad) (adjoint of the reflect halo), fused into the bulk kernels.
    The last case exceeds the fused kernel's stage size: the fold is materialised in bf16 first (one more rounding)."""
    fused = 2 * (W + 2) * p8(C_) * 2 <= 50 * 1024 and W * p8(C_) * 2 <= 16 * 1024
    _bn_relu_bwd_case(lib, training, C_, 0, p8(C_), 2, N, H, W, folded=True, tol=TOL if fused else 4e-3)


def _bn_relu_bwd_case(lib, training, C_, c_off, gcp, dy_pad, N, H, W, folded=False, tol=TOL):
    torch.manual_seed(8)
    y = bf16r(torch.randn(N, C_, H, W, device="cuda") * 1.5 + 0.3)
    G = bf16r(torch.randn(N, C_, H, W, device="cuda"))
    if folded:
        dpad = bf16r(torch.randn(N, C_, H + 2, W + 2, device="cuda"))
        t = torch.zeros(N, C_, H, W, device="cuda", requires_grad=True)
        F.pad(t, (1, 1, 1, 1), mode="reflect").backward(dpad)
        G = t.grad.clone()  # exact fp32 fold: the fused kernels never round it to bf16
    gamma, beta = torch.rand(C_, device="cuda") + 0.5, torch.randn(C_, device="cuda") * 0.3
    drop = (torch.rand(N, C_, device="cuda") > 0.2).float() / 0.8
    yq = y.clone().requires_grad_(True)
    gq, bq = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    if training:
        z, mean, var = O.batchnorm_train(yq, gq, bq)
        invstd = torch.rsqrt(var + 1e-5).detach()
        mean = mean.detach()
    else:
        mean, invstd = torch.randn(C_, device="cuda") * 0.2, torch.rsqrt(torch.rand(C_, device="cuda") + 0.5)
        z = (yq - mean[None, :, None, None]) * (invstd * gq)[None, :, None, None] + bq[None, :, None, None]
    (F.relu(z) * drop[:, :, None, None] * G).sum().backward()
    scale = (gamma * invstd).contiguous()
    shift = (beta - mean * gamma * invstd).contiguous()
    yb = make_buffer(N, H, W, 0, p8(C_), fill=0.0)
    put_nchw(yb, y, 0)
    gb = make_buffer(N, H, W, 0, gcp, fill=0.0)
    put_nchw(gb, G, 0, c_off)
    e = 2 if dy_pad else 0
    dyb = torch.zeros(N, H + e, W + e, p8(C_), dtype=torch.bfloat16, device="cuda")
    dyb[:, :H, :W] = float("nan")
    dya = _lib.Act(dyb.data_ptr(), N, H, W, dy_pad, p8(C_), 0, C_)
    part = torch.empty(int(lib.mimo_bn_bwd_scratch_floats(C_)), device="cuda")
    s1s2 = torch.empty(2 * C_, device="cuda")
    dgamma, dbeta, dbias = (torch.full((C_,), float("nan"), device="cuda") for _ in range(3))
    if folded:
        dpb = make_buffer(N, H + 2, W + 2, 0, p8(C_), fill=0.0)
        put_nchw(dpb, dpad, 0)
        gb.fill_(float("nan"))  # scratch only
        _lib.check(lib.mimo_bn_relu_bwd_folded(act_of(dpb, 0, 0, C_), act_of(gb, 0, c_off, C_), yb.data_ptr(), p8(C_), scale.data_ptr(),
                                               shift.data_ptr(), mean.contiguous().data_ptr(), invstd.contiguous().data_ptr(), drop.data_ptr(),
                                               training, part.data_ptr(), s1s2.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                               dbias.data_ptr(), 0, dya, stream()))
    else:
        _lib.check(lib.mimo_bn_relu_bwd(act_of(gb, 0, c_off, C_), yb.data_ptr(), p8(C_), scale.data_ptr(), shift.data_ptr(), mean.contiguous().data_ptr(),
                                        invstd.contiguous().data_ptr(), drop.data_ptr(), training, part.data_ptr(), s1s2.data_ptr(), dgamma.data_ptr(),
                                        dbeta.data_ptr(), dbias.data_ptr(), 0, dya, stream()))
    assert torch.all(dyb[:, H:] == 0) and torch.all(dyb[:, :, W:] == 0)  # the zero tail is never written
    dyb = dyb[:, :H, :W].contiguous()
    assert rel_l2(get_nchw(dyb, 0, 0, C_), bf16r(yq.grad)) <= tol
    assert rel_l2(dgamma, gq.grad) <= (1e-4 if tol <= TOL else tol / 2) and rel_l2(dbeta, bq.grad) <= (1e-4 if tol <= TOL else tol / 2)
    if training:
        assert torch.all(dbias == 0)
    else:
        assert rel_l2(dbias, yq.grad.sum(dim=(0, 2, 3))) <= 1e-2
    if p8(C_) > C_:
        assert torch.all(dyb[..., C_:] == 0)


def test_head_fwd_bwd(lib):
    torch.manual_seed(9)
    B, S, f, K, H, W = 3, 2, 21, 2, 9, 11
    feat = bf16r(torch.rand(B, f, H, W, device="cuda"))
    w, b = torch.randn(K, f, device="cuda") * 0.3, torch.randn(K, device="cuda")
    fb = make_buffer(B, H, W, 1, p8(f))
    put_nchw(fb, feat, 1)
    out = torch.full((B, S, K, H, W), float("nan"), device="cuda")
    s = 1
    _lib.check(lib.mimo_head1x1(act_of(fb, 1, 0, f), w.data_ptr(), b.data_ptr(), K, out[:, s].data_ptr(), S * K * H * W, stream()))
    ref = F.conv2d(feat, w[:, :, None, None], b)
    assert torch.allclose(out[:, s], ref, atol=1e-5, rtol=1e-5) and torch.isnan(out[:, 0]).all()
    dout = torch.randn(B, S, K, H, W, device="cuda")
    gs = torch.tensor(4.0, device="cuda")
    gb = make_buffer(B, H, W, 0, p8(f))
    part = torch.empty(int(lib.mimo_head1x1_bwd_scratch_floats(K, f)), device="cuda")
    dw, db = torch.full((K, f), float("nan"), device="cuda"), torch.full((K,), float("nan"), device="cuda")
    _lib.check(lib.mimo_head1x1_bwd(act_of(fb, 1, 0, f), w.data_ptr(), K, dout[:, s].data_ptr(), S * K * H * W, gs.data_ptr(),
                                    act_of(gb, 0, 0, f), part.data_ptr(), dw.data_ptr(), db.data_ptr(), 0, stream()))
    fq = feat.clone().requires_grad_(True)
    wq, bq = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    (F.conv2d(fq, wq[:, :, None, None], bq) * dout[:, s] * 4.0).sum().backward()
    assert rel_l2(get_nchw(gb, 0, 0, f), bf16r(fq.grad)) <= TOL
    assert rel_l2(dw, wq.grad) <= 1e-5 and rel_l2(db, bq.grad) <= 1e-5
    assert torch.all(gb[..., f:] == 0)


@pytest.mark.parametrize("C_,c_off,cp,pad", [(21, 0, 24, 1), (336, 0, 336, 1), (13, 8, 32, 0), (42, 21, 64, 0)])
def test_mask_mul_elementwise_dropout(lib, C_, c_off, cp, pad):
    """nn.Dropout as a keep-mask multiply (reference model.py:239,294): interior scaled in fp32, everything else untouched."""
    torch.manual_seed(21)
    N, H, W = 2, 5, 7
    x = bf16r(torch.randn(N, C_, H, W, device="cuda"))
    buf = make_buffer(N, H, W, pad, cp, fill=3.0)
    put_nchw(buf, x, pad, c_off)
    before = buf.clone()
    keep = (torch.rand(N, H, W, p8(C_), device="cuda") >= 0.3).to(torch.bfloat16).contiguous()
    scale = 1.0 / 0.7
    _lib.check(lib.mimo_mask_mul(act_of(buf, pad, c_off, C_), keep.data_ptr(), p8(C_), scale, stream()))
    got = get_nchw(buf, pad, c_off, C_)
    ref = bf16r(x * keep[..., :C_].permute(0, 3, 1, 2).float() * scale)
    assert torch.equal(got, ref)
    # channels outside the view and the halo are untouched
    after = buf.clone()
    o = 1 if pad == 1 else 0
    after[:, o:o + H, o:o + W, c_off:c_off + C_] = before[:, o:o + H, o:o + W, c_off:c_off + C_]
    assert torch.equal(after, before)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_conv_flat2_opt_in_kernel_parity(mode):
    """The bulk-copy-loader small-channel kernel (conv_flat2.cu) is opt-in (MIMO_CONV_FLAT2 = 1 single-CTA MMAs, 2 CTA pairs; read
    once per process): run the fprop / dgrad parity cases in a subprocess with the switch on."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MIMO_CONV_FLAT2=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_kernels_gpu.py"), "-q", "-m", "gpu", "-x", "-p",
                        "no:cacheprovider", "-k", "test_weight_pack_and_conv_fprop or test_conv_dgrad"], capture_output=True, text=True,
                       env=env, cwd=root, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
