"""Stand-alone execution of the reference's building blocks (DoubleConv / Down / Up / OutConv) on the CUDA
kernels, NCHW fp32 in and out like the reference modules (components.py:8-129).  Forward only: training of
isolated blocks is not part of the reference's path (the whole network trains through the C++ executor).
PyTorch only allocates the buffers; every byte is produced by libmimo_b200.so kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Act, check
from .engine import stream_ptr


def _p8(c: int) -> int:
    return (c + 7) // 8 * 8


def _require_cuda(x):
    if not x.is_cuda:
        raise _lib.MimoError("mimo_unet_b200 blocks need CUDA tensors (no CPU fallback)")
    if torch.is_grad_enabled() and x.requires_grad:
        raise NotImplementedError("stand-alone blocks are forward-only; train through MimoUNet")


class _Buf:
    """NHWC bf16 buffer (+halo) owned by torch."""

    def __init__(self, n, h, w, pad, c, device):
        self.n, self.h, self.w, self.pad, self.c, self.cp = n, h, w, pad, c, _p8(c)
        self.t = torch.empty((n, h + 2 * pad, w + 2 * pad, self.cp), dtype=torch.bfloat16, device=device)

    def act(self, c_off=0, c=None) -> Act:
        return Act(self.t.data_ptr(), self.n, self.h, self.w, self.pad, self.cp, c_off, self.c if c is None else c)


def _pack(x: torch.Tensor, buf: _Buf, c_off=0):
    x = x.detach().float().contiguous()
    n, c, h, w = x.shape
    check(_lib.lib().mimo_pack_input(x.data_ptr(), c * h * w, h * w, None, buf.act(c_off, c), stream_ptr()), "mimo_pack_input")


def _unpack(buf: _Buf) -> torch.Tensor:
    out = torch.empty((buf.n, buf.c, buf.h, buf.w), dtype=torch.float32, device=buf.t.device)
    check(_lib.lib().mimo_unpack_nchw(buf.act(), out.data_ptr(), stream_ptr()), "mimo_unpack_nchw")
    return out


def _conv_bn_relu(conv, bn, src: _Buf, dst: _Buf, drop: Optional[torch.Tensor]):
    lib = _lib.lib()
    dev = src.t.device
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    wf = torch.empty(9 * cout * _p8(cin), dtype=torch.bfloat16, device=dev)
    check(lib.mimo_weight_pack(conv.weight.detach().contiguous().data_ptr(), cout, cin, wf.data_ptr(), _p8(cin), None, 0, stream_ptr()))
    y = torch.empty((src.n, src.h, src.w, _p8(cout)), dtype=torch.bfloat16, device=dev)
    vec = torch.empty((4, _p8(cout)), dtype=torch.float32, device=dev)
    if bn.training:
        tiles = lib.mimo_conv3x3_m_tiles(src.n, src.h, src.w)
        ps = torch.empty((2, tiles, _p8(cout)), dtype=torch.float32, device=dev)
        check(lib.mimo_conv3x3(src.act(), 0, wf.data_ptr(), cout, _p8(cin), y.data_ptr(), _p8(cout), ps[0].data_ptr(), ps[1].data_ptr(),
                               None, 0, stream_ptr()), "mimo_conv3x3")
        check(lib.mimo_bn_finalize(ps[0].data_ptr(), ps[1].data_ptr(), tiles, _p8(cout), cout, float(src.n * src.h * src.w),
                                   bn.weight.data_ptr(), bn.bias.data_ptr(), conv.bias.data_ptr(), bn.running_mean.data_ptr(),
                                   bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), bn.momentum, bn.eps, vec[0].data_ptr(),
                                   vec[1].data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), stream_ptr()), "mimo_bn_finalize")
    else:
        check(lib.mimo_conv3x3(src.act(), 0, wf.data_ptr(), cout, _p8(cin), y.data_ptr(), _p8(cout), None, None, None, 0, stream_ptr()),
              "mimo_conv3x3")
        check(lib.mimo_bn_eval_affine(cout, bn.weight.data_ptr(), bn.bias.data_ptr(), conv.bias.data_ptr(), bn.running_mean.data_ptr(),
                                      bn.running_var.data_ptr(), bn.eps, vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(),
                                      vec[3].data_ptr(), stream_ptr()), "mimo_bn_eval_affine")
    check(lib.mimo_bn_relu_apply(y.data_ptr(), _p8(cout), vec[0].data_ptr(), vec[1].data_ptr(), None if drop is None else drop.data_ptr(),
                                 dst.act(), None, stream_ptr()), "mimo_bn_relu_apply")


def _double_conv_on(dc, src: _Buf) -> _Buf:
    seq = dc.double_conv
    dev = src.t.device
    mid = _Buf(src.n, src.h, src.w, 1, seq[0].weight.shape[0], dev)
    out = _Buf(src.n, src.h, src.w, 1, seq[3].weight.shape[0], dev)
    drop = None
    if dc.dropout.training and dc.dropout.p > 0:
        p = dc.dropout.p
        drop = ((torch.rand(src.n, out.c, device=dev) >= p).float() / (1 - p)).contiguous()
    _conv_bn_relu(seq[0], seq[1], src, mid, None)
    _conv_bn_relu(seq[3], seq[4], mid, out, drop)
    return out


def double_conv_forward(dc, x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    n, c, h, w = x.shape
    src = _Buf(n, h, w, 1, c, x.device)
    _pack(x, src)
    return _unpack(_double_conv_on(dc, src))


def down_forward(down, x: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    _require_cuda(x)
    n, c, h, w = x.shape
    src = _Buf(n, h, w, 1, c, x.device)
    _pack(x, src)
    pooled = _Buf(n, h // 2, w // 2, 1, c, x.device)
    idx = torch.empty((n, c, h // 2, w // 2), dtype=torch.int64, device=x.device) if down.use_pooling_indices else None
    check(_lib.lib().mimo_maxpool2x2(src.act(), pooled.act(), None if idx is None else idx.data_ptr(), stream_ptr()), "mimo_maxpool2x2")
    return _unpack(_double_conv_on(down.conv, pooled)), idx


def up_forward(up, x1: torch.Tensor, x2: torch.Tensor, pooling_indices=None) -> torch.Tensor:
    _require_cuda(x1)
    _require_cuda(x2)
    lib = _lib.lib()
    n, c1, h1, w1 = x1.shape
    _, c2, h2, w2 = x2.shape
    src = _Buf(n, h1, w1, 1, c1, x1.device)
    _pack(x1, src)
    c_up = c1 if (up.bilinear or up.use_pooling_indices) else up.up.weight.shape[1]
    cat = _Buf(n, h2, w2, 1, c2 + c_up, x1.device)
    _pack(x2, cat, 0)  # skip channels first (components.py:119)
    dst = cat.act(c2, c_up)
    if up.bilinear:
        check(lib.mimo_upsample_bilinear2x(src.act(), dst, stream_ptr()), "mimo_upsample_bilinear2x")
    elif up.use_pooling_indices:
        idx = pooling_indices.to(torch.int64).contiguous()
        check(lib.mimo_maxunpool2x2(src.act(), idx.data_ptr(), dst, stream_ptr()), "mimo_maxunpool2x2")
    else:
        wt, b = up.up.weight.detach().contiguous(), up.up.bias
        check(lib.mimo_convtranspose2x2(src.act(), wt.data_ptr(), None if b is None else b.data_ptr(), dst, stream_ptr()),
              "mimo_convtranspose2x2")
    return _unpack(_double_conv_on(up.conv, cat))


def outconv_forward(oc, x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    n, c, h, w = x.shape
    src = _Buf(n, h, w, 0, c, x.device)
    _pack(x, src)
    k = oc.conv.weight.shape[0]
    out = torch.empty((n, k, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().mimo_head1x1(src.act(), oc.conv.weight.detach().contiguous().data_ptr(), oc.conv.bias.data_ptr(), k, out.data_ptr(),
                                  k * h * w, stream_ptr()), "mimo_head1x1")
    return out
