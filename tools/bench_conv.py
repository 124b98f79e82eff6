"""Micro-benchmark + self-check of the convolution kernels on single layer shapes (development tool).

Usage (GPU box): [MIMO_CONV_FLAT=0|1] [MIMO_FLAT_BO=0|1] python tools/bench_conv.py [--set full|half|core|all] [--reps 20]
For every layer shape: fprop (+BatchNorm statistics), dgrad and wgrad through the C ABI, CUDA-event timed, with the
relative L2 error against torch's fp32 convolution of the same bf16-rounded operands.
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimo_unet_b200 import _lib  # noqa: E402
from mimo_unet_b200._lib import Act  # noqa: E402

SETS = {
    "full": [(64, 3, 21, 128, 160), (64, 21, 21, 128, 160), (64, 63, 31, 128, 160), (64, 31, 21, 128, 160)],
    "half": [(64, 21, 42, 64, 80), (64, 42, 42, 64, 80), (64, 168, 84, 64, 80), (64, 84, 42, 64, 80)],
    "core": [(64, 84, 168, 32, 40), (64, 168, 168, 32, 40), (64, 336, 336, 16, 20), (64, 672, 336, 16, 20), (64, 336, 336, 8, 10)],
    # every layer of C2 that the CTA-pair kernel serves (cin, cout as fprop sees them; dgrad swaps the roles)
    "c2all": [(64, 168, 84, 64, 80), (64, 84, 42, 64, 80), (64, 84, 168, 32, 40), (64, 168, 168, 32, 40), (64, 336, 168, 32, 40),
              (64, 168, 84, 32, 40), (64, 168, 336, 16, 20), (64, 336, 336, 16, 20), (64, 672, 336, 16, 20), (64, 336, 168, 16, 20),
              (64, 336, 336, 8, 10)],
    # C3 (fbc 30, 256 x 256 maps): the wide-map decoder layers
    "c3dec": [(32, 90, 45, 256, 256), (32, 45, 30, 256, 256), (32, 30, 30, 256, 256)],
    "probe": [(64, 21, 21, 128, 160), (64, 63, 31, 128, 160)],
    "small": [(2, 21, 21, 37, 45), (3, 63, 31, 16, 24), (2, 3, 21, 32, 32)],
}


def p8(c):
    return (c + 7) // 8 * 8


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def buf(n, h, w, pad, cp, fill=float("nan")):
    e = 2 if pad else 0
    return torch.full((n, h + e, w + e, cp), fill, dtype=torch.bfloat16, device="cuda")


def act(t, h, w, pad, c):
    return Act(t.data_ptr(), t.shape[0], h, w, pad, t.shape[3], 0, c)


def timeit(fn, reps):
    """GPU time per launch in microseconds: `reps` launches captured into ONE CUDA graph and replayed, so the host side of a launch
    (five tensor-map encodes, attribute call, ctypes) is outside the measurement -- back-to-back eager launches measure
    max(host, GPU), which hides everything below ~50 us."""
    fn()
    torch.cuda.synchronize()
    if os.environ.get("BENCH_CONV_EAGER", "0") == "1":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    global _STREAM
    with torch.cuda.stream(side):
        _STREAM = side.cuda_stream
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
        _STREAM = torch.cuda.current_stream().cuda_stream
    _STREAM = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


_STREAM = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="full")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fprop-only", action="store_true")
    ap.add_argument("--dy-pad", type=int, default=2, help="layout of the dY buffer: 0 dense, 2 zero tail (flat dgrad)")
    a = ap.parse_args()
    lib = _lib.lib()
    global _STREAM
    _STREAM = torch.cuda.current_stream().cuda_stream
    names = list(SETS) if a.set == "all" else a.set.split(",")
    print(f"# MIMO_CONV_FLAT={os.environ.get('MIMO_CONV_FLAT', '1')} MIMO_FLAT_BO={os.environ.get('MIMO_FLAT_BO', '1')} dy_pad={a.dy_pad}")
    print(f"{'shape':>28} | {'fprop us':>9} {'TF/s':>6} {'GB/s':>6} {'err':>8} {'stat':>8} | {'dgrad us':>9} {'TF/s':>6} {'err':>8} | {'wgrad us':>9} {'TF/s':>6} {'err':>8}")
    for name in names:
        for (N, Ci, Co, H, W) in SETS[name]:
            torch.manual_seed(0)
            x = torch.randn(N, Ci, H, W, device="cuda").bfloat16().float()
            w = torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)
            dy = torch.randn(N, Co, H, W, device="cuda").bfloat16().float()
            wq = w.bfloat16().float()
            wf = torch.zeros(9, Co, p8(Ci), dtype=torch.bfloat16, device="cuda")
            wd = torch.zeros(9, Ci, p8(Co), dtype=torch.bfloat16, device="cuda")
            _lib.check(lib.mimo_weight_pack(w.data_ptr(), Co, Ci, wf.data_ptr(), p8(Ci), wd.data_ptr(), p8(Co), _STREAM))
            xb = buf(N, H, W, 1, p8(Ci))
            xb[...] = 0
            xb[..., :Ci] = F.pad(x, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1).bfloat16()
            yb = buf(N, H, W, 0, p8(Co))
            rows = lib.mimo_conv3x3_m_tiles(N, H, W)
            ssum = torch.zeros(rows, p8(Co), device="cuda")
            ssq = torch.zeros(rows, p8(Co), device="cuda")
            flops = 2.0 * N * H * W * Ci * Co * 9

            def fprop():
                _lib.check(lib.mimo_conv3x3(act(xb, H, W, 1, Ci), 0, wf.data_ptr(), Co, p8(Ci), yb.data_ptr(), p8(Co), ssum.data_ptr(),
                                            ssq.data_ptr(), None, 0, _STREAM), "fprop")
            t_f = timeit(fprop, a.reps)
            ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), wq)
            got = yb[..., :Co].permute(0, 3, 1, 2).float()
            e_f = rel(got, ref.bfloat16().float())
            e_s = rel(ssum.sum(0)[:Co], got.sum(dim=(0, 2, 3)))
            e_s = max(e_s, rel(ssq.sum(0)[:Co], (got * got).sum(dim=(0, 2, 3))))
            bytes_f = N * H * W * (p8(Ci) + p8(Co)) * 2
            if a.fprop_only:
                print(f"{str((N, Ci, Co, H, W)):>28} | {t_f:9.1f} {flops / t_f * 1e-6:6.1f} {bytes_f / t_f * 1e-3:6.0f} {e_f:8.1e} {e_s:8.1e}", flush=True)
                continue

            dyb = buf(N, H, W, a.dy_pad, p8(Co), fill=0.0)
            dyb[:, :H, :W, :Co] = dy.permute(0, 2, 3, 1).bfloat16()
            dpad = buf(N, H + 2, W + 2, 0, p8(Ci))

            def dgrad():
                _lib.check(lib.mimo_conv3x3(act(dyb, H, W, a.dy_pad, Co), 1, wd.data_ptr(), Ci, p8(Co), dpad.data_ptr(), p8(Ci), None, None,
                                            None, 0, _STREAM), "dgrad")
            t_d = timeit(dgrad, a.reps)
            refd = F.conv_transpose2d(dy, wq)
            e_d = rel(dpad[..., :Ci].permute(0, 3, 1, 2).float(), refd.bfloat16().float())

            scratch = torch.empty(9 * Co * p8(Ci), device="cuda")
            grad = torch.zeros(Co, Ci, 3, 3, device="cuda")

            def wgrad():
                _lib.check(lib.mimo_conv3x3_wgrad(act(dyb, H, W, a.dy_pad, Co), act(xb, H, W, 1, Ci), scratch.data_ptr(), p8(Ci),
                                                  grad.data_ptr(), 0, _STREAM), "wgrad")
            t_w = timeit(wgrad, a.reps)
            wref = torch.zeros(Co, Ci, 3, 3, device="cuda", requires_grad=True)
            F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), wref).backward(dy)
            e_w = rel(grad, wref.grad)
            print(f"{str((N, Ci, Co, H, W)):>28} | {t_f:9.1f} {flops / t_f * 1e-6:6.1f} {bytes_f / t_f * 1e-3:6.0f} {e_f:8.1e} {e_s:8.1e} | "
                  f"{t_d:9.1f} {flops / t_d * 1e-6:6.1f} {e_d:8.1e} | {t_w:9.1f} {flops / t_w * 1e-6:6.1f} {e_w:8.1e}", flush=True)


if __name__ == "__main__":
    main()
