"""Achieved HBM bandwidth of the fused loss / aggregation kernels (north star: >= 70 % of HBM on these memory-bound kernels).

Usage (GPU box): python tools/bench_loss.py [--reps 50]
  laplace_nll_train : fused LaplaceNLL + per-subnetwork mean + loss-buffer weights + gradient seed; algorithmic bytes =
                      20 B / element of [B,S,C,H,W] (read mu, log_s, y; write d mu, d log_s; fp32)  
  ensemble_aggregate: compute_uncertainties; algorithmic bytes = (2*S*4 + 12) B per output pixel
Inputs are larger than the 126 MB L2 or the L2 is flushed between repetitions.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimo_unet_b200 import functional as Fn  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)).get("hbm_gbs", 6559.7), "measured"
    return 6650.0, "fallback"


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        for _ in range(3):  # 256 MB writes: evict the previous repetition from the L2 and keep the GPU ~250 us behind the host,
            flush.add_(1)   # so the timed kernels are already enqueued when the GPU reaches e0 (no host launch gaps inside)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda")
    peak, src = peak_gbs()
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    print(f"# HBM peak {peak:.0f} GB/s ({src})")
    print(f"{'kernel':<22} {'shape':<28} {'us':>9} {'GB/s':>8} {'frac':>6}")
    for name, (B, S, H, W) in {"C2 train loss": (64, 2, 128, 160), "C3 train loss": (32, 2, 256, 256), "B256 train loss": (256, 2, 128, 160)}.items():
        out = torch.randn(B, S, 2, H, W, device=dev, requires_grad=True)
        y = torch.rand(B, S, 1, H, W, device=dev)
        lb = Fn.DeviceLossBuffer(S, 0.3, 10, dev)

        def f():
            Fn.laplace_train_loss(out, y, loss_buffer=lb)
        t = timed(f, a.reps, flush)
        bytes_ = B * S * H * W * 20.0
        print(f"{'laplace_nll_train':<22} {name + ' ' + str((B, S, H, W)):<28} {t:9.1f} {bytes_ / t * 1e-3:8.0f} {bytes_ / t * 1e-3 / peak:6.2f}")
    for name, (B, S, H, W) in {"C4 B64 S4": (64, 4, 128, 160), "C4 B256 S4": (256, 4, 128, 160), "C4 B64 S32 (mc 8)": (64, 32, 128, 160)}.items():
        p1 = torch.randn(B, S, 1, H, W, device=dev)
        p2 = torch.randn(B, S, 1, H, W, device=dev) * 0.3

        def g():
            Fn.ensemble_aggregate(p1, p2)
        t = timed(g, a.reps, flush)
        bytes_ = B * H * W * (2 * S * 4 + 12.0)
        print(f"{'ensemble_aggregate':<22} {name + ' ' + str((B, S, H, W)):<28} {t:9.1f} {bytes_ / t * 1e-3:8.0f} {bytes_ / t * 1e-3 / peak:6.2f}")


if __name__ == "__main__":
    main()
