"""Per-launch CUDA-event profile of one training step of the C++ executor, grouped by layer.

Usage (GPU box): python tools/profile_layers.py [--cfg C2|C1|C3] [--out gpurun_out/layers.txt]
Prints, for every conv of the network, the live duration of its fprop / dgrad / wgrad / BN kernels, the algorithmic
FLOPs (true channel counts) and the resulting TFLOP/s, plus the memory-bound kernels with their GB/s.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimo_unet_b200 import _lib  # noqa: E402
from mimo_unet_b200.engine import UNetPlan  # noqa: E402

CFGS = {
    "C1": dict(cin=3, S=2, f=21, B=8, H=256, W=256),
    "C2": dict(cin=3, S=2, f=21, B=64, H=128, W=160),
    "C3": dict(cin=2, S=2, f=30, B=32, H=256, W=256),
    "C4": dict(cin=3, S=4, f=21, B=64, H=128, W=160),
}


def node_shapes(cfg):
    """(name, cin, cmid, cout, level) per double conv in executor order (mirrors csrc/engine.cu plan_create)."""
    S, f, cin = cfg["S"], cfg["f"], cfg["cin"]
    c = 2 * f * S
    d = c // 2 + f
    L = []
    for s in range(S):
        L.append((f"encoder.in_convs.{s}", cin, f, f, 0))
    for s in range(S):
        L.append((f"encoder.down1s.{s}", f, 2 * f, 2 * f, 1))
    L += [("core.down2", c, 2 * c, 2 * c, 2), ("core.down3", 2 * c, 4 * c, 4 * c, 3), ("core.down4", 4 * c, 4 * c, 4 * c, 4),
          ("core.up1", 8 * c, 4 * c, 2 * c, 3), ("core.up2", 4 * c, 2 * c, c, 2), ("core.up3", 2 * c, c, c // 2, 1)]
    for s in range(S):
        L.append((f"decoder.up4s.{s}", d, d // 2, f, 0))
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--eval", action="store_true")
    a = ap.parse_args()
    cfg = CFGS[a.cfg]
    S, f, cin, B, H, W = cfg["S"], cfg["f"], cfg["cin"], cfg["B"], cfg["H"], cfg["W"]
    dev = torch.device("cuda")
    lib = _lib.lib()
    plan = UNetPlan(cin, 2, S, f, B, H, W, dev)
    torch.manual_seed(0)
    shapes = node_shapes(cfg)
    state = []
    for name, ci, cm, co, lvl in shapes:
        for (i, o) in ((ci, cm), (cm, co)):
            state += [torch.randn(o, i, 3, 3, device=dev) * (2.0 / (9 * i)) ** 0.5, torch.zeros(o, device=dev), torch.ones(o, device=dev),
                      torch.zeros(o, device=dev), torch.zeros(o, device=dev), torch.ones(o, device=dev), torch.zeros((), dtype=torch.int64, device=dev)]
    for s in range(S):
        state += [torch.randn(2, f, 1, 1, device=dev) * 0.2, torch.zeros(2, device=dev)]
    grads = [torch.zeros_like(t) if t.dtype == torch.float32 and t.dim() != 0 else None for t in state]
    # running stats get no gradient
    k = 0
    for name, ci, cm, co, lvl in shapes:
        for _ in range(2):
            grads[k + 4] = None
            grads[k + 5] = None
            k += 7
    plan.bind(state, grads)
    x = torch.rand(B, S, cin, H, W, device=dev)
    out = torch.empty(B, S, 2, H, W, device=dev)
    dout = torch.randn(B, S, 2, H, W, device=dev) * 1e-3
    for _ in range(2):
        plan.forward(x, out, not a.eval)
        plan.backward(dout)
    torch.cuda.synchronize()
    lib.mimo_unet_profile_enable(plan.handle, 1)
    N = 4096
    ms, cls, tag = (C.c_float * N)(), (C.c_int * N)(), (C.c_int * N)()
    acc = {}
    total = 0.0
    for r in range(a.reps):
        plan.forward(x, out, not a.eval)
        plan.backward(dout)
        n = lib.mimo_unet_profile_read_launches(plan.handle, N, ms, cls, tag)
        for i in range(n):
            key = (tag[i], cls[i])
            acc[key] = acc.get(key, 0.0) + ms[i] / a.reps
            total += ms[i] / a.reps
    names = [lib.mimo_unet_profile_class_name(i).decode() for i in range(lib.mimo_unet_profile_classes())]
    Hs = [H >> l for l in range(5)]
    Ws = [W >> l for l in range(5)]
    lines = []
    hdr = f"{'layer':28s} {'cin':>5s} {'cout':>5s} {'HxW':>9s} {'GF':>8s} | " + " ".join(f"{n[:10]:>10s}" for n in
          ("conv_fprop", "conv_dgrad", "conv_wgrad", "bn_final", "bn_apply", "bn_bwd", "gather")) + " | fprop/dgrad/wgrad TF/s"
    lines.append(f"# {a.cfg} {cfg}  total kernel time per step {total:.3f} ms ({'eval' if a.eval else 'train'})")
    lines.append(hdr)
    cidx = {n: i for i, n in enumerate(names)}
    sums = {}
    for ni, (name, ci, cm, co, lvl) in enumerate(shapes):
        for j, (i_, o_) in enumerate(((ci, cm), (cm, co))):
            t = 2 * ni + j
            fl = 2.0 * B * Hs[lvl] * Ws[lvl] * i_ * o_ * 9
            g = lambda n: acc.get((t, cidx[n]), 0.0)
            tf = lambda n: (fl / (g(n) * 1e-3) / 1e12) if g(n) > 0 else 0.0
            cols = [g("conv_fprop"), g("conv_dgrad"), g("conv_wgrad") + g("wgrad_unpack"), g("bn_finalize"), g("bn_relu_apply"), g("bn_relu_bwd"), g("grad_gather")]
            lines.append(f"{name + '.c' + str(j + 1):28s} {i_:5d} {o_:5d} {Hs[lvl]:4d}x{Ws[lvl]:<4d} {fl / 1e9:8.1f} | " +
                         " ".join(f"{v * 1e3:10.1f}" for v in cols) + f" | {tf('conv_fprop'):6.0f} {tf('conv_dgrad'):6.0f} {tf('conv_wgrad'):6.0f}")
    by_cls = {}
    for (t, c), v in acc.items():
        by_cls[names[c]] = by_cls.get(names[c], 0.0) + v
    lines.append("# per class (ms/step): " + json.dumps({k: round(v, 3) for k, v in sorted(by_cls.items(), key=lambda kv: -kv[1])}))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
