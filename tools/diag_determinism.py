"""Runs forward + backward twice on identical inputs and reports the first intermediate buffer / gradient that is not
bit-identical between the two runs (backward order). Usage: python tools/diag_determinism.py [S f H W B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.profile_layers import node_shapes  # noqa: E402
from mimo_unet_b200.engine import UNetPlan  # noqa: E402


def main():
    args = [int(v) for v in sys.argv[1:]]
    S, f, H, W, B = (args + [2, 8, 32, 48, 2][len(args):])[:5]
    cfg = dict(S=S, f=f, cin=3)
    dev = torch.device("cuda")
    plan = UNetPlan(3, 2, S, f, B, H, W, dev)
    torch.manual_seed(0)
    shapes = node_shapes(cfg)
    state, names = [], []
    for name, ci, cm, co, lvl in shapes:
        for j, (i, o) in enumerate(((ci, cm), (cm, co))):
            state += [torch.randn(o, i, 3, 3, device=dev) * (2.0 / (9 * i)) ** 0.5, torch.zeros(o, device=dev), torch.ones(o, device=dev),
                      torch.zeros(o, device=dev), torch.zeros(o, device=dev), torch.ones(o, device=dev), torch.zeros((), dtype=torch.int64, device=dev)]
            names += [f"{name}.c{j + 1}.{k}" for k in ("w", "b", "gamma", "beta", "rm", "rv", "nbt")]
    for s in range(S):
        state += [torch.randn(2, f, 1, 1, device=dev) * 0.2, torch.zeros(2, device=dev)]
        names += [f"head{s}.w", f"head{s}.b"]
    grads = [torch.zeros_like(t) if t.dtype == torch.float32 and not n.endswith((".rm", ".rv")) else None for t, n in zip(state, names)]
    plan.bind(state, grads)
    x = torch.rand(B, S, 3, H, W, device=dev)
    out = torch.empty(B, S, 2, H, W, device=dev)
    dout = torch.randn(B, S, 2, H, W, device=dev) * 1e-3

    def run():
        plan.forward(x, out, True)
        plan.backward(dout)
        torch.cuda.synchronize()
        snap = {}
        for name, *_ in shapes:
            for buf in ("g2", "c2.dy", "c2.dpad", "g1", "c1.dy", "c1.dpad"):
                try:
                    snap[f"{name}.{buf}"] = plan.debug_tensor(f"{name}.{buf}")
                except Exception as e:  # noqa: BLE001
                    snap[f"{name}.{buf}"] = None
        for n, g in zip(names, grads):
            if g is not None:
                snap["grad:" + n] = g.clone()
        return snap

    a = run()
    for rep in range(3):
        b = run()
        bad = 0
        order = list(reversed([n for n, *_ in shapes]))
        for node in order:
            for buf in ("g2", "c2.dy", "c2.dpad", "g1", "c1.dy", "c1.dpad"):
                k = f"{node}.{buf}"
                if a[k] is None:
                    continue
                if not torch.equal(a[k], b[k]):
                    d = (a[k] - b[k]).abs()
                    print(f"rep {rep}: MISMATCH {k}: max abs {float(d.max()):.3e} (ref max {float(a[k].abs().max()):.3e}), {int((d > 0).sum())} of {d.numel()} elements")
                    bad += 1
        for k in a:
            if k.startswith("grad:") and not torch.equal(a[k], b[k]):
                d = (a[k] - b[k]).abs()
                rel = float(d.norm() / (a[k].norm() + 1e-30))
                if rel > 1e-5:
                    print(f"rep {rep}: grad {k}: rel {rel:.3e}")
        print(f"rep {rep}: {bad} mismatching buffers")


if __name__ == "__main__":
    main()
