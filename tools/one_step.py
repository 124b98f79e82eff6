"""One training step of the C++ executor (forward + backward at a named config) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`. Usage: python tools/one_step.py [C2|C1|C3|C4] [--eval]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.profile_layers import CFGS, node_shapes  # noqa: E402
from mimo_unet_b200.engine import UNetPlan  # noqa: E402


def build(cfg_name):
    cfg = CFGS[cfg_name]
    S, f, cin, B, H, W = cfg["S"], cfg["f"], cfg["cin"], cfg["B"], cfg["H"], cfg["W"]
    dev = torch.device("cuda")
    plan = UNetPlan(cin, 2, S, f, B, H, W, dev)
    torch.manual_seed(0)
    state = []
    for name, ci, cm, co, lvl in node_shapes(cfg):
        for (i, o) in ((ci, cm), (cm, co)):
            state += [torch.randn(o, i, 3, 3, device=dev) * (2.0 / (9 * i)) ** 0.5, torch.zeros(o, device=dev), torch.ones(o, device=dev),
                      torch.zeros(o, device=dev), torch.zeros(o, device=dev), torch.ones(o, device=dev), torch.zeros((), dtype=torch.int64, device=dev)]
    for s in range(S):
        state += [torch.randn(2, f, 1, 1, device=dev) * 0.2, torch.zeros(2, device=dev)]
    grads = []
    for i, t in enumerate(state):
        k = i % 7 if i < 7 * 2 * len(node_shapes(cfg)) else 0
        grads.append(torch.zeros_like(t) if t.dtype == torch.float32 and k not in (4, 5) else None)
    plan.bind(state, grads)
    x = torch.rand(B, S, cin, H, W, device=dev)
    out = torch.empty(B, S, 2, H, W, device=dev)
    dout = torch.randn(B, S, 2, H, W, device=dev) * 1e-3
    return plan, x, out, dout


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "C2"
    training = "--eval" not in sys.argv
    plan, x, out, dout = build(name)
    plan.forward(x, out, training)
    plan.backward(dout)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.forward(x, out, training)
    plan.backward(dout)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("one step done", name, "launches", plan.last_launches)


if __name__ == "__main__":
    main()
