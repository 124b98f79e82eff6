"""Second baseline (SURVEY 8d "GPU reference timing"): the reference's algorithm through stock PyTorch on the SAME B200 --
the oracle's functional restatement (F.conv2d / batch_norm / max_pool / interpolate = cuDNN + ATen kernels), fp32 and
torch.autocast(bf16), eager; C2 training step (forward + Laplace NLL + loss-buffer weights + backward + fused Adam).
Test/measurement infrastructure only (imports oracle/). Lives under tests/ because only test infrastructure may execute oracle/. Usage: python tests/perf_torch_gpu_baseline.py [--steps 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from oracle import mimo_oracle as O  # noqa: E402


def run(autocast: bool, steps: int, warmup: int, channels_last: bool):
    dev = torch.device("cuda")
    S, f, cin, B, H, W = 2, 21, 3, 64, 128, 160
    torch.manual_seed(1)
    sd = O.make_state_dict(cin, 2, S, f, seed=1)
    params = {k: (v.to(dev).clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.to(dev).clone())
              for k, v in sd.items()}
    if channels_last:
        for k, v in params.items():
            if v.dim() == 4:
                params[k] = v.detach().contiguous(memory_format=torch.channels_last).requires_grad_(v.requires_grad)
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3, fused=True)
    x, y = torch.rand(B, cin, H, W, device=dev), torch.rand(B, 1, H, W, device=dev)
    w = torch.ones(S, device=dev)

    def step():
        idx = [torch.randperm(B, device=dev) for _ in range(S)]
        xs = torch.stack([x[i] for i in idx], dim=1)
        ys = torch.stack([y[i] for i in idx], dim=1)
        ns = {}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = O.mimo_unet_forward(xs, params, S, training=True, emulate_bf16=False, new_stats=ns)
        loss, total = O.train_loss(out.float(), ys, None, w)
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"autocast_bf16": autocast, "channels_last_weights": channels_last, "ms_per_step": ms, "images_per_s": B / ms * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    res = [run(False, a.steps, a.warmup, False), run(True, a.steps, a.warmup, False), run(True, a.steps, a.warmup, True)]
    print(json.dumps({"workload": "C2 train step, stock PyTorch (cuDNN/ATen) on the same GPU, eager", "torch": torch.__version__, "results": res}))


if __name__ == "__main__":
    main()
