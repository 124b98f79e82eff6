"""Second baseline (SURVEY 8d "GPU reference timing"): the REFERENCE ITSELF on the same B200 through stock PyTorch (cuDNN / ATen
/ Inductor). Uses the reference's own modules (MimoUNet, apply_input_transform, LaplaceNLL, LossBuffer -- loaded by
oracle/_refload.py from /root/reference or from the copy staged into the git-ignored oracle/_ref/) in the reference's
Lightning-free training loop (notebook cell 13-14 == mimo_unet.py:115-144), C2 shape, with the speed knobs SURVEY 8d lists:
eager fp32, eager autocast(bf16) + channels_last, and torch.compile (what the reference's model.compile() does,
mimo_unet.py:89-91) under autocast(bf16), all with cudnn.benchmark.
Test / measurement infrastructure only (it executes oracle/). Usage: python tests/perf_torch_gpu_baseline.py [--steps 10]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from oracle import _refload  # noqa: E402


def run(mode: str, steps: int, warmup: int):
    R = _refload.load()
    dev = torch.device("cuda")
    S, f, cin, B, H, W = 2, 21, 3, 64, 128, 160
    torch.manual_seed(1)
    net = R.model.MimoUNet(in_channels=cin, out_channels=2, num_subnetworks=S, filter_base_count=f).to(dev)
    net.train()
    autocast = mode != "eager_fp32"
    if mode != "eager_fp32":
        net = net.to(memory_format=torch.channels_last)
    fwd = torch.compile(net) if mode == "compile_bf16" else net
    loss_fn = R.losses.LaplaceNLL()
    lb = R.loss_buffer.LossBuffer(subnetworks=S, temperature=0.3, buffer_size=10)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    x, y = torch.rand(B, cin, H, W, device=dev), torch.rand(B, 1, H, W, device=dev)

    def step():
        image_t, label_t, _ = R.utils.apply_input_transform(x, y, None, num_subnetworks=S, input_repetition_probability=0.0, batch_repetitions=1)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = fwd(image_t)
        out = out.float()
        loss = loss_fn.forward(out[:, :, :1], out[:, :, 1:], label_t, reduce_mean=False).mean(dim=(0, 2, 3, 4))
        w = lb.get_weights().to(dev)
        lb.add(loss.detach())
        opt.zero_grad(set_to_none=True)
        (loss * w).mean().backward()
        opt.step()

    t0 = time.perf_counter()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t_warm = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "warmup_s": round(t_warm, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    assert _refload.reference_available(), "reference modules not found (run oracle/stage_reference.py in the build container)"
    torch.backends.cudnn.benchmark = True
    res = []
    for mode in ("eager_fp32", "eager_bf16_channels_last", "compile_bf16"):
        try:
            res.append(run(mode, a.steps, a.warmup))
        except Exception as e:  # e.g. no compiler toolchain for Inductor on the box
            res.append({"mode": mode, "error": f"{type(e).__name__}: {str(e)[:300]}"})
    print(json.dumps({"workload": "C2 train step (M=2 fbc=21 3x128x160 batch 64), the reference's own modules through stock PyTorch on the same GPU",
                      "torch": torch.__version__, "cudnn_benchmark": True, "results": res}))


if __name__ == "__main__":
    main()
