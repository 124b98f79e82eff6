"""Where does the C2 training step spend GPU time outside the executor? torch.profiler (CUDA activities) over a few steps of
MimoUnetModel.training_step + backward + Adam; prints kernels grouped by name with per-step totals, and the GPU idle share."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimo.models.mimo_unet import MimoUnetModel  # noqa: E402


def main():
    dev = torch.device("cuda")
    torch.manual_seed(1)
    model = MimoUnetModel(3, 2, 2, 21, 0.0, 0.0, 0.0, 0.0, 0.0, "laplace_nll", 0.0, 1e-3, 1, 10, 0.3).to(dev)
    model.train()
    opt = model.configure_optimizers()["optimizer"]
    image, label = torch.rand(64, 3, 128, 160, device=dev), torch.rand(64, 1, 128, 160, device=dev)

    def step():
        out = model.training_step({"image": image, "label": label}, 0)
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 5
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            step()
        torch.cuda.synchronize()
    # host-side cost of enqueueing a step (the GPU queue is never awaited inside): if this is close to the step time the
    # step is launch-bound, not kernel-bound
    import time
    t_fwd = t_bwd = t_opt = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        a = time.perf_counter()
        out = model.training_step({"image": image, "label": label}, 0)
        b = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        c = time.perf_counter()
        opt.step()
        d = time.perf_counter()
        t_fwd += b - a; t_bwd += c - b; t_opt += d - c
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"# host enqueue per step {t_enq / 20 * 1e3:.3f} ms (training_step {t_fwd / 20 * 1e3:.3f}, backward {t_bwd / 20 * 1e3:.3f}, "
          f"adam {t_opt / 20 * 1e3:.3f}); wall per step incl. final sync {t_all / 20 * 1e3:.3f} ms")
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t / n, e.count / n, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"# GPU kernel time per step {tot / 1e3:.3f} ms over {sum(r[1] for r in rows):.0f} launches")
    ours = sum(r[0] for r in rows if "mimo" in r[2])
    print(f"# mimo_b200 kernels {ours / 1e3:.3f} ms, everything else {(tot - ours) / 1e3:.3f} ms")
    for t, c, k in rows[:60]:
        print(f"{t:9.1f} us {c:6.1f}x  {k[:110]}")


if __name__ == "__main__":
    main()
