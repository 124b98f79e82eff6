"""Per-node cost of dependent kernels inside a CUDA graph on this GPU: N tiny element-wise kernels captured back to back and replayed.
(Answers: how much of a 227-launch training step is launch gap, i.e. what programmatic dependent launch could win at most.)"""
import torch

x = torch.zeros(1024, device="cuda")
big = torch.zeros(64 * 1024 * 1024, device="cuda")
for name, fn, n in (("tiny add_ (1 block)", lambda: x.add_(1.0), 400), ("256 MB add_ (fills the GPU)", lambda: big.add_(1.0), 50)):
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    print(f"{name}: {best:.2f} us per graph node")
