"""2-GPU check of the overlapped gradient all-reduce (stage events recorded between the backward stage graphs):
for several iterations (eager, then CUDA-graph replays) the bucketed, overlapped result must equal a plain
all-reduce(AVG) of the ranks' local gradients, and be identical on every rank.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gpu/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mimo.models.mimo_unet import MimoUnetModel  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)
    model = MimoUnetModel(3, 2, 2, 21, 0.0, 0.0, 0.0, 0.0, 0.0, "laplace_nll", 0.0, 1e-3, 1, 0, 0.3).to(dev)  # loss_buffer_size 0: weights == 1 in both passes
    model.train()
    rt = model.model.runtime()
    torch.manual_seed(100 + rank)
    image, label = torch.rand(16, 3, 64, 96, device=dev), torch.rand(16, 1, 64, 96, device=dev)
    worst = 0.0
    for it in range(6):
        # pass 1: local gradients, no synchroniser
        rt.grad_sync = None
        torch.manual_seed(7 + it)  # same shuffle in both passes
        model.zero_grad(set_to_none=True)
        model.training_step({"image": image, "label": label}, 0)["loss"].backward()
        ref = rt.flat_grads.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        # pass 2: overlapped bucketed all-reduce
        sync = rt.enable_overlapped_allreduce()
        torch.manual_seed(7 + it)
        model.zero_grad(set_to_none=True)
        model.training_step({"image": image, "label": label}, 0)["loss"].backward()
        sync.wait()
        torch.cuda.synchronize()
        got = rt.flat_grads
        err = float((got - ref).norm() / ref.norm())
        other = got.clone()
        dist.broadcast(other, src=0)
        same = bool(torch.equal(other, got))
        plan = next(iter(rt.plans.values()))
        if rank == 0:
            print(f"iter {it}: rel err vs plain all-reduce {err:.2e}, identical across ranks {same}, graph_state {plan.graph_state:#x}", flush=True)
        worst = max(worst, err)
        assert same, "ranks disagree after the overlapped all-reduce"
    assert worst < 1e-4, worst
    if rank == 0:
        print("dp_check OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
