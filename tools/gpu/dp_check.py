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

    # ---- zero_grad(set_to_none=False): .grad stays aliased to the flat buffer -> the accumulate branch must all-reduce too
    sync = rt.enable_overlapped_allreduce()
    torch.manual_seed(50)
    for p_ in model.parameters():
        if p_.grad is not None:
            p_.grad.zero_()
    model.training_step({"image": image, "label": label}, 0)["loss"].backward()
    sync.wait()
    torch.cuda.synchronize()
    got = rt.flat_grads.clone()
    rt.grad_sync = None
    torch.manual_seed(50)
    model.zero_grad(set_to_none=True)
    model.training_step({"image": image, "label": label}, 0)["loss"].backward()
    ref = rt.flat_grads.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    err = float((got - ref).norm() / ref.norm())
    if rank == 0:
        print(f"set_to_none=False (accumulate branch): rel err vs plain all-reduce {err:.2e}", flush=True)
    assert err < 1e-4, err

    # ---- SURVEY 8e parity definition: R replicas of the ORACLE (per-replica BatchNorm statistics, gradients averaged) against
    # the R-GPU run of the product, same weights and shards; plus identical loss buffers on every rank after a few steps
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import mimo_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    world = dist.get_world_size()
    S, f = 2, 21
    sd = O.make_state_dict(3, 2, S, f, seed=4)
    m2 = MimoUnetModel(3, 2, S, f, 0.0, 0.0, 0.0, 0.0, 0.0, "laplace_nll", 0.0, 1e-3, 1, 10, 0.3, input_repetition_probability=1.0).to(dev)
    m2.model.load_state_dict(sd)
    m2.train()
    sync2 = m2.model.runtime().enable_overlapped_allreduce()
    torch.manual_seed(200 + rank)
    x, y = torch.rand(8, 3, 64, 96, device=dev), torch.rand(8, 1, 64, 96, device=dev)
    m2.zero_grad(set_to_none=True)
    step = m2.training_step({"image": x, "label": y}, 0)
    step["loss"].backward()
    sync2.wait()
    torch.cuda.synchronize()
    names = [n for n, _, _ in O.state_dict_spec(3, 2, S, f)]
    p = {k: (v.to(dev).clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.to(dev).clone()) for k, v in sd.items()}
    xs, ys = torch.stack([x] * S, dim=1), torch.stack([y] * S, dim=1)
    out = O.mimo_unet_forward(xs, p, S, training=True, emulate_bf16=True)
    loss_ref, total = O.train_loss(out, ys, None, torch.ones(S, device=dev))
    total.backward()
    cos = []
    for k in names:
        t = p[k]
        if not (isinstance(t, torch.Tensor) and t.requires_grad) or t.grad is None or k.endswith("double_conv.0.bias") or k.endswith("double_conv.3.bias"):
            continue
        g_ref = t.grad.clone()
        dist.all_reduce(g_ref, op=dist.ReduceOp.AVG)     # the R-replica average of the oracle's gradients
        g = dict(m2.model.named_parameters())[k].grad
        cos.append(float(torch.dot(g.flatten().double(), g_ref.flatten().double()) / (g.norm().double() * g_ref.norm().double() + 1e-30)))
    mean_cos, min_cos = sum(cos) / len(cos), min(cos)
    for _ in range(3):   # a few more steps: the loss buffers must stay identical on all ranks (side-stream loss exchange)
        m2.zero_grad(set_to_none=True)
        m2.training_step({"image": x, "label": y}, 0)["loss"].backward()
        sync2.wait()
    sync2.join_loss_exchange()
    torch.cuda.synchronize()
    buf = m2.loss_buffer.device_state(dev).buffer.clone()
    other = buf.clone()
    dist.broadcast(other, src=0)
    same_buf = bool(torch.equal(other, buf))
    if rank == 0:
        print(f"R-replica oracle ({world} replicas, bf16-emulating, per-replica BN): gradient cosine mean {mean_cos:.4f} min {min_cos:.4f}; "
              f"loss buffers identical across ranks {same_buf}", flush=True)
    assert mean_cos >= 0.97 and min_cos >= 0.8, (mean_cos, min_cos)
    assert same_buf
    if rank == 0:
        print("dp_check OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
