"""GPU diagnostic: whole-model forward/backward of the executor vs the oracle (fp32 and bf16-emulating) run on
the same GPU with autograd. Prints per-tensor errors. Usage: python tests/diag_model_vs_oracle.py [S f H W B cin]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from mimo_unet_b200.engine import UNetPlan  # noqa: E402
from oracle import mimo_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def main():
    args = [int(v) for v in sys.argv[1:]]
    S, f, H, W, B, cin = (args + [2, 8, 32, 32, 2, 3][len(args):])[:6]
    training = True
    torch.manual_seed(0)
    sd = {k: v.cuda() for k, v in O.make_state_dict(cin, 2, S, f, 17).items()}
    x = torch.rand(B, S, cin, H, W, device="cuda")
    y = torch.rand(B, S, 1, H, W, device="cuda")
    names = [n for n, _, _ in O.state_dict_spec(cin, 2, S, f)]

    def oracle_run(emu):
        p = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone()) for k, v in sd.items()}
        rec = O.Recorder()
        ns = {}
        out = O.mimo_unet_forward(x, p, S, training=training, emulate_bf16=emu, new_stats=ns, rec=rec)
        l = O.laplace_nll_elementwise(out[:, :, :1], out[:, :, 1:], y).mean(dim=(0, 2, 3, 4))
        return p, out, l, rec, ns

    p32, out32, l32, rec32, ns32 = oracle_run(False)
    pemu, outemu, lemu, recemu, nsemu = oracle_run(True)
    w = torch.softmax(torch.arange(S, dtype=torch.float32, device="cuda") * 0.3, 0) * S

    plan = UNetPlan(cin, 2, S, f, B, H, W, torch.device("cuda"))
    state = [sd[n].clone().contiguous() for n in names]
    grads = [torch.full_like(t, float("nan")) if t.dtype == torch.float32 and "running" not in n else None for n, t in zip(names, state)]
    plan.bind(state, grads)
    out = torch.empty(B, S, 2, H, W, device="cuda")
    plan.forward(x, out, training)
    torch.cuda.synchronize()
    print(f"config S={S} f={f} {H}x{W} B={B} cin={cin}; launches fwd {plan.last_launches}")
    print("out: vs fp32 oracle %.3e  vs bf16 oracle %.3e   (bf16 oracle vs fp32 oracle %.3e)" % (rel(out, out32), rel(out, outemu), rel(outemu, out32)))
    l = O.laplace_nll_elementwise(out[:, :, :1], out[:, :, 1:], y).mean(dim=(0, 2, 3, 4))
    print("loss", l.tolist(), "fp32", l32.tolist(), "emu", lemu.tolist())
    # intermediates
    for key in recemu.t:
        if key.startswith("decoder.feat"):
            continue
        # key like 'encoder.in_convs.0.double_conv.0.raw'
        base, idx, kind = key.rsplit(".", 2)
        node = base.replace(".conv.double_conv", "").replace(".double_conv", "")
        cn = "c1" if idx == "0" else "c2"
        name = f"{node}.{cn}.y" if kind == "raw" else (f"{node}.a1" if idx == "0" else f"{node}.out")
        got = plan.debug_tensor(name)
        print("  %-44s vs bf16-oracle %.3e   vs fp32 %.3e" % (key, rel(got, recemu.t[key]), rel(got, rec32.t[key])))
    # running stats
    worst = 0
    for k in ns32:
        if k in sd and "num_batches" not in k:
            worst = max(worst, float((state[names.index(k)] - ns32[k]).abs().max() / (ns32[k].abs().max() + 1e-6)))
    print("running stats worst rel-max vs fp32 oracle: %.3e" % worst)

    # backward from the SAME upstream gradient (computed at the executor's output)
    o = out.clone().requires_grad_(True)
    (O.laplace_nll_elementwise(o[:, :, :1], o[:, :, 1:], y).mean(dim=(0, 2, 3, 4)) * w).mean().backward()
    dout = o.grad.contiguous()
    plan.backward(dout)
    torch.cuda.synchronize()
    print("launches bwd", plan.last_launches)
    out32.backward(dout)
    outemu.backward(dout)
    print("param grads:  name  | cos/rel vs fp32 oracle | cos/rel vs bf16 oracle | (bf16 oracle vs fp32: cos/rel)")
    for n, g in zip(names, grads):
        if g is None:
            continue
        g32, gem = p32[n].grad, pemu[n].grad
        if g32 is None:
            continue
        print("  %-46s %7.4f %.2e | %7.4f %.2e | %7.4f %.2e  |g|=%.3e" % (n, cos(g, g32), rel(g, g32), cos(g, gem), rel(g, gem), cos(gem, g32), rel(gem, g32), float(g.norm())))


if __name__ == "__main__":
    main()
