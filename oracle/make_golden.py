"""Generates tests/golden/*.pt by RUNNING THE REFERENCE (/root/reference) on CPU, fp32.

Run in the build container only:  python oracle/make_golden.py
The fixtures are data (tensors), not code; the GPU box checks the CUDA path against them.
Weights are NOT stored: they are regenerated from ``oracle.mimo_oracle.make_state_dict(seed)``.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _refload, mimo_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def model_case(ref, S, f, H, W, B, seed, cin=3):
    torch.manual_seed(seed)
    sd = O.make_state_dict(cin, 2, S, f, seed)
    x = torch.rand(B, S, cin, H, W)
    y = torch.rand(B, S, 1, H, W)
    w = torch.softmax(torch.arange(S, dtype=torch.float32) * 0.3, 0) * S  # arbitrary loss weights
    case = {"cfg": dict(S=S, f=f, H=H, W=W, B=B, seed=seed, cin=cin), "x": x, "y": y, "w": w}
    crit = ref.losses.LaplaceNLL()
    for mode in ("train", "eval"):
        m = ref.model.MimoUNet(cin, 2, S, f)
        m.load_state_dict(sd)
        m.train(mode == "train")
        xin = x.clone().requires_grad_(True)
        out = m(xin)
        p1, p2 = out[:, :, :1], out[:, :, 1:]
        loss = crit.forward(p1, p2, y, reduce_mean=False).mean(dim=(0, 2, 3, 4))
        (loss * w).mean().backward()
        rec = {"out": out.detach().clone(), "loss": loss.detach().clone(), "x_grad": xin.grad.clone(),
               "grads": {k: O.grad_digest(p.grad) for k, p in m.named_parameters()}}
        if mode == "train":
            rec["new_stats"] = {k: v.clone() for k, v in m.state_dict().items()
                                if "running_" in k or "num_batches" in k}
        case[mode] = rec
    return case


def component_cases(ref):
    torch.manual_seed(11)
    out = {}
    # Down with pooling indices (components.py:36-57), odd size -> floor mode
    x = torch.rand(2, 5, 9, 11)
    x[0, 0, 0, 0] = x[0, 0, 0, 1] = 2.0  # tie -> first max wins
    d = ref.components.Down(5, 6, use_pooling_indices=True)
    pooled, idx = d.maxpool(x)
    out["maxpool"] = {"x": x, "pooled": pooled, "idx": idx}
    # MaxUnpool2d (components.py:87)
    xe = torch.rand(2, 5, 8, 12)
    pe, ie = torch.nn.functional.max_pool2d(xe, 2, return_indices=True)
    out["unpool"] = {"x": pe, "idx": ie, "y": torch.nn.MaxUnpool2d(2)(pe, ie)}
    # bilinear x2 align_corners (components.py:78) + pad to odd skip (components.py:112-115)
    xl = torch.rand(2, 6, 4, 5)
    up = torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)(xl)
    out["bilinear"] = {"x": xl, "y": up, "y_pad_9x11": O.pad_to(up, 9, 11)}
    # ConvTranspose2d k2 s2 (components.py:96-98)
    ct = torch.nn.ConvTranspose2d(6, 3, kernel_size=2, stride=2)
    out["convtranspose"] = {"x": xl, "w": ct.weight.detach().clone(), "b": ct.bias.detach().clone(),
                            "y": ct(xl).detach()}
    # Up module whole (bilinear) with odd skip
    upm = ref.components.Up(10, 4, bilinear=True)
    upm.train()
    x1, x2 = torch.rand(2, 6, 4, 5), torch.rand(2, 4, 9, 11)
    out["up_module"] = {"x1": x1, "x2": x2, "sd": {k: v.clone() for k, v in upm.state_dict().items()},
                        "y": upm(x1, x2).detach()}
    return out


def loss_cases(ref):
    crit = ref.losses.LaplaceNLL()
    log_s = torch.tensor([-20.0, -11.6, -11.5, 0.0, 0.5, 6.9, 6.91, 10.0])
    d = torch.tensor([-0.4, 0.0, 0.3, 1.7])
    LS, D = torch.meshgrid(log_s, d, indexing="ij")
    mu = D.clone().requires_grad_(True)
    ls = LS.clone().requires_grad_(True)
    y = torch.zeros_like(D)
    l = crit.forward(mu, ls, y, reduce_mean=False)
    l.sum().backward()
    mask = (torch.arange(l.numel()).reshape(l.shape) % 3 != 0).float()
    out = {"mu": D, "log_s": LS, "y": y, "loss": l.detach(), "g_mu": mu.grad.clone(), "g_ls": ls.grad.clone(),
           "mask": mask, "loss_masked_mean": crit.forward(D, LS, y, mask=mask).detach(),
           "std": crit.std(D, LS), "dist_param_log": crit.calculate_dist_param(crit.std(D, LS), log=True)}
    return out


def buffer_cases(ref):
    out = {}
    g = torch.Generator().manual_seed(5)
    losses = torch.rand(25, 3, generator=g) * 2
    for size, T in ((10, 0.3), (10, 1.0), (0, 1.0), (4, 0.5)):
        lb = ref.loss_buffer.LossBuffer(subnetworks=3, temperature=T, buffer_size=size)
        ws = []
        for t in range(25):
            ws.append(lb.get_weights().clone())
            lb.add(losses[t])
        out[f"size{size}_T{T}"] = torch.stack(ws)
    out["losses"] = losses
    return out


def uncertainty_cases(ref):
    crit = ref.losses.LaplaceNLL()
    out = {}
    g = torch.Generator().manual_seed(9)
    for S in (1, 2, 4, 8):
        p1 = torch.randn(2, S, 1, 6, 7, generator=g)
        p2 = torch.randn(2, S, 1, 6, 7, generator=g) * 0.5
        m, a, e = ref.utils.compute_uncertainties(crit, p1, p2)
        out[f"S{S}"] = {"p1": p1, "p2": p2, "mean": m, "alea": a, "epi": e}
    return out


def transform_cases(ref):
    out = {}
    for p in (0.0, 0.5, 1.0):
        for rep in (1, 2):
            torch.manual_seed(123)
            img = torch.arange(6 * 2 * 2 * 2, dtype=torch.float32).reshape(6, 2, 2, 2)
            lab = torch.arange(6, dtype=torch.float32).reshape(6, 1, 1, 1).expand(6, 1, 2, 2).contiguous()
            a, b, _ = ref.utils.apply_input_transform(img, lab, None, 3, p, rep)
            out[f"p{p}_rep{rep}"] = {"image": a, "label": b}
    return out


def main():
    assert _refload.reference_available(), "reference checkout not found"
    ref = _refload.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    models = {}
    for name, (S, f, H, W, B, cin) in {
        "m1_f8_32x32": (1, 8, 32, 32, 2, 3),
        "m2_f8_32x32": (2, 8, 32, 32, 2, 3),
        "m2_f8_37x45": (2, 8, 37, 45, 2, 3),
        "m2_f21_32x48": (2, 21, 32, 48, 2, 3),
        "m4_f8_32x32": (4, 8, 32, 32, 2, 3),
        "m2_f30_c2_32x32": (2, 30, 32, 32, 2, 2),
    }.items():
        models[name] = model_case(ref, S, f, H, W, B, seed=17, cin=cin)
        print("model case", name, models[name]["train"]["loss"])
    torch.save(models, os.path.join(OUT, "model_cases.pt"))
    torch.save({"components": component_cases(ref), "loss": loss_cases(ref), "buffer": buffer_cases(ref),
                "uncertainty": uncertainty_cases(ref), "transform": transform_cases(ref)},
               os.path.join(OUT, "op_cases.pt"))
    print("written to", OUT)


if __name__ == "__main__":
    main()
