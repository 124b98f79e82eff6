"""CPU tests that PIN the oracle: against the reference itself (when /root/reference is mounted, i.e. in the
build container) and against the committed golden fixtures the reference generated (always)."""
import math

import pytest
import torch

from oracle import _refload, mimo_oracle as O

needs_ref = pytest.mark.skipif(not _refload.reference_available(), reason="reference checkout not mounted")


@pytest.fixture(scope="module")
def cases(golden_dir):
    return torch.load(f"{golden_dir}/model_cases.pt")


@pytest.fixture(scope="module")
def ops(golden_dir):
    return torch.load(f"{golden_dir}/op_cases.pt")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("name", ["m1_f8_32x32", "m2_f8_37x45", "m2_f21_32x48", "m4_f8_32x32", "m2_f30_c2_32x32"])
def test_oracle_forward_matches_golden(cases, name):
    c = cases[name]
    cfg = c["cfg"]
    sd = O.make_state_dict(cfg["cin"], 2, cfg["S"], cfg["f"], cfg["seed"])
    out = O.mimo_unet_forward(c["x"], sd, cfg["S"], training=False)
    assert (out - c["eval"]["out"]).abs().max() <= 1e-6          # eval: <= 1e-7-ish
    ns = {}
    out = O.mimo_unet_forward(c["x"], sd, cfg["S"], training=True, new_stats=ns)
    assert (out - c["train"]["out"]).abs().max() <= 1e-4         # train: BN summation order
    for k, v in c["train"]["new_stats"].items():
        assert torch.allclose(ns[k].float(), v.float(), atol=2e-6), k
    loss = O.laplace_nll_elementwise(out[:, :, :1], out[:, :, 1:], c["y"]).mean(dim=(0, 2, 3, 4))
    assert torch.allclose(loss, c["train"]["loss"], rtol=1e-5)


def test_oracle_gradients_match_golden(cases):
    c = cases["m2_f8_32x32"]
    cfg = c["cfg"]
    sd = O.make_state_dict(cfg["cin"], 2, cfg["S"], cfg["f"], cfg["seed"])
    p = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v) for k, v in sd.items()}
    x = c["x"].clone().requires_grad_(True)
    out = O.mimo_unet_forward(x, p, cfg["S"], training=False)
    l = O.laplace_nll_elementwise(out[:, :, :1], out[:, :, 1:], c["y"]).mean(dim=(0, 2, 3, 4))
    (l * c["w"]).mean().backward()
    assert rel(x.grad, c["eval"]["x_grad"]) <= 1e-4
    for k, dg in c["eval"]["grads"].items():
        d = O.grad_digest(p[k].grad)
        assert abs(float(d["norm"]) - float(dg["norm"])) <= 1e-3 * float(dg["norm"]) + 1e-7, k
        assert torch.allclose(d["head"], dg["head"], rtol=2e-3, atol=1e-6), k


def test_laplace_formulas_match_golden(ops):
    g = ops["loss"]
    l = O.laplace_nll_elementwise(g["mu"], g["log_s"], g["y"])
    assert torch.allclose(l, g["loss"], rtol=1e-6)
    gm, gl = O.laplace_nll_grads(g["mu"], g["log_s"], g["y"])
    assert torch.allclose(gm, g["g_mu"], rtol=1e-6) and torch.allclose(gl, g["g_ls"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(O.laplace_std(g["log_s"]), g["std"])
    assert torch.allclose(O.laplace_dist_param(g["std"], log=True), g["dist_param_log"], rtol=1e-6, atol=1e-6)
    # SURVEY App. C.5 probe values
    lm, gmu, gls = (f(torch.tensor([0.3]), torch.tensor([-20.0]), torch.tensor([0.0])) for f in
                    (O.laplace_nll_elementwise, lambda *a: O.laplace_nll_grads(*a)[0], lambda *a: O.laplace_nll_grads(*a)[1]))
    assert abs(float(lm) - 29988.488) < 0.01 and abs(float(gmu) - 1e5) < 1 and abs(float(gls) + 6.18326) < 1e-3


def test_loss_buffer_matches_golden(ops):
    g = ops["buffer"]
    for key, expect in g.items():
        if key == "losses":
            continue
        size, T = int(key.split("_")[0][4:]), float(key.split("_T")[1])
        lb = O.LossBufferOracle(3, T, size)
        got = []
        for t in range(25):
            got.append(lb.get_weights())
            lb.add(g["losses"][t])
        assert torch.allclose(torch.stack(got), expect, rtol=1e-6), key


def test_uncertainties_match_golden(ops):
    for key, g in ops["uncertainty"].items():
        m, a, e = O.compute_uncertainties(g["p1"], g["p2"])
        assert torch.allclose(m, g["mean"]) and torch.allclose(a, g["alea"], rtol=1e-6) and torch.allclose(e, g["epi"], rtol=1e-5, atol=1e-7)


def test_input_transform_matches_golden(ops):
    from mimo.models.utils import apply_input_transform
    for key, g in ops["transform"].items():
        p = float(key.split("_")[0][1:])
        rep = int(key.split("rep")[1])
        torch.manual_seed(123)
        img = torch.arange(6 * 2 * 2 * 2, dtype=torch.float32).reshape(6, 2, 2, 2)
        lab = torch.arange(6, dtype=torch.float32).reshape(6, 1, 1, 1).expand(6, 1, 2, 2).contiguous()
        a, b, _ = apply_input_transform(img, lab, None, 3, p, rep)   # product-side implementation, CPU tensors (index math only)
        assert torch.equal(a, g["image"]) and torch.equal(b, g["label"]), key  # permutation indices are bit exact


def test_flop_table_matches_survey():
    fwd, train = O.flops_per_sample(3, 2, 21, 256, 256)
    assert abs(fwd / 1e9 - 31.794) < 0.01 and abs(train / 1e9 - 95.234) < 0.01
    fwd, _ = O.flops_per_sample(3, 4, 21, 128, 160)
    assert abs(fwd / 1e9 - 39.999) < 0.01
    assert len(O.state_dict_spec(3, 2, 2, 21)) == 172


@needs_ref
@pytest.mark.parametrize("S,f,H,W", [(2, 8, 37, 45), (3, 8, 32, 32)])
def test_oracle_equals_reference_live(S, f, H, W):
    ref = _refload.load()
    torch.manual_seed(0)
    sd = O.make_state_dict(3, 2, S, f, seed=9)
    m = ref.model.MimoUNet(3, 2, S, f)
    assert list(m.state_dict().keys()) == [n for n, _, _ in O.state_dict_spec(3, 2, S, f)]
    m.load_state_dict(sd)
    x = torch.rand(2, S, 3, H, W)
    m.eval()
    with torch.no_grad():
        assert (m(x) - O.mimo_unet_forward(x, sd, S, training=False)).abs().max() <= 1e-6
    m.train()
    ns = {}
    with torch.no_grad():
        assert (m(x) - O.mimo_unet_forward(x, sd, S, training=True, new_stats=ns)).abs().max() <= 1e-4
    msd = m.state_dict()
    for k in msd:
        if "running" in k:
            assert torch.allclose(msd[k], ns[k], atol=2e-6), k
    # loss + gradient formulas against the reference's autograd
    crit = ref.losses.LaplaceNLL()
    mu = torch.randn(50, requires_grad=True)
    ls = (torch.randn(50) * 6).requires_grad_(True)
    y = torch.randn(50)
    crit.forward(mu, ls, y, reduce_mean=False).sum().backward()
    gm, gl = O.laplace_nll_grads(mu.detach(), ls.detach(), y)
    assert torch.allclose(mu.grad, gm, rtol=1e-6) and torch.allclose(ls.grad, gl, rtol=1e-5, atol=1e-6)


@needs_ref
def test_gaussian_and_evidential_oracle_equal_reference_live():
    """GaussianNLL (losses.py:39-121) and EvidentialLoss (losses.py:195-271): oracle restatement vs the reference classes,
    values and autograd gradients, including the clamp edge cases of the Gaussian variance."""
    ref = _refload.load()
    torch.manual_seed(3)
    mu = torch.randn(200, requires_grad=True)
    lv = torch.cat([torch.randn(180) * 4, torch.tensor([-20.0, -11.6, 0.0, 6.9, 10.0] * 4)]).requires_grad_(True)
    y = torch.randn(200)
    mask = (torch.rand(200) > 0.3).float()
    a = ref.losses.GaussianNLL().forward(mu, lv, y, mask=mask, reduce_mean=False)
    mu2, lv2 = mu.detach().clone().requires_grad_(True), lv.detach().clone().requires_grad_(True)
    b = O.gaussian_nll_elementwise(mu2, lv2, y, mask)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    a.sum().backward()
    b.sum().backward()
    assert torch.allclose(mu.grad, mu2.grad, rtol=1e-6, atol=1e-7) and torch.allclose(lv.grad, lv2.grad, rtol=1e-5, atol=1e-6)
    raw = torch.randn(3, 4, 5, 6, requires_grad=True)
    yy = torch.rand(3, 1, 5, 6)
    par = O.evidential_head(raw)
    mu_, logv, loga, logb = torch.unbind(raw, dim=1)
    sp = torch.nn.Softplus()
    assert torch.equal(par, torch.stack([mu_, sp(logv), sp(loga) + 1, sp(logb)], dim=1))   # evidential_unet.py:90-96
    la = ref.losses.EvidentialLoss(coeff=1.0).forward(par, yy, reduce_mean=False)
    assert torch.allclose(la, O.evidential_loss_elementwise(par, yy), rtol=1e-6, atol=1e-7)


@needs_ref
def test_product_surface_matches_reference_names():
    """state_dict layout and constructor surface of the product module vs the live reference."""
    ref = _refload.load()
    from mimo.models.mimo_components.model import MimoUNet
    for S, f in ((1, 8), (2, 21), (4, 8)):
        a, b = MimoUNet(3, 2, S, f), ref.model.MimoUNet(3, 2, S, f)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    with pytest.raises(ValueError):
        MimoUNet(3, 2, 2, 8, encoder_dropout_rate=0.1, center_dropout_rate=0.1)


@needs_ref
def test_oracle_elementwise_dropout_placement_equals_reference_live():
    """center_dropout (reference model.py:239) and final_dropouts (model.py:294): the keep masks the reference's nn.Dropout
    modules actually drew are captured with forward hooks (output / input) and replayed through the oracle's `elem_masks`."""
    ref = _refload.load()
    S, f, H, W = 2, 8, 32, 48
    sd = O.make_state_dict(3, 2, S, f, seed=4)
    m = ref.model.MimoUNet(3, 2, S, f, center_dropout_rate=0.3, final_dropout_rate=0.2)
    m.load_state_dict(sd)
    m.train()
    masks = {}

    def grab(name):
        def hook(mod, inp, out):
            x = inp[0]
            scale = 1.0 / (1.0 - mod.p)
            keep = torch.where(x != 0, out / (x * scale), torch.ones_like(x))  # x == 0: mask irrelevant
            masks[name] = keep.round() * scale
        return hook

    m.core.center_dropout.register_forward_hook(grab("center"))
    for i, d in enumerate(m.decoder.final_dropouts):
        d.register_forward_hook(grab(f"final.{i}"))
    torch.manual_seed(3)
    x = torch.rand(2, S, 3, H, W)
    with torch.no_grad():
        y_ref = m(x)
    assert set(masks) == {"center", "final.0", "final.1"}
    assert 0.05 < float((masks["center"] == 0).float().mean()) < 0.6
    y = O.mimo_unet_forward(x, sd, S, training=True, elem_masks=masks)
    assert (y - y_ref).abs().max() <= 1e-4
    y_plain = O.mimo_unet_forward(x, sd, S, training=True)
    assert (y_plain - y_ref).abs().max() > 1e-2  # the masks matter
