"""GPU parity of the fused memory-bound kernels (Laplace NLL fwd/bwd, loss buffer, fused training loss, ensemble
aggregation) against the reference-generated golden fixtures and the oracle. fp32 kernels: tolerance 1e-5 rel."""
import pytest
import torch

from mimo_unet_b200 import functional as Fn
from oracle import mimo_oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(golden_dir):
    return torch.load(f"{golden_dir}/op_cases.pt")


def test_laplace_grid_golden(ops):
    g = ops["loss"]
    mu = g["mu"].cuda().requires_grad_(True)
    ls = g["log_s"].cuda().requires_grad_(True)
    y = g["y"].cuda()
    l = Fn.laplace_nll(mu, ls, y, reduce_mean=False)
    assert torch.allclose(l.cpu(), g["loss"], rtol=2e-6, atol=1e-6)
    l.sum().backward()
    assert torch.allclose(mu.grad.cpu(), g["g_mu"], rtol=2e-6, atol=1e-7)
    assert torch.allclose(ls.grad.cpu(), g["g_ls"], rtol=1e-5, atol=1e-6)
    lm = Fn.laplace_nll(mu.detach(), ls.detach(), y, mask=g["mask"].cuda(), reduce_mean=True)
    assert torch.allclose(lm.cpu(), g["loss_masked_mean"], rtol=1e-5)
    # probe values quoted in SURVEY App. C.5
    one = Fn.laplace_nll(torch.tensor([0.3], device="cuda"), torch.tensor([-20.0], device="cuda"), torch.zeros(1, device="cuda"), reduce_mean=False)
    assert abs(float(one) - 29988.488) < 0.05


def test_laplace_strided_views_and_mean_backward():
    torch.manual_seed(0)
    out = torch.randn(3, 2, 2, 17, 19, device="cuda", requires_grad=True)
    y = torch.rand(3, 2, 1, 17, 19, device="cuda")
    p1, p2 = out[:, :, :1], out[:, :, 1:]
    l = Fn.laplace_nll(p1, p2, y, reduce_mean=True)
    l.backward()
    oc = out.detach().cpu().requires_grad_(True)
    lr = O.laplace_nll_elementwise(oc[:, :, :1], oc[:, :, 1:], y.cpu()).mean()
    lr.backward()
    assert abs(float(l) - float(lr)) <= 1e-5 * abs(float(lr))
    assert rel_l2(out.grad.cpu(), oc.grad) <= 1e-5
    # elementwise + per-subnetwork mean, as the reference training loop does it
    out.grad = None
    le = Fn.laplace_nll(p1, p2, y, reduce_mean=False).mean(dim=(0, 2, 3, 4))
    (le * torch.tensor([0.4, 1.6], device="cuda")).mean().backward()
    oc.grad = None
    lre = O.laplace_nll_elementwise(oc[:, :, :1], oc[:, :, 1:], y.cpu()).mean(dim=(0, 2, 3, 4))
    (lre * torch.tensor([0.4, 1.6])).mean().backward()
    assert rel_l2(le.cpu(), lre) <= 1e-5 and rel_l2(out.grad.cpu(), oc.grad) <= 1e-5


def test_loss_buffer_trajectories_golden(ops):
    g = ops["buffer"]
    losses = g["losses"].cuda()
    for key, expect in g.items():
        if key == "losses":
            continue
        size = int(key.split("_")[0][4:])
        T = float(key.split("_T")[1])
        lb = Fn.DeviceLossBuffer(3, T, size, "cuda")
        got = []
        for t in range(losses.shape[0]):
            got.append(lb.weights().cpu())
            lb.add(losses[t])
        got = torch.stack(got)
        assert torch.allclose(got, expect, rtol=1e-5, atol=1e-6), key
        assert lb.index == (losses.shape[0] % size if size else 0)  # index sequence is bit exact


def test_fused_train_loss_matches_oracle_trajectory():
    torch.manual_seed(1)
    B, S, H, W = 4, 2, 33, 47
    lb = Fn.DeviceLossBuffer(S, 0.3, 10, "cuda")
    ob = O.LossBufferOracle(S, 0.3, 10)
    for step in range(4):
        out = torch.randn(B, S, 2, H, W, device="cuda", requires_grad=True)
        y = torch.rand(B, S, 1, H, W, device="cuda")
        mask = (torch.rand(B, S, 1, H, W, device="cuda") > 0.2).float() if step % 2 else None
        total, loss, w = Fn.laplace_train_loss(out, y, mask=mask, loss_buffer=lb)
        (total * 3.0).backward()
        oc = out.detach().cpu().requires_grad_(True)
        w_ref = ob.get_weights()
        l_ref, tot_ref = O.train_loss(oc, y.cpu(), None if mask is None else mask.cpu(), w_ref)
        ob.add(l_ref)
        (tot_ref * 3.0).backward()
        assert torch.allclose(w.cpu(), w_ref, rtol=1e-5), step
        assert rel_l2(loss.cpu(), l_ref.detach()) <= 1e-5
        assert abs(float(total) - float(tot_ref)) <= 1e-5 * abs(float(tot_ref))
        assert rel_l2(out.grad.cpu(), oc.grad) <= 1e-5
    assert torch.allclose(lb.buffer.cpu(), ob.buf, rtol=1e-5)


def test_fused_train_loss_gather_and_broadcast_labels():
    torch.manual_seed(2)
    B, S, H, W = 5, 3, 8, 12
    out = torch.randn(B, S, 2, H, W, device="cuda")
    label = torch.rand(B, 1, H, W, device="cuda")
    gather = torch.stack([torch.randperm(B) for _ in range(S)]).cuda()
    total, loss, w = Fn.laplace_train_loss(out, label, gather=gather, loss_buffer=None)
    y_t = torch.stack([label[gather[s]] for s in range(S)], dim=1)
    l_ref, tot_ref = O.train_loss(out.cpu(), y_t.cpu(), None, torch.ones(S))
    assert rel_l2(loss.cpu(), l_ref) <= 1e-5 and torch.all(w == 1)
    # validation-style broadcast (repeat_subnetworks without materialising the repeat)
    total2, loss2, _ = Fn.laplace_train_loss(out, label[:, None].expand(-1, S, -1, -1, -1), loss_buffer=None)
    l_ref2, _ = O.train_loss(out.cpu(), label[:, None].expand(-1, S, -1, -1, -1).cpu(), None, torch.ones(S))
    assert rel_l2(loss2.cpu(), l_ref2) <= 1e-5


def test_ensemble_aggregate_golden(ops):
    for key, g in ops["uncertainty"].items():
        m, a, e = Fn.ensemble_aggregate(g["p1"].cuda(), g["p2"].cuda())
        assert torch.allclose(m.cpu(), g["mean"], rtol=1e-5, atol=1e-6), key
        assert torch.allclose(a.cpu(), g["alea"], rtol=1e-5, atol=1e-6), key
        assert torch.allclose(e.cpu(), g["epi"], rtol=1e-4, atol=1e-6), key
    # strided p1/p2 views of a network output
    out = torch.randn(2, 4, 2, 9, 7, device="cuda")
    m, a, e = Fn.ensemble_aggregate(out[:, :, :1], out[:, :, 1:])
    mr, ar, er = O.compute_uncertainties(out[:, :, :1].cpu(), out[:, :, 1:].cpu())
    assert torch.allclose(m.cpu(), mr, atol=1e-6) and torch.allclose(a.cpu(), ar, rtol=1e-5) and torch.allclose(e.cpu(), er, rtol=1e-4, atol=1e-6)


def test_fused_regression_metrics_match_torch_formulas():
    """r2 / mae / mse / rmse of (mu, y) from the fused loss pass (reference mimo/metrics.py:22-34 computes them with four
    torchmetrics reductions per step) against the plain torch formulas in fp64; with a mask (which must not enter) and a gather."""
    from mimo_unet_b200 import functional as Fn
    torch.manual_seed(11)
    B, S, H, W = 6, 3, 37, 45
    out = torch.randn(B, S, 2, H, W, device="cuda", requires_grad=True)
    y = torch.rand(B, 1, H, W, device="cuda") * 3 + 0.5
    mask = (torch.rand(B, 1, H, W, device="cuda") > 0.3).float()
    gather = torch.stack([torch.randperm(B, device="cuda") for _ in range(S)])
    total, loss, weights, m = Fn.laplace_train_loss(out, y, mask=mask, gather=gather, with_metrics=True)
    total.backward()
    yt = torch.stack([y[g] for g in gather], dim=1).double()        # [B,S,1,H,W]
    mu = out.detach()[:, :, :1].double()
    e = mu - yt
    mse = (e * e).mean()
    ref = {"mae": e.abs().mean(), "mse": mse, "rmse": mse.sqrt(), "r2": 1 - (e * e).sum() / ((yt - yt.mean()) ** 2).sum()}
    assert list(m.keys()) == ["r2", "mae", "mse", "rmse"]           # the reference's logging order
    for k, v in ref.items():
        assert abs(float(m[k]) - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (k, float(m[k]), float(v))
    # the loss itself is unchanged by the extra accumulators
    total2, loss2, _ = Fn.laplace_train_loss(out.detach(), y, mask=mask, gather=gather)
    assert torch.equal(loss, loss2) and torch.equal(total.detach(), total2)


def test_gaussian_nll_vs_oracle_values_and_gradients():
    """GaussianNLL on the GPU (same kernels as the Laplace loss, Gaussian element math) against the oracle's autograd, incl. the
    clamp edge cases (the clamp changes the value the loss is evaluated at but is invisible to autograd), masks, mean and
    elementwise reductions, and the fused training pass (per-subnetwork means + gradient seed)."""
    from mimo.losses import GaussianNLL, UncertaintyLoss
    from mimo_unet_b200 import functional as Fn
    torch.manual_seed(4)
    crit = UncertaintyLoss.from_name("gaussian_nll")
    assert isinstance(crit, GaussianNLL)
    mu = torch.randn(4, 3, 1, 9, 11, device="cuda", requires_grad=True)
    lv = (torch.randn(4, 3, 1, 9, 11, device="cuda") * 5).requires_grad_(True)
    with torch.no_grad():
        lv.view(-1)[:5] = torch.tensor([-20.0, -11.6, 0.0, 6.9, 10.0], device="cuda")
    y = torch.randn(4, 3, 1, 9, 11, device="cuda")
    mask = (torch.rand(4, 3, 1, 9, 11, device="cuda") > 0.25).float()
    for reduce_mean in (False, True):
        mu_o, lv_o = mu.detach().double().requires_grad_(True), lv.detach().double().requires_grad_(True)
        ref = O.gaussian_nll_elementwise(mu_o, lv_o, y.double(), mask.double())
        got = crit(mu, lv, y, mask=mask, reduce_mean=reduce_mean)
        if reduce_mean:
            ref = ref.mean()
        assert rel_l2(got.detach(), ref.detach()) <= 1e-5
        mu.grad = lv.grad = None
        (got.sum() * 1.5).backward()
        (ref.sum() * 1.5).backward()
        assert rel_l2(mu.grad, mu_o.grad) <= 1e-5 and rel_l2(lv.grad, lv_o.grad) <= 1e-5
    # fused training pass with the Gaussian element math
    out = torch.randn(4, 3, 2, 9, 11, device="cuda", requires_grad=True)
    total, loss, w = Fn.laplace_train_loss(out, y, gaussian=True)
    total.backward()
    o2 = out.detach().double().requires_grad_(True)
    l_ref = O.gaussian_nll_elementwise(o2[:, :, :1], o2[:, :, 1:], y.double()).mean(dim=(0, 2, 3, 4))
    l_ref.mean().backward()
    assert rel_l2(loss, l_ref.detach()) <= 1e-5 and rel_l2(out.grad, o2.grad) <= 1e-5


def test_evidential_head_and_loss_vs_oracle():
    """Softplus head (evidential_unet.py:85-96) and EvidentialLoss (losses.py:195-271) fused kernels against the oracle in fp64:
    values, gradients w.r.t. the raw network output, mask, mean / elementwise."""
    from mimo.losses import EvidentialLoss
    from mimo_unet_b200 import functional as Fn
    torch.manual_seed(5)
    raw = (torch.randn(3, 4, 13, 17, device="cuda") * 1.5).requires_grad_(True)
    with torch.no_grad():
        raw[0, 1, 0, :3] = torch.tensor([25.0, -15.0, 0.0], device="cuda")   # softplus threshold / tiny evidence
    y = torch.rand(3, 1, 13, 17, device="cuda")
    mask = (torch.rand(3, 1, 13, 17, device="cuda") > 0.3).float()
    crit = EvidentialLoss(coeff=1.0)
    for reduce_mean in (False, True):
        par = Fn.evidential_head(raw)
        got = crit(par, y, mask=mask, reduce_mean=reduce_mean)
        r64 = raw.detach().double().requires_grad_(True)
        p64 = O.evidential_head(r64)
        ref = O.evidential_loss_elementwise(p64, y.double(), mask.double().squeeze(1))
        if reduce_mean:
            ref = ref.mean()
        assert rel_l2(par.detach(), p64.detach()) <= 1e-6
        assert rel_l2(got.detach(), ref.detach()) <= 2e-5
        raw.grad = None
        got.sum().backward()
        ref.sum().backward()
        assert rel_l2(raw.grad, r64.grad) <= 2e-4    # digamma series + fp32 lgamma against fp64
    assert torch.equal(crit.mode(par), par[:, 0])
    assert rel_l2(crit.aleatoric_var(par), par[:, 3] / (par[:, 2] - 1)) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("S,C,with_mask", [(2, 1, True), (1, 1, False), (4, 2, True), (16, 1, False)])
def test_fused_validation_pass_matches_oracle(S, C, with_mask):
    """MimoUnetModel.validation_step math (reference mimo_unet.py:146-183) in one kernel pass vs the oracle's restatement:
    per-subnetwork NLL means, ensemble aggregation, combined-scale NLL, the four maps, regression metrics, clipped std means.
    p1 / p2 are strided views of one [B, S, 2C, H, W] tensor like the module's output; log-scales reach both clamp bounds."""
    torch.manual_seed(21)
    B, H, W = 3, 19, 23
    out = torch.randn(B, S, 2 * C, H, W, device="cuda")
    out[:, :, C:] = out[:, :, C:] * 4.0              # exp() spans ~1e-7 .. 1e7: both clamps (1e-5, 1e3) are active somewhere
    label = torch.rand(B, C, H, W, device="cuda")
    mask = (torch.rand(B, C, H, W, device="cuda") > 0.2).float() if with_mask else None
    p1, p2 = out[:, :, :C], out[:, :, C:]
    v = Fn.validation_laplace(p1, p2, label, mask)
    lab5 = label[:, None].expand(-1, S, -1, -1, -1)
    vl, comb, mean, alea, epi = O.validation_math(out.cpu().double(), lab5.cpu().double(), None if mask is None else mask.cpu().double())
    assert rel_l2(v["val_loss"].cpu(), vl.float()) <= 1e-5
    assert abs(float(v["val_loss_combined"]) - float(comb)) <= 1e-5 * abs(float(comb))
    assert rel_l2(v["preds"].cpu(), mean.float()) <= 1e-6
    assert rel_l2(v["aleatoric_std"].cpu(), alea.sqrt().float()) <= 1e-5
    assert rel_l2(v["epistemic_std"].cpu(), epi.sqrt().float()) <= 1e-5
    assert rel_l2(v["err"].cpu(), (mean - label.cpu().double()).float()) <= 1e-5
    e = (mean - label.cpu().double()).flatten()
    y = label.cpu().double().flatten()
    ref = {"mae": e.abs().mean(), "mse": (e * e).mean(), "rmse": (e * e).mean().sqrt(),
           "r2": 1.0 - (e * e).sum() / ((y - y.mean()) ** 2).sum()}
    for k, r in ref.items():
        assert abs(float(v["metrics"][k]) - float(r)) <= 1e-4 * max(1.0, abs(float(r))), k
    assert abs(float(v["aleatoric_std_mean"]) - float(alea.sqrt().clip(0, 5).mean())) <= 1e-4
    assert abs(float(v["epistemic_std_mean"]) - float(epi.sqrt().clip(0, 5).mean())) <= 1e-4
