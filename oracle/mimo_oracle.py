"""CPU oracle for the MIMO U-Net hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain ``torch.nn.functional`` on CPU) of the
algorithm the reference implements for the north-star path.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``mimo_unet_b200/`` or
``mimo/`` imports it, and the product path fails loudly when the CUDA library is
missing.

Parity status: PINNED against the reference itself.  The reference ships no tests or
golden vectors, so the pin is (1) ``tests/test_oracle_vs_reference.py``,
which runs the real reference modules from ``/root/reference`` in the build container
(auto-skipped where the reference is absent) and (2) the committed fixtures under
``tests/golden/`` that ``oracle/make_golden.py`` generated from the reference.

Every function cites the reference file:line it restates (paths relative to the
reference checkout).

Two arithmetic modes:
  * ``emulate_bf16=False``  fp32 everywhere: equals the reference up to BN summation order.
  * ``emulate_bf16=True``   rounds tensors to bf16 at exactly the points where the CUDA
    path stores bf16 (image, conv weights, raw conv outputs, activations, upsampled
    maps); accumulation stays fp32.  This is what the rel <= 1e-3 kernel bar is judged with.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------
# bf16 storage emulation
# ----------------------------------------------------------------------------
def round_bf16(x: Tensor) -> Tensor:
    """Value-rounds to bf16 and returns fp32 (straight-through for autograd)."""
    r = x.detach().to(torch.bfloat16).to(x.dtype)
    if x.requires_grad:
        return x + (r - x.detach())
    return r


def _q(x: Tensor, on: bool) -> Tensor:
    return round_bf16(x) if on else x


# ----------------------------------------------------------------------------
# building blocks (components.py)
# ----------------------------------------------------------------------------
def conv3x3_reflect(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """components.py:23,26 -- Conv2d(k=3, padding=1, padding_mode='reflect')."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)


def batchnorm_train(y: Tensor, gamma: Tensor, beta: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """components.py:24,27 -- BatchNorm2d in training mode.  Returns (out, mean, biased var)."""
    mean = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False)
    inv = torch.rsqrt(var + BN_EPS)
    out = (y - mean[None, :, None, None]) * (inv * gamma)[None, :, None, None] + beta[None, :, None, None]
    return out, mean, var


def batchnorm_eval(y: Tensor, gamma: Tensor, beta: Tensor, rm: Tensor, rv: Tensor) -> Tensor:
    inv = torch.rsqrt(rv + BN_EPS)
    return (y - rm[None, :, None, None]) * (inv * gamma)[None, :, None, None] + beta[None, :, None, None]


def updated_running_stats(rm: Tensor, rv: Tensor, mean: Tensor, var_b: Tensor, n: int):
    """momentum 0.1, unbiased variance for the running estimate."""
    unbiased = var_b * (n / max(n - 1, 1))
    return (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean, (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unbiased


def maxpool2x2(x: Tensor, return_indices: bool = False):
    """components.py:48 -- MaxPool2d(2): floor mode, int64 flat indices h*W+w, first max wins."""
    return F.max_pool2d(x, 2, return_indices=return_indices)


def upsample_bilinear2x_ac(x: Tensor) -> Tensor:
    """components.py:78 -- Upsample(scale_factor=2, bilinear, align_corners=True)."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def pad_to(x1: Tensor, H: int, W: int) -> Tensor:
    """components.py:109-115 -- zero pad the up-sampled map to the skip's size."""
    dY, dX = H - x1.shape[2], W - x1.shape[3]
    return F.pad(x1, [dX // 2, dX - dX // 2, dY // 2, dY - dY // 2])


class Recorder:
    """Collects intermediates (by state_dict-style name) for teacher-forced kernel tests."""

    def __init__(self):
        self.t: Dict[str, Tensor] = {}

    def put(self, name: str, v: Tensor):
        self.t[name] = v


def double_conv(
    x: Tensor,
    sd: Dict[str, Tensor],
    prefix: str,
    training: bool,
    emulate_bf16: bool,
    new_stats: Optional[Dict[str, Tensor]],
    drop_mask: Optional[Tensor] = None,
    rec: Optional[Recorder] = None,
) -> Tensor:
    """components.py:8-33.  ``prefix`` ends with 'double_conv.'.

    drop_mask: optional [N, C] keep-mask already scaled by 1/(1-p) (Dropout2d semantics,
    SURVEY App. C.3) applied after the second ReLU.
    """
    h = x
    for ci, bi in ((0, 1), (3, 4)):
        w = _q(sd[f"{prefix}{ci}.weight"], emulate_bf16)
        b = sd[f"{prefix}{ci}.bias"]
        gamma, beta = sd[f"{prefix}{bi}.weight"], sd[f"{prefix}{bi}.bias"]
        if training:
            # The CUDA path leaves the conv bias out of the stored raw output (it cancels in
            # train-mode BN) and adds it back only into running_mean.
            y = conv3x3_reflect(h, w, None)
            y = _q(y, emulate_bf16)
            if rec is not None:
                rec.put(f"{prefix}{ci}.raw", y)
            out, mean, var = batchnorm_train(y, gamma, beta)
            if new_stats is not None:
                n = y.shape[0] * y.shape[2] * y.shape[3]
                rm, rv = updated_running_stats(
                    sd[f"{prefix}{bi}.running_mean"], sd[f"{prefix}{bi}.running_var"], (mean + b).detach(), var.detach(), n
                )
                new_stats[f"{prefix}{bi}.running_mean"] = rm
                new_stats[f"{prefix}{bi}.running_var"] = rv
                new_stats[f"{prefix}{bi}.num_batches_tracked"] = sd[f"{prefix}{bi}.num_batches_tracked"] + 1
                new_stats[f"{prefix}{bi}.batch_mean"] = mean.detach()
                new_stats[f"{prefix}{bi}.batch_var"] = var.detach()
        else:
            y = conv3x3_reflect(h, w, None)
            y = _q(y, emulate_bf16)
            if rec is not None:
                rec.put(f"{prefix}{ci}.raw", y)
            # eval: BN(y + b) with running stats == affine on the bias-free output
            out = batchnorm_eval(y + b[None, :, None, None], gamma, beta,
                                 sd[f"{prefix}{bi}.running_mean"], sd[f"{prefix}{bi}.running_var"])
        h = F.relu(out)
        if ci == 3 and drop_mask is not None:
            h = h * drop_mask[:, :, None, None]
        h = _q(h, emulate_bf16)
        if rec is not None:
            rec.put(f"{prefix}{ci}.act", h)
    return h


# ----------------------------------------------------------------------------
# whole model (model.py)
# ----------------------------------------------------------------------------
def mimo_unet_forward(
    x: Tensor,
    sd: Dict[str, Tensor],
    num_subnetworks: int,
    training: bool = True,
    emulate_bf16: bool = False,
    new_stats: Optional[Dict[str, Tensor]] = None,
    drop_masks: Optional[Dict[str, Tensor]] = None,
    rec: Optional[Recorder] = None,
    elem_masks: Optional[Dict[str, Tensor]] = None,
) -> Tensor:
    """model.py:94-117 (MimoUNet.forward), bilinear=True / use_pooling_indices=False path
    (the only one the reference scripts run, mimo_unet.py:73-74).

    x: [B, S, Cin, H, W] -> [B, S, Cout, H, W].
    drop_masks: optional {double_conv prefix: [N, C] scaled keep mask} for Dropout2d.
    elem_masks: optional scaled keep masks of the element-wise nn.Dropout layers: "center" [B, C5, H/16, W/16]
    (model.py:239, applied to x5) and "final.<i>" [B, f, H, W] (model.py:294, in front of head i).
    """
    S = num_subnetworks
    assert x.shape[1] == S
    dm = drop_masks or {}
    x = _q(x, emulate_bf16)

    def dc(h, prefix):
        return double_conv(h, sd, prefix, training, emulate_bf16, new_stats, dm.get(prefix), rec)

    # encoder  (model.py:150-175)
    x1s, x2s = [], []
    for i in range(S):
        x1 = dc(x[:, i], f"encoder.in_convs.{i}.double_conv.")
        x2 = dc(maxpool2x2(x1), f"encoder.down1s.{i}.conv.double_conv.")
        x1s.append(x1)
        x2s.append(x2)
    xc = torch.cat(x2s, dim=1)  # model.py:113

    # core  (model.py:232-243)
    x3 = dc(maxpool2x2(xc), "core.down2.conv.double_conv.")
    x4 = dc(maxpool2x2(x3), "core.down3.conv.double_conv.")
    x5 = dc(maxpool2x2(x4), "core.down4.conv.double_conv.")
    em = elem_masks or {}
    if "center" in em:  # model.py:239
        x5 = _q(x5 * em["center"], emulate_bf16)

    def up(x_low, skip, prefix):  # components.py:106-120
        u = _q(upsample_bilinear2x_ac(x_low), emulate_bf16)
        u = pad_to(u, skip.shape[2], skip.shape[3])
        return dc(torch.cat([skip, u], dim=1), prefix)

    u = up(x5, x4, "core.up1.conv.double_conv.")
    u = up(u, x3, "core.up2.conv.double_conv.")
    u = up(u, xc, "core.up3.conv.double_conv.")

    # decoder  (model.py:285-297)
    outs = []
    for i in range(S):
        d = up(u, x1s[i], f"decoder.up4s.{i}.conv.double_conv.")
        if f"final.{i}" in em:  # model.py:294
            d = _q(d * em[f"final.{i}"], emulate_bf16)
        if rec is not None:
            rec.put(f"decoder.feat.{i}", d)
        o = F.conv2d(d, sd[f"decoder.outcs.{i}.conv.weight"], sd[f"decoder.outcs.{i}.conv.bias"])
        outs.append(o)
    return torch.stack(outs, dim=1)


# ----------------------------------------------------------------------------
# loss, loss buffer, aggregation
# ----------------------------------------------------------------------------
def laplace_nll_elementwise(mu: Tensor, log_s: Tensor, y: Tensor, mask: Optional[Tensor] = None,
                            eps_min: float = 1e-5, eps_max: float = 1e3) -> Tensor:
    """losses.py:132-164.  l = log(clamp(exp(log_s))) + |mu - y| / clamp(exp(log_s))."""
    s_raw = torch.exp(log_s)
    s_c = s_raw + (s_raw.detach().clamp(eps_min, eps_max) - s_raw.detach())  # clamp under no_grad
    loss = torch.log(s_c) + (mu - y).abs() / s_c
    if mask is not None:
        loss = loss * mask
    return loss


def laplace_nll_grads(mu: Tensor, log_s: Tensor, y: Tensor, mask: Optional[Tensor] = None,
                      eps_min: float = 1e-5, eps_max: float = 1e3):
    """Closed-form gradients of the elementwise loss."""
    d = mu - y
    s_raw = torch.exp(log_s)
    s_c = s_raw.clamp(eps_min, eps_max)
    g_mu = torch.sign(d) / s_c
    g_ls = (1.0 / s_c - d.abs() / (s_c * s_c)) * s_raw
    if mask is not None:
        g_mu, g_ls = g_mu * mask, g_ls * mask
    return g_mu, g_ls


def gaussian_nll_elementwise(mu: Tensor, log_var: Tensor, y: Tensor, mask: Optional[Tensor] = None,
                             eps_min: float = 1e-5, eps_max: float = 1e3) -> Tensor:
    """losses.py:48-79 (GaussianNLL.forward).  l = log(clamp(exp(log_var))) + (mu - y)^2 / clamp(exp(log_var)); the clamp is
    applied in place under no_grad on a clone, i.e. it is invisible to autograd."""
    v_raw = torch.exp(log_var)
    v_c = v_raw + (v_raw.detach().clamp(eps_min, eps_max) - v_raw.detach())
    loss = torch.log(v_c) + (mu - y) ** 2 / v_c
    if mask is not None:
        loss = loss * mask
    return loss


def evidential_head(raw: Tensor) -> Tensor:
    """evidential_unet.py:85-96: raw [B,4,H,W] = (mu, log v, log alpha, log beta) -> (mu, softplus, softplus + 1, softplus)."""
    mu, logv, loga, logb = torch.unbind(raw, dim=1)
    return torch.stack([mu, F.softplus(logv), F.softplus(loga) + 1, F.softplus(logb)], dim=1)


def evidential_loss_elementwise(params: Tensor, y: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """losses.py:203-256 (EvidentialLoss.evidential_loss + forward): params [B,4,H,W] = (gamma, v, alpha, beta), y [B,1,H,W]
    -> [B,H,W].  Gamma(x) = exp(lgamma(x)) as in the reference."""
    mu, v, alpha, beta = torch.unbind(params, dim=1)
    t = y.squeeze(dim=1)
    coeff = torch.exp(torch.lgamma(alpha - 0.5)) / (4 * torch.exp(torch.lgamma(alpha)) * v * torch.sqrt(beta))
    second = 2 * beta * (1 + v) + (2 * alpha - 1) * v * (t - mu) ** 2
    loss = coeff * second + (t - mu) ** 2 * (2 * alpha + v)
    if mask is not None:
        loss = loss * mask
    return loss


def laplace_std(log_s: Tensor) -> Tensor:
    """losses.py:166-167."""
    return torch.exp(log_s) * math.sqrt(2.0)


def laplace_dist_param(std: Tensor, log: bool = False, eps_min: float = 1e-5, eps_max: float = 1e3) -> Tensor:
    """losses.py:172-192."""
    p = (std / math.sqrt(2.0)).clamp(eps_min, eps_max)
    return torch.log(p) if log else p


class LossBufferOracle:
    """loss_buffer.py:18-74 as plain python/torch on CPU."""

    def __init__(self, subnetworks: int, temperature: float, buffer_size: int):
        assert temperature > 0
        self.S, self.T, self.size = subnetworks, temperature, buffer_size
        self.buf = torch.zeros(buffer_size, subnetworks)
        self.index = 0

    def get_weights(self) -> Tensor:
        mean = self.buf.mean(dim=0) if self.size else torch.zeros(self.S)
        return torch.softmax(mean / self.T, dim=-1) * self.S

    def add(self, loss: Tensor):
        if self.size:
            self.buf[self.index] = loss.detach().cpu().float()
            self.index = (self.index + 1) % self.size


def train_loss(out: Tensor, y: Tensor, mask: Optional[Tensor], weights: Tensor):
    """mimo_unet.py:223-247 + :138.  out [B,S,2C,H,W]; returns (loss[S], scalar weighted loss)."""
    C = out.shape[2] // 2
    p1, p2 = out[:, :, :C], out[:, :, C:]
    l = laplace_nll_elementwise(p1, p2, y, mask)
    loss = l.mean(dim=(0, 2, 3, 4))
    return loss, (loss * weights.to(loss)).mean()


def compute_uncertainties(p1: Tensor, p2: Tensor):
    """models/utils.py:76-101: mean, aleatoric variance (mean 2 b^2), epistemic variance (unbiased)."""
    S = p1.shape[1]
    mean = p1.mean(dim=1)
    alea = torch.square(laplace_std(p2)).mean(dim=1)
    if S > 1:
        epi = torch.square(p1 - p1.mean(dim=1, keepdim=True)).sum(dim=1) / (S - 1)
    else:
        epi = torch.zeros_like(alea)
    return mean, alea, epi


def input_shuffle_indices(B: int, S: int, p_rep: float, batch_repetitions: int,
                          generator: Optional[torch.Generator] = None) -> List[Tensor]:
    """models/utils.py:27-36 index construction (CPU generator)."""
    main = torch.randperm(B, generator=generator).repeat(batch_repetitions)
    k = int(main.shape[0] * (1.0 - p_rep))
    return [torch.cat((main[:k][torch.randperm(k, generator=generator)], main[k:]), dim=0) for _ in range(S)]


def validation_math(out: Tensor, label: Tensor, mask: Optional[Tensor] = None):
    """mimo_unet.py:157-169 given out [B,S,2C,H,W] and label [B,S,C,H,W] (already repeated)."""
    C = out.shape[2] // 2
    p1, p2 = out[:, :, :C], out[:, :, C:]
    m5 = None if mask is None else mask[:, None].expand(-1, p1.shape[1], -1, -1, -1)
    val_loss = laplace_nll_elementwise(p1, p2, label, m5).mean(dim=(0, 2, 3, 4))
    mean, alea, epi = compute_uncertainties(p1, p2)
    y_mean = label.mean(dim=1)
    comb = laplace_dist_param(torch.sqrt(alea + epi), log=True)
    combined = laplace_nll_elementwise(mean, comb, y_mean, mask).mean()
    return val_loss, combined, mean, alea, epi


# ----------------------------------------------------------------------------
# algorithmic work -- used by bench.py for roofline numerators
# ----------------------------------------------------------------------------
def conv_layer_table(in_channels: int, S: int, f: int, H: int, W: int, out_channels: int = 2):
    """List of (name, instances, Cin, Cout, H, W, k) for every conv, bilinear path."""
    L = []
    L.append(("encoder.in_convs.0", S, in_channels, f, H, W, 3))
    L.append(("encoder.in_convs.3", S, f, f, H, W, 3))
    H2, W2 = H // 2, W // 2
    L.append(("encoder.down1s.0", S, f, 2 * f, H2, W2, 3))
    L.append(("encoder.down1s.3", S, 2 * f, 2 * f, H2, W2, 3))
    H4, W4, H8, W8, H16, W16 = H2 // 2, W2 // 2, H2 // 4, W2 // 4, H2 // 8, W2 // 8
    c = 2 * f * S
    L.append(("core.down2.0", 1, c, 2 * c, H4, W4, 3)); L.append(("core.down2.3", 1, 2 * c, 2 * c, H4, W4, 3))
    L.append(("core.down3.0", 1, 2 * c, 4 * c, H8, W8, 3)); L.append(("core.down3.3", 1, 4 * c, 4 * c, H8, W8, 3))
    L.append(("core.down4.0", 1, 4 * c, 4 * c, H16, W16, 3)); L.append(("core.down4.3", 1, 4 * c, 4 * c, H16, W16, 3))
    L.append(("core.up1.0", 1, 8 * c, 4 * c, H8, W8, 3)); L.append(("core.up1.3", 1, 4 * c, 2 * c, H8, W8, 3))
    L.append(("core.up2.0", 1, 4 * c, 2 * c, H4, W4, 3)); L.append(("core.up2.3", 1, 2 * c, c, H4, W4, 3))
    L.append(("core.up3.0", 1, 2 * c, c, H2, W2, 3)); L.append(("core.up3.3", 1, c, c // 2, H2, W2, 3))
    d = c // 2 + f
    L.append(("decoder.up4s.0", S, d, d // 2, H, W, 3)); L.append(("decoder.up4s.3", S, d // 2, f, H, W, 3))
    L.append(("decoder.outcs", S, f, out_channels, H, W, 1))
    return L


def flops_per_sample(in_channels: int, S: int, f: int, H: int, W: int, out_channels: int = 2):
    """(fwd, train) algorithmic FLOPs per sample with true channel counts."""
    fwd = train = 0.0
    for name, inst, ci, co, h, w, k in conv_layer_table(in_channels, S, f, H, W, out_channels):
        fl = 2.0 * h * w * co * ci * k * k * inst
        fwd += fl
        train += 3 * fl
        if name == "encoder.in_convs.0":
            train -= fl  # no dgrad into the image during training
    return fwd, train


# ----------------------------------------------------------------------------
# deterministic parameters (shared by the golden generator and the tests)
# ----------------------------------------------------------------------------
def state_dict_spec(in_channels: int, out_channels: int, S: int, f: int):
    """Ordered (name, shape, kind) list of MimoUNet.state_dict(), bilinear path."""
    spec = []

    def dconv(prefix, cin, cmid, cout):
        for idx, (ci, co) in zip((0, 3), ((cin, cmid), (cmid, cout))):
            spec.append((f"{prefix}{idx}.weight", (co, ci, 3, 3), "conv_w"))
            spec.append((f"{prefix}{idx}.bias", (co,), "conv_b"))
            b = idx + 1
            spec.append((f"{prefix}{b}.weight", (co,), "bn_w"))
            spec.append((f"{prefix}{b}.bias", (co,), "bn_b"))
            spec.append((f"{prefix}{b}.running_mean", (co,), "bn_rm"))
            spec.append((f"{prefix}{b}.running_var", (co,), "bn_rv"))
            spec.append((f"{prefix}{b}.num_batches_tracked", (), "bn_n"))

    for i in range(S):
        dconv(f"encoder.in_convs.{i}.double_conv.", in_channels, f, f)
    for i in range(S):
        dconv(f"encoder.down1s.{i}.conv.double_conv.", f, 2 * f, 2 * f)
    c = 2 * f * S
    dconv("core.down2.conv.double_conv.", c, 2 * c, 2 * c)
    dconv("core.down3.conv.double_conv.", 2 * c, 4 * c, 4 * c)
    dconv("core.down4.conv.double_conv.", 4 * c, 4 * c, 4 * c)
    dconv("core.up1.conv.double_conv.", 8 * c, 4 * c, 2 * c)
    dconv("core.up2.conv.double_conv.", 4 * c, 2 * c, c)
    dconv("core.up3.conv.double_conv.", 2 * c, c, c // 2)
    d = c // 2 + f
    for i in range(S):
        dconv(f"decoder.up4s.{i}.conv.double_conv.", d, d // 2, f)
    for i in range(S):
        spec.append((f"decoder.outcs.{i}.conv.weight", (out_channels, f, 1, 1), "conv_w"))
        spec.append((f"decoder.outcs.{i}.conv.bias", (out_channels,), "conv_b"))
    return spec


def make_state_dict(in_channels: int, out_channels: int, S: int, f: int, seed: int) -> Dict[str, Tensor]:
    """Deterministic, non-trivial parameters (incl. BN affine and running stats) from a CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape, kind in state_dict_spec(in_channels, out_channels, S, f):
        if kind == "conv_w":
            bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "conv_b":
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "bn_w":
            v = torch.rand(shape, generator=g) + 0.5
        elif kind in ("bn_b", "bn_rm"):
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        elif kind == "bn_rv":
            v = torch.rand(shape, generator=g) + 0.5
        else:
            v = torch.tensor(3, dtype=torch.int64)
        sd[name] = v
    return sd


def grad_digest(g: Tensor) -> Dict[str, Tensor]:
    """Compact fingerprint of a gradient tensor stored in fixtures instead of the full tensor."""
    flat = g.detach().reshape(-1).float()
    return {"norm": flat.norm().reshape(1), "head": flat[:64].clone(), "stride97": flat[::97].clone()}
