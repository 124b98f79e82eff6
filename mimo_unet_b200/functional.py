"""Tensor-level wrappers (with autograd) of the loss / loss-buffer / aggregation kernels.

All functions require CUDA tensors; CPU tensors raise (there is no CPU fallback in the product).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check
from .engine import stream_ptr, _ptr


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MimoError("mimo_unet_b200 kernels need CUDA tensors (no CPU fallback); got a CPU tensor")


def _factorisations(t: torch.Tensor):
    """All ways to see `t` as [rows][cols] with contiguous cols and ONE row stride: {k: (rows, cols, row_stride)}."""
    nd = t.dim()
    sizes, strides = list(t.shape), list(t.stride())
    res = {}
    for k in range(0, nd + 1):
        exp, ok = 1, True
        for d in range(nd - 1, k - 1, -1):
            if sizes[d] != 1 and strides[d] != exp:
                ok = False
                break
            exp *= sizes[d]
        if not ok:
            continue
        cols, rows, rs = exp, 1, None
        for d in range(k - 1, -1, -1):
            if sizes[d] == 1:
                continue
            if rs is None:
                rs, rows = strides[d], sizes[d]
            elif strides[d] != rs * rows:
                ok = False
                break
            else:
                rows *= sizes[d]
        if ok:
            res[k] = (rows, cols, rs if rs is not None else cols)
    return res


def _common_rows_view(tensors, shape):
    """Broadcasts every tensor to `shape` and finds one [rows][cols] factorisation valid for all of them
    (covers the strided p1/p2 channel-slice views of the network output without copies).
    Returns (tensors, rows, cols, [row_stride per tensor])."""
    ts = []
    for t in tensors:
        t = t.expand(shape) if tuple(t.shape) != tuple(shape) else t
        ts.append(t if t.dtype == torch.float32 else t.float())
    facts = [_factorisations(t) for t in ts]
    common = set(facts[0])
    for f in facts[1:]:
        common &= set(f)
    if common:
        k = min(common)
        rows, cols = facts[0][k][0], facts[0][k][1]
        return ts, rows, cols, [f[k][2] for f in facts]
    ts = [t.contiguous() for t in ts]
    n = ts[0].numel()
    return ts, 1, n, [n] * len(ts)


class _LaplaceNLLFn(torch.autograd.Function):
    """LaplaceNLL.forward (reference mimo/losses.py:132-164) on the GPU, elementwise or mean."""

    @staticmethod
    def forward(ctx, y_hat, log_scale, y, mask, reduce_mean: bool, eps_min: float, eps_max: float, gaussian: bool = False):
        _need_cuda(y_hat, log_scale, y, mask)
        shape = torch.broadcast_shapes(y_hat.shape, log_scale.shape, y.shape, *( [mask.shape] if mask is not None else []))
        ops = [y_hat.detach(), log_scale.detach(), y.detach()] + ([mask.detach()] if mask is not None else [])
        ts, rows, cols, rss = _common_rows_view(ops, shape)
        mu, ls, yy = ts[0], ts[1], ts[2]
        mu_rs, ls_rs, y_rs = rss[0], rss[1], rss[2]
        mm, m_rs = (ts[3], rss[3]) if mask is not None else (None, 0)
        lib = _lib.lib()
        fwd = lib.mimo_gaussian_nll_fwd if gaussian else lib.mimo_laplace_nll_fwd
        dev = mu.device
        n = rows * cols
        if reduce_mean:
            out = torch.empty((), dtype=torch.float32, device=dev)
            part = torch.empty(int(lib.mimo_laplace_scratch_floats()), dtype=torch.float32, device=dev)
            check(fwd(mu.data_ptr(), mu_rs, ls.data_ptr(), ls_rs, yy.data_ptr(), y_rs, _ptr(mm), m_rs, rows, cols,
                      eps_min, eps_max, None, part.data_ptr(), out.data_ptr(), stream_ptr()), "mimo_*_nll_fwd")
        else:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
            check(fwd(mu.data_ptr(), mu_rs, ls.data_ptr(), ls_rs, yy.data_ptr(), y_rs, _ptr(mm), m_rs, rows, cols,
                      eps_min, eps_max, out.data_ptr(), None, None, stream_ptr()), "mimo_*_nll_fwd")
        ctx.save_for_backward(mu, ls, yy, mm if mm is not None else torch.empty(0, device=dev))
        ctx.geom = (rows, cols, mu_rs, ls_rs, y_rs, m_rs, mm is not None, reduce_mean, eps_min, eps_max, shape,
                    tuple(y_hat.shape), tuple(log_scale.shape))
        ctx.gaussian = gaussian
        return out

    @staticmethod
    def backward(ctx, g):
        mu, ls, yy, mm = ctx.saved_tensors
        rows, cols, mu_rs, ls_rs, y_rs, m_rs, has_mask, reduce_mean, eps_min, eps_max, shape, s_mu, s_ls = ctx.geom
        lib = _lib.lib()
        g_mu = torch.empty(shape, dtype=torch.float32, device=mu.device)
        g_ls = torch.empty(shape, dtype=torch.float32, device=mu.device)
        g = g.contiguous().float()
        bwd = lib.mimo_gaussian_nll_bwd if ctx.gaussian else lib.mimo_laplace_nll_bwd
        check(bwd(mu.data_ptr(), mu_rs, ls.data_ptr(), ls_rs, yy.data_ptr(), y_rs, mm.data_ptr() if has_mask else None,
                  m_rs, rows, cols, eps_min, eps_max, g.data_ptr(), int(reduce_mean),
                  1.0 / float(rows * cols) if reduce_mean else 1.0, g_mu.data_ptr(), g_ls.data_ptr(), stream_ptr()),
              "mimo_*_nll_bwd")
        if tuple(shape) != s_mu:
            g_mu = g_mu.sum_to_size(s_mu)
        if tuple(shape) != s_ls:
            g_ls = g_ls.sum_to_size(s_ls)
        return g_mu, g_ls, None, None, None, None, None, None


def laplace_nll(y_hat, log_scale, y, mask=None, reduce_mean=True, eps_min=1e-5, eps_max=1e3):
    return _LaplaceNLLFn.apply(y_hat, log_scale, y, mask, bool(reduce_mean), float(eps_min), float(eps_max), False)


def gaussian_nll(y_hat, log_variance, y, mask=None, reduce_mean=True, eps_min=1e-5, eps_max=1e3):
    """GaussianNLL.forward (reference mimo/losses.py:48-79): same kernels, Gaussian element math."""
    return _LaplaceNLLFn.apply(y_hat, log_variance, y, mask, bool(reduce_mean), float(eps_min), float(eps_max), True)


class DeviceLossBuffer:
    """Device-resident state of the reference LossBuffer (loss_buffer.py:18-74); no host round trips."""

    def __init__(self, subnetworks: int, temperature: float, buffer_size: int, device):
        assert temperature > 0, "Temperature should be positive."
        lib = _lib.lib()
        self.S, self.T, self.size = subnetworks, float(temperature), int(buffer_size)
        nbytes = int(lib.mimo_lossbuffer_bytes(subnetworks, buffer_size))
        self.state = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)
        check(lib.mimo_lossbuffer_init(self.state.data_ptr(), subnetworks, buffer_size, self.T, stream_ptr()), "mimo_lossbuffer_init")

    def weights(self) -> torch.Tensor:
        w = torch.empty(self.S, dtype=torch.float32, device=self.state.device)
        check(_lib.lib().mimo_lossbuffer_get_weights(self.state.data_ptr(), w.data_ptr(), stream_ptr()), "mimo_lossbuffer_get_weights")
        return w

    def add(self, loss: torch.Tensor):
        loss = loss.detach().to(device=self.state.device, dtype=torch.float32).contiguous()
        check(_lib.lib().mimo_lossbuffer_add(self.state.data_ptr(), loss.data_ptr(), stream_ptr()), "mimo_lossbuffer_add")

    @property
    def index(self) -> int:
        return int(self.state[0].item())

    @property
    def buffer(self) -> torch.Tensor:
        return self.state[4: 4 + self.S * max(self.size, 1)].view(torch.float32).view(max(self.size, 1), self.S)[: self.size]


class _TrainLossFn(torch.autograd.Function):
    """_calculate_train_loss (+ loss_weighted.mean()) of the reference (mimo_unet.py:223-247,138) fused with
    the seed of backward: one pass produces loss[S], weights[S], the scalar weighted loss and d/d out."""

    @staticmethod
    def forward(ctx, out, y, mask, gather, lb_state, fixed_w, update_buffer: bool, eps_min: float, eps_max: float,
                want_metrics: bool = False, gaussian: bool = False):
        _need_cuda(out, y, mask)
        lib = _lib.lib()
        B, S, C2, H, W = out.shape
        C = C2 // 2
        HW = H * W
        outc = out.detach().contiguous().float()
        # y: [B,S,C,H,W] (transformed labels) or [B,C,H,W] (+gather / broadcast over S)
        y = y.detach().float()
        if y.dim() == 5:
            yc = y if y.is_contiguous() else (y if (y.stride(1) == 0 and y[:, 0].is_contiguous()) else y.contiguous())
            y_bs, y_ss = yc.stride(0), yc.stride(1)
        else:
            yc = y.contiguous()
            y_bs, y_ss = yc.stride(0), 0
        mc, m_bs, m_ss = None, 0, 0
        if mask is not None:
            m = mask.detach().float()
            if m.dim() == 5:
                mc = m if m.is_contiguous() else (m if (m.stride(1) == 0 and m[:, 0].is_contiguous()) else m.contiguous())
                m_bs, m_ss = mc.stride(0), mc.stride(1)
            else:
                mc = m.contiguous()
                m_bs, m_ss = mc.stride(0), 0
            if mc.shape[-3] != C:  # [B,(S,)1,H,W] mask against C label channels
                mc = mc.expand(*mc.shape[:-3], C, H, W).contiguous()
                m_bs, m_ss = mc.stride(0), (mc.stride(1) if mc.dim() == 5 else 0)
        dev = outc.device
        need_grad = out.requires_grad
        dout = torch.empty_like(outc) if need_grad else None
        res = torch.empty(2 * S + 1 + 4, dtype=torch.float32, device=dev)
        loss, weights, weighted, metrics = res[:S], res[S:2 * S], res[2 * S:2 * S + 1], res[2 * S + 1:]
        if want_metrics or gaussian:
            # r2 / mae / mse / rmse of (mu, y) come out of the same pass (reference mimo/metrics.py: four more reductions)
            part = torch.empty(int(lib.mimo_laplace_train_metrics_scratch_floats(B, S, C, HW)), dtype=torch.float32, device=dev)
            fn = lib.mimo_gaussian_nll_train_metrics if gaussian else lib.mimo_laplace_nll_train_metrics
            check(fn(outc.data_ptr(), yc.data_ptr(), y_bs, y_ss, _ptr(mc), m_bs, m_ss, _ptr(gather), B, S, C,
                     HW, eps_min, eps_max, _ptr(lb_state), _ptr(fixed_w), int(update_buffer), _ptr(dout),
                     part.data_ptr(), loss.data_ptr(), weights.data_ptr(), weighted.data_ptr(),
                     metrics.data_ptr(), stream_ptr()), "mimo_*_nll_train_metrics")
        else:
            part = torch.empty(int(lib.mimo_laplace_train_scratch_floats(B, S, C, HW)), dtype=torch.float32, device=dev)
            check(lib.mimo_laplace_nll_train(outc.data_ptr(), yc.data_ptr(), y_bs, y_ss, _ptr(mc), m_bs, m_ss, _ptr(gather), B, S, C, HW,
                                             eps_min, eps_max, _ptr(lb_state), _ptr(fixed_w), int(update_buffer), _ptr(dout),
                                             part.data_ptr(), loss.data_ptr(), weights.data_ptr(), weighted.data_ptr(), stream_ptr()),
                  "mimo_laplace_nll_train")
        ctx.dout = dout
        ctx.mark_non_differentiable(loss, weights, metrics)
        return weighted.reshape(()), loss, weights, metrics

    @staticmethod
    def backward(ctx, g_weighted, g_loss, g_weights, g_metrics):
        dout = ctx.dout
        ctx.dout = None
        if dout is None:
            return (None,) * 11
        g = g_weighted.detach().reshape(1).float().contiguous()
        check(_lib.lib().mimo_scale_by_scalar(dout.data_ptr(), dout.numel(), g.data_ptr(), stream_ptr()), "mimo_scale_by_scalar")
        return (dout,) + (None,) * 10


def laplace_train_loss(out, y, mask=None, gather=None, loss_buffer: Optional[DeviceLossBuffer] = None, fixed_weights=None,
                       update_buffer=True, eps_min=1e-5, eps_max=1e3, with_metrics=False, gaussian=False):
    """Returns (weighted_mean_loss scalar [differentiable], loss[S], weights[S]); with_metrics appends the regression metrics
    {"mae", "mse", "rmse", "r2"} of (mu, y) as 0-d tensors (reference mimo/metrics.py:22-34), produced by the same kernel pass."""
    total, loss, weights, metrics = _TrainLossFn.apply(out, y, mask, gather, None if loss_buffer is None else loss_buffer.state,
                                                       fixed_weights, bool(update_buffer), float(eps_min), float(eps_max), bool(with_metrics), bool(gaussian))
    if with_metrics:
        return total, loss, weights, {"r2": metrics[3], "mae": metrics[0], "mse": metrics[1], "rmse": metrics[2]}
    return total, loss, weights


class _EvidentialHeadFn(torch.autograd.Function):
    """(mu, log v, log alpha, log beta) -> (mu, softplus, softplus + 1, softplus): reference evidential_unet.py:85-96."""

    @staticmethod
    def forward(ctx, raw):
        _need_cuda(raw)
        r = raw.detach().contiguous().float()
        B, C, H, W = r.shape
        assert C == 4, "the evidential head needs 4 output channels (mu, v, alpha, beta)"
        out = torch.empty_like(r)
        check(_lib.lib().mimo_evidential_head(r.data_ptr(), out.data_ptr(), B, H * W, stream_ptr()), "mimo_evidential_head")
        ctx.save_for_backward(r)
        return out

    @staticmethod
    def backward(ctx, g):
        (r,) = ctx.saved_tensors
        B, _, H, W = r.shape
        g = g.contiguous().float()
        g_raw = torch.empty_like(r)
        check(_lib.lib().mimo_evidential_head_bwd(r.data_ptr(), g.data_ptr(), g_raw.data_ptr(), B, H * W, stream_ptr()), "mimo_evidential_head_bwd")
        return g_raw


def evidential_head(raw: torch.Tensor) -> torch.Tensor:
    return _EvidentialHeadFn.apply(raw)


class _EvidentialLossFn(torch.autograd.Function):
    """EvidentialLoss.forward (reference mimo/losses.py:203-256), elementwise [B,H,W] or mean."""

    @staticmethod
    def forward(ctx, params, y, mask, reduce_mean: bool):
        _need_cuda(params, y, mask)
        p = params.detach().contiguous().float()
        B, C, H, W = p.shape
        assert C == 4
        yy = y.detach().float().reshape(B, H * W).contiguous()
        mm = None if mask is None else mask.detach().float().expand(B, *mask.shape[1:]).reshape(B, H * W).contiguous()
        lib = _lib.lib()
        if reduce_mean:
            out = torch.empty((), dtype=torch.float32, device=p.device)
            part = torch.empty(int(lib.mimo_laplace_scratch_floats()), dtype=torch.float32, device=p.device)
            check(lib.mimo_evidential_loss_fwd(p.data_ptr(), yy.data_ptr(), _ptr(mm), B, H * W, None, part.data_ptr(), out.data_ptr(),
                                               stream_ptr()), "mimo_evidential_loss_fwd")
        else:
            out = torch.empty(B, H, W, dtype=torch.float32, device=p.device)
            check(lib.mimo_evidential_loss_fwd(p.data_ptr(), yy.data_ptr(), _ptr(mm), B, H * W, out.data_ptr(), None, None, stream_ptr()),
                  "mimo_evidential_loss_fwd")
        ctx.save_for_backward(p, yy, mm if mm is not None else torch.empty(0, device=p.device))
        ctx.meta = (mm is not None, reduce_mean)
        return out

    @staticmethod
    def backward(ctx, g):
        p, yy, mm = ctx.saved_tensors
        has_mask, reduce_mean = ctx.meta
        B, _, H, W = p.shape
        g = g.contiguous().float()
        gp = torch.empty_like(p)
        check(_lib.lib().mimo_evidential_loss_bwd(p.data_ptr(), yy.data_ptr(), mm.data_ptr() if has_mask else None, B, H * W, g.data_ptr(),
                                                  int(reduce_mean), 1.0 / float(B * H * W) if reduce_mean else 1.0, gp.data_ptr(),
                                                  stream_ptr()), "mimo_evidential_loss_bwd")
        return gp, None, None, None


def evidential_loss(params, y, mask=None, reduce_mean=False):
    return _EvidentialLossFn.apply(params, y, mask, bool(reduce_mean))


def ensemble_aggregate(p1: torch.Tensor, p2: torch.Tensor):
    """compute_uncertainties (reference mimo/models/utils.py:76-101) for Laplace members on the GPU."""
    _need_cuda(p1, p2)
    B, S = p1.shape[0], p1.shape[1]
    inner_shape = p1.shape[2:]
    inner = 1
    for d in inner_shape:
        inner *= d

    def prep(t):
        t = t.detach().float()
        if not t[0, 0].is_contiguous():
            t = t.contiguous()
        return t, t.stride(0), t.stride(1)

    a, a_bs, a_ss = prep(p1)
    b, b_bs, b_ss = prep(p2)
    mean = torch.empty((B,) + tuple(inner_shape), dtype=torch.float32, device=p1.device)
    alea = torch.empty_like(mean)
    epi = torch.empty_like(mean)
    check(_lib.lib().mimo_ensemble_aggregate(a.data_ptr(), a_bs, a_ss, b.data_ptr(), b_bs, b_ss, B, S, inner, mean.data_ptr(),
                                             alea.data_ptr(), epi.data_ptr(), stream_ptr()), "mimo_ensemble_aggregate")
    return mean, alea, epi


def validation_laplace(p1: torch.Tensor, p2: torch.Tensor, label: torch.Tensor, mask: torch.Tensor = None, eps_min: float = 1e-5,
                       eps_max: float = 1e3):
    """The math of MimoUnetModel.validation_step (reference mimo/models/mimo_unet.py:146-183 with LaplaceNLL) in one pass over
    (p1, p2, label): p1 / p2 [B, S, C, H, W] (location, log-scale), label / mask the UN-repeated [B, C, H, W] tensors.
    Returns a dict: val_loss [S], val_loss_combined, the maps preds / aleatoric_std / epistemic_std / err ([B, C, H, W]),
    metrics {mae, mse, rmse, r2} of (preds, label), aleatoric_std_mean / epistemic_std_mean (means of the stds clipped to [0, 5])."""
    _need_cuda(p1, p2, label)
    B, S = p1.shape[0], p1.shape[1]
    inner_shape = tuple(p1.shape[2:])
    inner = 1
    for d in inner_shape:
        inner *= d

    def prep(t):
        t = t.detach().float()
        if not t[0, 0].is_contiguous():
            t = t.contiguous()
        return t

    a, b = prep(p1), prep(p2)
    if (a.stride(0), a.stride(1)) != (b.stride(0), b.stride(1)):
        a, b = a.contiguous(), b.contiguous()
    y = label.detach().float().contiguous()
    m = mask.detach().float().contiguous() if mask is not None else None
    if tuple(y.shape) != (B,) + inner_shape:
        raise _lib.MimoError(f"validation_laplace: label shape {tuple(y.shape)} != {(B,) + inner_shape}")
    lib = _lib.lib()
    maps = [torch.empty((B,) + inner_shape, dtype=torch.float32, device=p1.device) for _ in range(4)]
    scratch = torch.empty(lib.mimo_validation_scratch_floats(S), dtype=torch.float32, device=p1.device)
    scalars = torch.empty(S + 7, dtype=torch.float32, device=p1.device)
    check(lib.mimo_validation_laplace(a.data_ptr(), b.data_ptr(), a.stride(0), a.stride(1), y.data_ptr(),
                                      m.data_ptr() if m is not None else None, B, S, inner, eps_min, eps_max, maps[0].data_ptr(),
                                      maps[1].data_ptr(), maps[2].data_ptr(), maps[3].data_ptr(), scratch.data_ptr(), scalars.data_ptr(),
                                      stream_ptr()), "mimo_validation_laplace")
    return {"val_loss": scalars[:S], "val_loss_combined": scalars[S], "preds": maps[0], "aleatoric_std": maps[1],
            "epistemic_std": maps[2], "err": maps[3],
            "metrics": {"mae": scalars[S + 1], "mse": scalars[S + 2], "rmse": scalars[S + 3], "r2": scalars[S + 4]},
            "aleatoric_std_mean": scalars[S + 5], "epistemic_std_mean": scalars[S + 6]}
