"""Uncertainty losses with the reference's interface (reference: mimo/losses.py).

``LaplaceNLL`` -- the loss named by the north star -- runs as fused CUDA kernels (forward and autograd backward)
from libmimo_b200.so; ``GaussianNLL`` (``UncertaintyLoss.from_name("gaussian_nll")``) shares those kernels with its own
element math, and ``EvidentialLoss`` (the deep-evidential-regression baseline) has its own elementwise kernels. On CPU
tensors the stock torch formulas are used (the losses are plain elementwise math; the network itself has no CPU path).
"""
import math
from abc import ABC, abstractmethod

import torch


class UncertaintyLoss(torch.nn.Module, ABC):
    @abstractmethod
    def forward(self, y_hat, log_variance, y, mask) -> torch.Tensor: ...

    @abstractmethod
    def std(self, mu, log_variance) -> torch.Tensor: ...

    @abstractmethod
    def mode(self, mu, log_variance) -> torch.Tensor: ...

    @abstractmethod
    def calculate_dist_param(self, std: torch.Tensor, *, log: bool = False) -> torch.Tensor: ...

    @property
    @abstractmethod
    def num_distribution_params(self) -> int: ...

    @classmethod
    def from_name(cls, name: str) -> "UncertaintyLoss":
        table = {"gaussian_nll": GaussianNLL, "laplace_nll": LaplaceNLL}
        if name not in table:
            raise ValueError(f"Unknown loss function: {name}")
        return table[name]()


def _clamped_no_grad(t: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """Value-clamp with the identity gradient of the reference's `.clone()` + in-place clamp under no_grad."""
    return t + (t.detach().clamp(min=lo, max=hi) - t.detach())


class LaplaceNLL(UncertaintyLoss):
    """log(b) + |y_hat - y| / b with b = clamp(exp(log_scale), eps_min, eps_max) (clamp invisible to autograd)."""
    num_distribution_params = 2

    def __init__(self, eps_min: float = 1e-5, eps_max: float = 1e3):
        super().__init__()
        self.eps_min, self.eps_max = eps_min, eps_max

    def forward(self, y_hat: torch.Tensor, log_scale: torch.Tensor, y: torch.Tensor, *, mask: torch.Tensor = None,
                reduce_mean: bool = True):
        from mimo_unet_b200 import functional as Fn
        return Fn.laplace_nll(y_hat, log_scale, y, mask=mask, reduce_mean=reduce_mean, eps_min=self.eps_min, eps_max=self.eps_max)

    def std(self, mu, log_scale):
        return torch.exp(log_scale) * math.sqrt(2.0)

    def mode(self, mu, log_scale):
        return mu

    def calculate_dist_param(self, std: torch.Tensor, *, log: bool = False) -> torch.Tensor:
        param = _clamped_no_grad(std / math.sqrt(2.0), self.eps_min, self.eps_max)
        return torch.log(param) if log else param


class GaussianNLL(UncertaintyLoss):
    """log(var) + (y_hat - y)^2 / var, var = clamp(exp(log_variance)) (clamp invisible to autograd)."""
    num_distribution_params = 2

    def __init__(self, eps_min: float = 1e-5, eps_max: float = 1e3):
        super().__init__()
        self.eps_min, self.eps_max = eps_min, eps_max

    def forward(self, y_hat, log_variance, y, *, mask=None, reduce_mean: bool = True):
        if y_hat.is_cuda:
            from mimo_unet_b200 import functional as Fn
            return Fn.gaussian_nll(y_hat, log_variance, y, mask=mask, reduce_mean=reduce_mean, eps_min=self.eps_min, eps_max=self.eps_max)
        var = _clamped_no_grad(torch.exp(log_variance), self.eps_min, self.eps_max)
        loss = torch.log(var) + (y_hat - y) ** 2 / var
        if mask is not None:
            loss = loss * mask
        return loss.mean() if reduce_mean else loss

    def std(self, mu, log_variance):
        return torch.exp(log_variance) ** 0.5

    def mode(self, mu, log_variance):
        return mu

    def calculate_dist_param(self, std, *, log: bool = False):
        param = _clamped_no_grad(std ** 2, self.eps_min, self.eps_max)
        return torch.log(param) if log else param


class EvidentialLoss(torch.nn.Module):
    """Deep evidential regression loss (reference mimo/losses.py:195-271): sum-of-squares NIG loss + regulariser on
    evidential_output [B, 4, H, W] = (gamma, v, alpha, beta). One fused elementwise CUDA pass each way on the GPU."""
    num_distribution_params = 4

    def __init__(self, coeff: float) -> None:
        super().__init__()
        self.coeff = coeff

    @staticmethod
    def evidential_loss(mu, v, alpha, beta, targets):
        def gamma_fn(x):
            return torch.exp(torch.lgamma(x))
        coeff = gamma_fn(alpha - 0.5) / (4 * gamma_fn(alpha) * v * torch.sqrt(beta))
        second = 2 * beta * (1 + v) + (2 * alpha - 1) * v * torch.pow(targets - mu, 2)
        return coeff * second + torch.pow(targets - mu, 2) * (2 * alpha + v)

    def forward(self, evidential_output, y_true, *, mask=None, reduce_mean=False) -> torch.Tensor:
        if evidential_output.is_cuda:
            from mimo_unet_b200 import functional as Fn
            m = mask
            if m is not None and m.dim() == 4:
                m = m.squeeze(1)
            return Fn.evidential_loss(evidential_output, y_true.squeeze(dim=1), mask=m, reduce_mean=reduce_mean)
        gamma, v, alpha, beta = torch.unbind(evidential_output, dim=1)
        loss = self.evidential_loss(gamma, v, alpha, beta, y_true.squeeze(dim=1))
        if mask is not None:
            loss = loss * mask
        return torch.mean(loss) if reduce_mean else loss

    @staticmethod
    def mode(evidential_output):
        return torch.unbind(evidential_output, dim=1)[0]

    @staticmethod
    def aleatoric_var(evidential_output):
        gamma, v, alpha, beta = torch.unbind(evidential_output, dim=1)
        return beta / (alpha - 1)

    @staticmethod
    def epistemic_var(evidential_output):
        gamma, v, alpha, beta = torch.unbind(evidential_output, dim=1)
        return beta / (v * (alpha - 1))
