"""Regression metrics with the reference's interface (reference: mimo/metrics.py). Uses torchmetrics when it
is installed; otherwise the same quantities from torch ops (logging only, outside the hot path)."""
from functools import partial
from typing import Dict, List, Optional

import torch

try:  # pragma: no cover - depends on the environment
    import torchmetrics.functional as _tmf
except Exception:  # torchmetrics is not installed in the build container
    _tmf = None


def _mse(y_hat, y, squared=True):
    v = torch.mean((y_hat - y) ** 2)
    return v if squared else torch.sqrt(v)


def _r2(y_hat, y):
    ss_res = torch.sum((y - y_hat) ** 2)
    ss_tot = torch.sum((y - y.mean()) ** 2)
    return 1 - ss_res / ss_tot


def get_metric(metric: str):
    if _tmf is not None:
        table = {"mae": _tmf.mean_absolute_error, "mse": _tmf.mean_squared_error,
                 "rmse": partial(_tmf.mean_squared_error, squared=False), "r2": _tmf.r2_score,
                 "mape": _tmf.mean_absolute_percentage_error}
    else:
        table = {"mae": lambda a, b: torch.mean(torch.abs(a - b)), "mse": _mse, "rmse": partial(_mse, squared=False),
                 "r2": _r2, "mape": lambda a, b: torch.mean(torch.abs(a - b) / torch.clamp(torch.abs(b), min=1.17e-06))}
    if metric not in table:
        raise ValueError(f"Unknown metric: {metric}")
    return table[metric]


def compute_regression_metrics(y_hat: torch.Tensor, y: torch.Tensor,
                               metrics: Optional[List[str]] = ("r2", "mae", "mse", "rmse")) -> Dict[str, float]:
    y, y_hat = y.detach(), y_hat.detach()
    return {m: get_metric(m)(y_hat, y) for m in metrics}
