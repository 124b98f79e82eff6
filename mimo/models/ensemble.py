"""EnsembleModule with the reference's interface (reference: mimo/models/ensemble.py): several MimoUnetModel
checkpoints and/or Monte-Carlo dropout passes, aggregated into mean / aleatoric / epistemic variance.

Differences in mechanics only: predictions stay on the GPU and are aggregated by one fused kernel (the reference
copies every member to the host and aggregates there); the repeated input is not materialised per member."""
from typing import List

import torch

from ._lightning_compat import LightningModule
from .mimo_unet import MimoUnetModel
from .utils import compute_uncertainties, repeat_subnetworks


class EnsembleModule(LightningModule):
    def __init__(self, checkpoint_paths: List[str], monte_carlo_steps: int = 0, return_raw_predictions=False, models=None):
        super().__init__()
        # `models` (already constructed MimoUnetModel instances) is an addition for checkpoint-free use
        self.models = list(models) if models is not None else [MimoUnetModel.load_from_checkpoint(p) for p in checkpoint_paths]
        self.monte_carlo_steps = monte_carlo_steps
        self.return_raw_predictions = return_raw_predictions
        for m in self.models:
            m.eval()
            if monte_carlo_steps > 0:
                self._activate_mc_dropout(m)

    @staticmethod
    def _activate_mc_dropout(model: torch.nn.Module):
        """Puts every Dropout* submodule back into training mode (BatchNorm stays in eval)."""
        for sub in model.modules():
            if type(sub).__name__.startswith("Dropout"):
                sub.train()

    @property
    def num_subnetworks(self):
        return sum(m.num_subnetworks for m in self.models)

    @property
    def loss_fn(self):
        return self.models[0].loss_fn

    def to(self, *a, **k):
        for m in self.models:
            m.to(*a, **k)
        return super().to(*a, **k)

    @property
    def device(self):
        return next(self.models[0].parameters()).device

    def forward(self, x: torch.Tensor):
        """x [B,C_in,H,W] -> (mean, aleatoric_variance, epistemic_variance) each [B,C_out,H,W], or raw (p1, p2)
        with all members stacked on dim 1 when `return_raw_predictions`."""
        p1s, p2s = [], []
        for m in self.models:
            m.to(x.device)
            x_rep = repeat_subnetworks(x, num_subnetworks=m.num_subnetworks)
            for _ in range(max(1, self.monte_carlo_steps)):
                p1, p2 = m(x_rep)
                p1s.append(p1)
                p2s.append(p2)
        p1, p2 = (p1s[0], p2s[0]) if len(p1s) == 1 else (torch.cat(p1s, dim=1), torch.cat(p2s, dim=1))
        if self.return_raw_predictions:
            return p1, p2
        return compute_uncertainties(self.loss_fn, y_preds=p1, log_params=p2)
