"""EvidentialUnetModel with the reference's interface (reference: mimo/models/evidential_unet.py): the deep-evidential-
regression baseline = the B200-native MimoUNet with one subnetwork and four output channels, a softplus head and
EvidentialLoss. Constructor arguments, step-output dictionaries, logged keys, optimizer setup and argparse flags follow
the reference; the head and the loss are fused elementwise CUDA kernels (mimo_evidential_head / mimo_evidential_loss_*)."""
from argparse import ArgumentParser
from typing import Any, Dict, Literal

import torch

from mimo.losses import EvidentialLoss
from mimo.metrics import compute_regression_metrics
from mimo.utils import count_trainable_parameters
from ._lightning_compat import LightningModule
from .mimo_components.model import MimoUNet


class EvidentialUnetModel(LightningModule):
    def __init__(self, in_channels: int, out_channels: int, filter_base_count: int, center_dropout_rate: float,
                 final_dropout_rate: float, encoder_dropout_rate: float, core_dropout_rate: float, decoder_dropout_rate: float,
                 weight_decay: float, learning_rate: float, seed: int, scheduler_step_size: int = 20, scheduler_gamma: float = 0.5):
        super().__init__()
        for k, v in dict(in_channels=in_channels, out_channels=out_channels, filter_base_count=filter_base_count,
                         center_dropout_rate=center_dropout_rate, final_dropout_rate=final_dropout_rate,
                         encoder_dropout_rate=encoder_dropout_rate, core_dropout_rate=core_dropout_rate,
                         decoder_dropout_rate=decoder_dropout_rate, weight_decay=weight_decay, learning_rate=learning_rate, seed=seed,
                         scheduler_step_size=scheduler_step_size, scheduler_gamma=scheduler_gamma).items():
            setattr(self, k, v)
        self.loss_fn = EvidentialLoss(coeff=1.0)
        self.model = MimoUNet(in_channels=in_channels, out_channels=out_channels, num_subnetworks=1, filter_base_count=filter_base_count,
                              center_dropout_rate=center_dropout_rate, final_dropout_rate=final_dropout_rate,
                              encoder_dropout_rate=encoder_dropout_rate, core_dropout_rate=core_dropout_rate,
                              decoder_dropout_rate=decoder_dropout_rate, bilinear=True, use_pooling_indices=False)
        self.save_hyperparameters()
        self.save_hyperparameters({"loss": "evidential", "trainable_params": count_trainable_parameters(self.model)})

    def compile(self):
        """No-op (see MimoUnetModel.compile): the network already is a fixed, pre-planned kernel sequence."""
        return None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        fixed = {k.replace("model._orig_mod.", "model."): v for k, v in state_dict.items()}
        return super().load_state_dict(fixed, strict=strict, **kw)

    def forward(self, x: torch.Tensor):
        """x [B, C_in, H, W] -> [B, 4, H, W] = (mu, v, alpha, beta)."""
        B, C_in, H, W = x.shape
        assert C_in == self.in_channels, "channel dimension must match in_channels"
        out = torch.squeeze(self.model(torch.unsqueeze(x, dim=1)), dim=1)
        if out.is_cuda:
            from mimo_unet_b200 import functional as Fn
            return Fn.evidential_head(out)
        mu, logv, logalpha, logbeta = torch.unbind(out, dim=1)
        sp = torch.nn.functional.softplus
        return torch.stack([mu, sp(logv), sp(logalpha) + 1, sp(logbeta)], dim=1)

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int) -> Dict[str, torch.Tensor]:
        image, label = batch["image"], batch["label"]
        mask = batch.get("mask")
        out = self(image)
        loss = self.loss_fn(out, label, mask=mask)
        y_pred = self.loss_fn.mode(out).unsqueeze(dim=1)
        aleatoric_std = self.loss_fn.aleatoric_var(out).unsqueeze(dim=1) ** 0.5
        self._log_metrics(y_pred=y_pred, y_true=label, stage="train")
        return {"loss": loss.mean(), "label": label, "preds": y_pred, "aleatoric_std_map": aleatoric_std,
                "err_map": y_pred - label, "mask": mask}

    def validation_step(self, batch: Dict[str, torch.Tensor], batch_idx: int) -> Dict[str, torch.Tensor]:
        image, label = batch["image"], batch["label"]
        mask = batch.get("mask")
        out = self(image)
        loss = self.loss_fn.forward(out, label, mask=mask, reduce_mean=False)
        y_pred = self.loss_fn.mode(out).unsqueeze(dim=1)
        aleatoric_std = self.loss_fn.aleatoric_var(out).unsqueeze(dim=1) ** 0.5
        epistemic_std = self.loss_fn.epistemic_var(out).unsqueeze(dim=1) ** 0.5
        self.log("val_loss", loss.mean(), batch_size=self._batch_size())
        self._log_metrics(y_pred=y_pred, y_true=label, stage="val")
        self._log_uncertainties(aleatoric_std, epistemic_std)
        return {"loss": loss.mean(), "label": label, "preds": y_pred, "aleatoric_std_map": aleatoric_std,
                "epistemic_std_map": epistemic_std, "err_map": y_pred - label, "mask": mask}

    def configure_optimizers(self) -> Dict[str, Any]:
        params = list(self.parameters())
        fused = bool(params) and all(p.is_cuda for p in params)
        optimizer = torch.optim.Adam(params, lr=self.learning_rate, weight_decay=self.weight_decay, fused=fused)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=self.scheduler_step_size, gamma=self.scheduler_gamma)
        return dict(optimizer=optimizer, lr_scheduler=scheduler, monitor="val_loss")

    def _batch_size(self):
        try:
            return self.trainer.datamodule.batch_size
        except Exception:
            return None

    def _log_metrics(self, y_pred: torch.Tensor, y_true: torch.Tensor, stage: Literal["train", "val"] = "train") -> None:
        for name, value in compute_regression_metrics(y_pred.flatten(), y_true.flatten()).items():
            self.log(f"metric_{stage}/{name}", value, on_step=(stage == "train"), on_epoch=True, metric_attribute=name,
                     batch_size=self._batch_size())

    def _log_uncertainties(self, aleatoric_std: torch.Tensor, epistemic_std: torch.Tensor) -> None:
        bs = self._batch_size()
        self.log("metric_val/aleatoric_std_mean", aleatoric_std.clip(0, 5).mean(), batch_size=bs)
        self.log("metric_val/epistemic_std_mean", epistemic_std.clip(0, 5).mean(), batch_size=bs)

    @staticmethod
    def add_model_specific_args(parent_parser: ArgumentParser) -> ArgumentParser:
        p = parent_parser.add_argument_group(title="MIMO UNet Model")
        for name, typ, default in (("filter_base_count", int, 32), ("center_dropout_rate", float, 0.0), ("final_dropout_rate", float, 0.0),
                                   ("encoder_dropout_rate", float, 0.0), ("core_dropout_rate", float, 0.0),
                                   ("decoder_dropout_rate", float, 0.0), ("learning_rate", float, 1e-3), ("weight_decay", float, 0.0),
                                   ("scheduler_step_size", int, 20), ("scheduler_gamma", float, 0.5)):
            p.add_argument(f"--{name}", type=typ, default=default)
        return parent_parser
