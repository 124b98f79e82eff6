"""MimoUnetModel: the reference's LightningModule surface (reference: mimo/models/mimo_unet.py) on top of the
B200-native MimoUNet.  Constructor arguments, attributes, step-output dictionaries, logged keys, optimizer setup
and argparse flags follow the reference so scripts/train and scripts/test run unchanged; the math of the steps
goes through fused CUDA kernels (input gather folded into the first convolution, fused Laplace-NLL + loss-buffer
weighting without host synchronisation, fused ensemble aggregation)."""
from argparse import ArgumentParser
from typing import Any, Dict, Literal, Tuple

import torch

from mimo.losses import GaussianNLL, LaplaceNLL, UncertaintyLoss
from mimo.metrics import compute_regression_metrics
from mimo.utils import count_trainable_parameters
from ._lightning_compat import LightningModule
from .mimo_components.loss_buffer import LossBuffer
from .mimo_components.model import MimoUNet
from .utils import compute_uncertainties, flatten_subnetwork_dimension, repeat_subnetworks, shuffle_indices


class MimoUnetModel(LightningModule):
    def __init__(self, in_channels: int, out_channels: int, num_subnetworks: int, filter_base_count: int,
                 center_dropout_rate: float, final_dropout_rate: float, encoder_dropout_rate: float,
                 core_dropout_rate: float, decoder_dropout_rate: float, loss: str, weight_decay: float,
                 learning_rate: float, seed: int, loss_buffer_size: int, loss_buffer_temperature: float,
                 input_repetition_probability: float = 0.0, batch_repetitions: int = 1, scheduler_step_size: int = 20,
                 scheduler_gamma: float = 0.5):
        super().__init__()
        for k, v in dict(in_channels=in_channels, out_channels=out_channels, num_subnetworks=num_subnetworks,
                         filter_base_count=filter_base_count, center_dropout_rate=center_dropout_rate,
                         final_dropout_rate=final_dropout_rate, encoder_dropout_rate=encoder_dropout_rate,
                         core_dropout_rate=core_dropout_rate, decoder_dropout_rate=decoder_dropout_rate,
                         weight_decay=weight_decay, learning_rate=learning_rate, seed=seed,
                         loss_buffer_size=loss_buffer_size, loss_buffer_temperature=loss_buffer_temperature,
                         input_repetition_probability=input_repetition_probability, batch_repetitions=batch_repetitions,
                         scheduler_step_size=scheduler_step_size, scheduler_gamma=scheduler_gamma).items():
            setattr(self, k, v)
        self.loss_fn = UncertaintyLoss.from_name(loss)
        self.model = MimoUNet(in_channels=in_channels, out_channels=out_channels, num_subnetworks=num_subnetworks,
                              filter_base_count=filter_base_count, center_dropout_rate=center_dropout_rate,
                              final_dropout_rate=final_dropout_rate, encoder_dropout_rate=encoder_dropout_rate,
                              core_dropout_rate=core_dropout_rate, decoder_dropout_rate=decoder_dropout_rate,
                              bilinear=True, use_pooling_indices=False)
        self.loss_buffer = LossBuffer(buffer_size=loss_buffer_size, temperature=loss_buffer_temperature,
                                      subnetworks=num_subnetworks)
        self.save_hyperparameters()
        self.save_hyperparameters({"loss": loss, "trainable_params": count_trainable_parameters(self.model)})

    def compile(self):
        """The reference wraps the network in torch.compile here. The B200 network already is a fixed, pre-planned
        kernel sequence, so there is nothing for a tracing compiler to do: deliberately a no-op (state_dict keys
        therefore stay `model.*`; `model._orig_mod.*` checkpoints are accepted on load)."""
        return None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        fixed = {k.replace("model._orig_mod.", "model."): v for k, v in state_dict.items()}
        return super().load_state_dict(fixed, strict=strict, **kw)

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        """x [B,S,C_in,H,W] -> (p1, p2), each [B,S,C_out/2,H,W]: Laplace location and log-scale per subnetwork."""
        B, S, C_in, H, W = x.shape
        assert S == self.num_subnetworks, "subnetwork dimension must match num_subnetworks"
        assert C_in == self.in_channels, "channel dimension must match in_channels"
        out = self.model(x)
        half = self.out_channels // 2
        return out[:, :, :half, ...], out[:, :, half:, ...]

    def _batch_size(self):
        try:
            return self.trainer.datamodule.batch_size
        except Exception:
            return None

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int) -> Dict[str, torch.Tensor]:
        image, label = batch["image"], batch["label"]
        mask = batch.get("mask")
        S, half = self.num_subnetworks, self.out_channels // 2
        # MIMO shuffling: build the per-subnetwork permutations once and let the kernels gather through them
        idx = torch.stack(shuffle_indices(image.shape[0], S, self.input_repetition_probability, self.batch_repetitions,
                                          image.device)).to(image.device)
        out = self.model(image, gather=idx)  # [B*rep, S, C_out, H, W]
        p1, p2 = out[:, :, :half], out[:, :, half:]
        label_t = torch.stack([label.index_select(0, i) for i in idx], dim=1)
        mask_t = None if mask is None else torch.stack([mask.index_select(0, i) for i in idx], dim=1)
        y_pred = self.loss_fn.mode(p1, p2)
        aleatoric_std = self.loss_fn.std(p1, p2)
        self._fused_metrics = None
        loss, loss_weighted, weights = self._calculate_train_loss(p1, p2, y_true=label_t, mask=mask_t, _out=out, _metrics=True)
        self._log_train_loss_and_weights(loss, weights)
        self._log_metrics(y_pred=y_pred, y_true=label_t, stage="train", precomputed=self._fused_metrics)
        return {
            "loss": loss_weighted.mean() if loss_weighted.dim() else loss_weighted,
            "label": flatten_subnetwork_dimension(label_t),
            "preds": flatten_subnetwork_dimension(y_pred),
            "aleatoric_std_map": flatten_subnetwork_dimension(aleatoric_std),
            "err_map": flatten_subnetwork_dimension(y_pred - label_t),
            "mask": flatten_subnetwork_dimension(mask_t) if mask_t is not None else None,
        }

    def validation_step(self, batch: Dict[str, torch.Tensor], batch_idx: int) -> Dict[str, torch.Tensor]:
        image, label = batch["image"], batch["label"]
        mask = batch.get("mask")
        S = self.num_subnetworks
        image = repeat_subnetworks(image, num_subnetworks=S)
        label = repeat_subnetworks(label, num_subnetworks=S)
        mask_t = repeat_subnetworks(mask, num_subnetworks=S) if mask is not None else None
        p1, p2 = self(image)
        if isinstance(self.loss_fn, LaplaceNLL) and p1.is_cuda and S <= 16 and not torch.is_grad_enabled():
            # one pass over (p1, p2, label): per-subnetwork NLL means, ensemble aggregation, combined-scale NLL, regression
            # metrics and the clipped std means (reference mimo_unet.py:157-183); the label is NOT repeated for the kernel
            from mimo_unet_b200 import functional as Fn
            v = Fn.validation_laplace(p1, p2, batch["label"], mask, self.loss_fn.eps_min, self.loss_fn.eps_max)
            y_mean = batch["label"].float()
            self._log_val_loss(v["val_loss"], v["val_loss_combined"])
            self._log_metrics(y_pred=v["preds"], y_true=y_mean, stage="val", precomputed=v["metrics"])
            bs = self._batch_size()
            self.log("metric_val/aleatoric_std_mean", v["aleatoric_std_mean"], batch_size=bs)
            self.log("metric_val/epistemic_std_mean", v["epistemic_std_mean"], batch_size=bs)
            return {"loss": v["val_loss"].mean(), "label": y_mean, "preds": v["preds"], "aleatoric_std_map": v["aleatoric_std"],
                    "epistemic_std_map": v["epistemic_std"], "err_map": v["err"], "mask": mask}
        val_loss = self.loss_fn.forward(p1, p2, label, mask=mask_t, reduce_mean=False).mean(dim=(0, 2, 3, 4))
        y_pred_mean, aleatoric_var, epistemic_var = compute_uncertainties(self.loss_fn, p1, p2)
        y_mean = label.mean(dim=1)
        combined_std = torch.sqrt(aleatoric_var + epistemic_var)
        aleatoric_std, epistemic_std = torch.sqrt(aleatoric_var), torch.sqrt(epistemic_var)
        combined_log_scale = self.loss_fn.calculate_dist_param(std=combined_std, log=True)
        val_loss_combined = self.loss_fn.forward(y_pred_mean, combined_log_scale, y_mean, mask=mask, reduce_mean=True)
        self._log_val_loss(val_loss, val_loss_combined)
        self._log_metrics(y_pred=y_pred_mean, y_true=y_mean, stage="val")
        self._log_uncertainties(aleatoric_std, epistemic_std)
        return {"loss": val_loss.mean(), "label": y_mean, "preds": y_pred_mean, "aleatoric_std_map": aleatoric_std,
                "epistemic_std_map": epistemic_std, "err_map": y_pred_mean - y_mean, "mask": mask}

    def configure_optimizers(self) -> Dict[str, Any]:
        # same optimizer as the reference (mimo_unet.py:185-201: Adam with L2-in-gradient weight decay); on CUDA the
        # single-kernel "fused" implementation of the identical update is used instead of ~10 foreach launches
        params = list(self.parameters())
        fused = bool(params) and all(p.is_cuda for p in params)
        optimizer = torch.optim.Adam(params, lr=self.learning_rate, weight_decay=self.weight_decay, fused=fused)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=self.scheduler_step_size, gamma=self.scheduler_gamma)
        return dict(optimizer=optimizer, lr_scheduler=scheduler, monitor="val_loss")

    @staticmethod
    def _compute_epistemic_std(y_hat: torch.Tensor) -> torch.Tensor:
        """[B,S,C,H,W] -> [B,C,H,W] unbiased std over subnetworks (zeros for S == 1)."""
        B, S, C, H, W = y_hat.shape
        if S == 1:
            return torch.zeros((B, C, H, W), device=y_hat.device)
        return (torch.sum((y_hat - y_hat.mean(dim=1, keepdim=True)) ** 2, dim=1) / (S - 1)) ** 0.5

    def _calculate_train_loss(self, p1: torch.Tensor, p2: torch.Tensor, y_true: torch.Tensor, mask: torch.Tensor = None,
                              _out: torch.Tensor = None, _metrics: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns (loss[S], loss*weights [S], weights[S]). Weights are read from the loss buffer BEFORE the new
        loss is added (reference mimo_unet.py:243-245). On the GPU with the Laplace loss everything -- NLL,
        per-subnetwork means, buffer read, buffer update and the gradient seed -- is one fused kernel pass."""
        if isinstance(self.loss_fn, (LaplaceNLL, GaussianNLL)) and p1.is_cuda:
            from mimo_unet_b200 import functional as Fn
            out = _out if _out is not None else torch.cat([p1, p2], dim=2)
            dev_buf = self.loss_buffer.device_state(p1.device)
            from mimo_unet_b200 import parallel as Par
            dp = Par.world_size() > 1
            sync = getattr(self.model, "_runtime", None) and self.model._runtime.grad_sync
            if dp and sync is not None:
                sync.join_loss_exchange()   # the previous step's buffer update (side stream) before this step reads the weights
            # data parallel: every rank's loss buffer receives the MEAN per-subnetwork loss over ranks, so the softmax
            # weights stay identical everywhere (the reference defines no multi-GPU behaviour)
            res = Fn.laplace_train_loss(out, y_true, mask=mask, loss_buffer=dev_buf, update_buffer=not dp,
                                        eps_min=self.loss_fn.eps_min, eps_max=self.loss_fn.eps_max, with_metrics=_metrics,
                                        gaussian=isinstance(self.loss_fn, GaussianNLL))
            total, loss, weights = res[:3]
            # r2/mae/mse/rmse of (mode, label) out of the same kernel pass (reference: compute_regression_metrics, 4 reductions)
            self._fused_metrics = res[3] if _metrics else None
            if dp:
                if sync is not None:
                    sync.exchange_loss(loss, dev_buf)     # all-reduce + buffer update on the side stream, joined next step
                else:
                    dev_buf.add(Par.allreduce_mean_(loss.detach().clone()))
            # `total` (= mean_s w_s loss_s) carries the autograd graph; expose it through the reference's
            # (loss, loss*weights, weights) triple so that loss_weighted.mean() == total, value and gradient
            loss_weighted = loss * weights + (total - (loss * weights).mean())
            return loss, loss_weighted, weights
        forward = self.loss_fn.forward(p1, p2, y_true, reduce_mean=False, mask=mask)
        loss = forward.mean(dim=(0, 2, 3, 4))
        weights = self.loss_buffer.get_weights().to(loss.device)
        self.loss_buffer.add(loss.detach())
        return loss, loss * weights, weights

    # ------------------------------------------------------------------------------------------ logging
    def _log_train_loss_and_weights(self, loss: torch.Tensor, weights: torch.Tensor) -> None:
        bs = self._batch_size()
        self.log("train_loss", loss.mean(), batch_size=bs)
        for i in range(loss.shape[0]):
            self.log(f"train_loss_{i}", loss[i], batch_size=bs)
            self.log(f"train_weight_{i}", weights[i], batch_size=bs)

    def _log_metrics(self, y_pred: torch.Tensor, y_true: torch.Tensor, stage: Literal["train", "val"] = "train",
                     precomputed: Dict[str, torch.Tensor] = None) -> None:
        metrics = precomputed if precomputed is not None else compute_regression_metrics(y_pred.flatten(), y_true.flatten())
        for name, value in metrics.items():
            self.log(f"metric_{stage}/{name}", value, on_step=(stage == "train"), on_epoch=True, metric_attribute=name,
                     batch_size=self._batch_size())

    def _log_val_loss(self, val_loss: torch.Tensor, val_loss_combined: torch.Tensor) -> None:
        bs = self._batch_size()
        self.log("val_loss", val_loss.mean(), batch_size=bs)
        for i in range(val_loss.shape[0]):
            self.log(f"val_loss_{i}", val_loss[i], batch_size=bs)
        self.log("val_loss_combined", val_loss_combined, batch_size=bs)

    def _log_uncertainties(self, aleatoric_std: torch.Tensor, epistemic_std: torch.Tensor) -> None:
        bs = self._batch_size()
        self.log("metric_val/aleatoric_std_mean", aleatoric_std.clip(0, 5).mean(), batch_size=bs)
        self.log("metric_val/epistemic_std_mean", epistemic_std.clip(0, 5).mean(), batch_size=bs)

    @staticmethod
    def add_model_specific_args(parent_parser: ArgumentParser) -> ArgumentParser:
        p = parent_parser.add_argument_group(title="MIMO UNet Model")
        for name, typ, default in (("num_subnetworks", int, 3), ("filter_base_count", int, 32), ("center_dropout_rate", float, 0.0),
                                   ("final_dropout_rate", float, 0.0), ("encoder_dropout_rate", float, 0.0),
                                   ("core_dropout_rate", float, 0.0), ("decoder_dropout_rate", float, 0.0),
                                   ("input_repetition_probability", float, 0.0), ("batch_repetitions", int, 1),
                                   ("loss", str, "laplace_nll"), ("learning_rate", float, 1e-3), ("weight_decay", float, 0.0),
                                   ("loss_buffer_size", int, 10), ("loss_buffer_temperature", float, 1.0),
                                   ("scheduler_step_size", int, 20), ("scheduler_gamma", float, 0.5)):
            p.add_argument(f"--{name}", type=typ, default=default)
        return parent_parser
