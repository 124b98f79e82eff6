"""`lightning` is optional: when it is missing (as in the build container) a minimal stand-in keeps the
LightningModule-facing surface importable and testable."""
import torch

try:  # pragma: no cover - depends on the environment
    import lightning.pytorch as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    pl = None
    HAVE_LIGHTNING = False

    class _Namespace(dict):
        __getattr__ = dict.get

    class LightningModule(torch.nn.Module):
        """Just enough of pl.LightningModule for MimoUnetModel / EnsembleModule outside a Trainer."""

        def __init__(self):
            super().__init__()
            self.hparams = _Namespace()
            self.logged = {}
            self.trainer = None

        def save_hyperparameters(self, *args, **kwargs):
            import inspect
            if args and isinstance(args[0], dict):
                self.hparams.update(args[0])
                return
            frame = inspect.currentframe().f_back
            names = inspect.getfullargspec(type(self).__init__).args[1:]
            self.hparams.update({n: frame.f_locals[n] for n in names if n in frame.f_locals})

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        @classmethod
        def load_from_checkpoint(cls, path, **kwargs):
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            import inspect
            names = inspect.getfullargspec(cls.__init__).args[1:]
            model = cls(**{k: v for k, v in hp.items() if k in names})
            sd = {k.replace("model._orig_mod.", "model."): v for k, v in ckpt["state_dict"].items()}
            model.load_state_dict(sd)
            return model
