"""Import-only stand-in for `pytorch_lightning` (TEST INFRASTRUCTURE, not product code).

The image has no Lightning. The reference's hot path (`fdiff.models.score_models`,
`fdiff.sampling.sampler`) only needs `LightningModule` to behave like an `nn.Module` with
`save_hyperparameters`, `log_dict` and a `device` property, so this stub lets the UNMODIFIED
reference import and run its CPU sampler in the build container. It is used only by
`oracle/ref_loader.py` (golden-vector generation + oracle pinning); it never ships on a product path.
"""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, *args, **kwargs):
        return None

    def log_dict(self, *args, **kwargs):
        return None

    def log(self, *args, **kwargs):
        return None

    @property
    def device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @classmethod
    def load_from_checkpoint(cls, *args, **kwargs):
        raise NotImplementedError("stub: real Lightning is not installed")


class LightningDataModule:
    def __init__(self, *args, **kwargs) -> None:
        pass


class Callback:
    pass


class Trainer:
    def __init__(self, *args, **kwargs) -> None:
        raise NotImplementedError("stub: real Lightning is not installed")
