"""Host mirror of the reference's loss factory (src/fdiff/utils/losses.py:12-127), forward / evaluation half.

`get_sde_loss_fn(scheduler, train=False, ...)` returns a `loss_fn(model, batch)` with the reference's signature, RNG consumption order
(`torch.rand` for the diffusion times when the batch has none, then `torch.randn_like(X)`) and value; the arithmetic — perturbation,
score network at per-series times, weighted squared error and both reductions — is ONE library call (fd_sde_loss).  The training loss
(`train=True`) additionally needs dropout and a backward pass through the score network, which this library does not have
(DESIGN.md §6): asking for it raises instead of returning a forward value that could not be trained on.
"""
from __future__ import annotations

from typing import Callable

import torch

from .batch import DiffusableBatch
from .schedulers import SDE


def get_sde_loss_fn(scheduler: SDE, train: bool, reduce_mean: bool = True,
                    likelihood_weighting: bool = False) -> Callable[[torch.nn.Module, DiffusableBatch], torch.Tensor]:
    def loss_fn(model: torch.nn.Module, batch: DiffusableBatch) -> torch.Tensor:
        if train:
            raise NotImplementedError(
                "fourierdiffusion_b200 evaluates the score-matching loss forward only (validation); the training step needs "
                "dropout and the backward pass of the score network, which are outside this library")
        model.eval()
        X = batch.X
        timesteps = batch.timesteps
        if timesteps is None:  # losses.py:58-62
            timesteps = torch.rand(X.shape[0], device=X.device) * (scheduler.T - scheduler.eps) + scheduler.eps
        z = torch.randn_like(X)  # losses.py:65
        if getattr(model, "noise_scheduler", scheduler) is not scheduler:
            raise ValueError("the loss's scheduler must be the score model's noise_scheduler (the library handle holds one scheduler)")
        loss, _ = model.engine(X.device).sde_loss(X, timesteps, z, likelihood_weighting=likelihood_weighting, reduce_mean=reduce_mean)
        return loss.to(X.device)

    return loss_fn
