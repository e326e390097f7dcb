"""Boundary type of the hot path.

`ScoreModule.forward(batch)` and `DiffusionSampler.reverse_diffusion_step(batch)` take the reference's batch object
(`fdiff.utils.dataclasses.DiffusableBatch`, src/fdiff/utils/dataclasses.py:7-18); they only ever read `.X`, `.y`, `.timesteps`,
`len(batch)` and `.device`, so any object with those attributes works (the reference's own class included).  This host-side stand-in
additionally checks the shapes the CUDA path relies on.
"""
from __future__ import annotations

from typing import Optional

import torch


class DiffusableBatch:
    """X: (batch, max_len, n_channels) series; y: optional labels (unused on the sampling path); timesteps: (batch,) diffusion times."""

    __slots__ = ("X", "y", "timesteps")

    def __init__(self, X: torch.Tensor, y: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None):
        if X.dim() != 3:
            raise ValueError(f"X must be (batch, max_len, n_channels), got {tuple(X.shape)}")
        if timesteps is not None and (timesteps.dim() != 1 or timesteps.shape[0] != X.shape[0]):
            raise ValueError(f"timesteps must be ({X.shape[0]},), got {tuple(timesteps.shape)}")
        self.X, self.y, self.timesteps = X, y, timesteps

    def __len__(self) -> int:
        return int(self.X.shape[0])

    def __repr__(self) -> str:
        t = None if self.timesteps is None else tuple(self.timesteps.shape)
        return f"DiffusableBatch(X={tuple(self.X.shape)}, y={'None' if self.y is None else tuple(self.y.shape)}, timesteps={t})"

    @property
    def device(self) -> torch.device:
        return self.X.device
