"""`DiffusionDataset` with the reference's interface (src/fdiff/dataloaders/datamodules.py:23-65): the DFT of the series, the per-feature
mean / unbiased standard deviation and the standardisation run on the GPU (fd_dft, fd_feature_stats, fd_standardise); tensors are returned
on the device the input lives on, like the reference's."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch.utils.data import Dataset

from . import _lib
from .fourier import _device_for, dft


def feature_mean_and_std(X: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """`X.mean(dim=0), X.std(dim=0)` of a (n, L, C) tensor (datamodules.py:52-53,153-161)."""
    assert X.dim() == 3, f"expected (n, max_len, n_channels), got {tuple(X.shape)}"
    lib = _lib.load()
    dev = _device_for(X)
    xd = X.detach().to(device=dev, dtype=torch.float32).contiguous()
    n, L, Cc = xd.shape
    mean = torch.empty(L, Cc, device=dev, dtype=torch.float32)
    std = torch.empty(L, Cc, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.fd_feature_stats(C.c_void_p(xd.data_ptr()), C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), n, L * Cc, dev.index,
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return mean.to(X.device), std.to(X.device)


def standardise(X: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """(X - mean) / std, or X * std + mean with `inverse=True`; X: (n, L, C), mean / std: (L, C)."""
    lib = _lib.load()
    dev = _device_for(X)
    xd = X.detach().to(device=dev, dtype=torch.float32).contiguous()
    n, L, Cc = xd.shape
    md = mean.detach().to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
    sd = std.detach().to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
    out = torch.empty_like(xd)
    with torch.cuda.device(dev):
        _lib.check(lib.fd_standardise(C.c_void_p(xd.data_ptr()), C.c_void_p(md.data_ptr()), C.c_void_p(sd.data_ptr()), C.c_void_p(out.data_ptr()),
                                      n, L * Cc, int(bool(inverse)), dev.index, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.to(X.device)


class DiffusionDataset(Dataset):
    """datamodules.py:23-65.  `standardized()` (not in the reference) returns the whole standardised tensor from one GPU pass."""

    def __init__(self, X: torch.Tensor, y: Optional[torch.Tensor] = None, fourier_transform: bool = False, standardize: bool = False,
                 X_ref: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        if fourier_transform:
            X = dft(X).detach()
        self.X = X
        self.y = y
        self.standardize = standardize
        if X_ref is None:
            X_ref = X
        elif fourier_transform:
            X_ref = dft(X_ref).detach()
        assert isinstance(X_ref, torch.Tensor)
        self.feature_mean, self.feature_std = feature_mean_and_std(X_ref)
        self._std_cache: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return len(self.X)

    def standardized(self) -> torch.Tensor:
        if self._std_cache is None:
            self._std_cache = standardise(self.X, self.feature_mean, self.feature_std)
        return self._std_cache

    def __getitem__(self, index: int) -> dict[str, torch.Tensor]:
        data = {"X": self.standardized()[index] if self.standardize else self.X[index]}
        if self.y is not None:
            data["y"] = self.y[index]
        return data
