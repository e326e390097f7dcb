"""dft / idft / spectral_density with the reference's signatures (src/fdiff/utils/fourier.py:8-124), computed by the CUDA library.

`dft`: ortho rFFT along dim 1 of (batch, max_len, n_channels), packed real [Re X_0..X_{L//2} | Im X_1..X_{ceil(L/2)-1}].
`idft`: the inverse; optionally fuses the de-standardisation `x * std + mean` of cmd/sample.py:76-78 in front.
Input on the CPU -> result on the CPU (like the reference, whose idft is CPU-only, fourier.py:66); input on a CUDA
device -> result stays there.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _device_for(x: torch.Tensor) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.FdError("no CUDA device visible: fourierdiffusion_b200 has no CPU fallback")
    if x.device.type == "cuda":
        return x.device
    return torch.device("cuda", torch.cuda.current_device())


def _run(x: torch.Tensor, inverse: bool, mean: Optional[torch.Tensor], std: Optional[torch.Tensor]) -> torch.Tensor:
    assert x.dim() == 3, f"expected (batch_size, max_len, n_channels), got {tuple(x.shape)}"
    lib = _lib.load()
    dev = _device_for(x)
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, L, Cc = xd.shape
    out = torch.empty_like(xd)
    if B == 0:
        return out.to(x.device)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        if inverse:
            md = sd = None
            if mean is not None or std is not None:
                assert mean is not None and std is not None, "mean and std must be given together"
                md = mean.detach().to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
                sd = std.detach().to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
            _lib.check(lib.fd_idft(C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), B, L, Cc,
                                   None if md is None else C.c_void_p(md.data_ptr()),
                                   None if sd is None else C.c_void_p(sd.data_ptr()), dev.index, stream))
        else:
            _lib.check(lib.fd_dft(C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), B, L, Cc, dev.index, stream))
    return out.to(x.device)


def dft(x: torch.Tensor) -> torch.Tensor:
    """fourier.py:8-45."""
    return _run(x, False, None, None)


def idft(x: torch.Tensor, mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fourier.py:48-87; with (mean, std) of shape (max_len, n_channels): idft(x * std + mean) (cmd/sample.py:76-82)."""
    return _run(x, True, mean, std)


def spectral_density(x: torch.Tensor, apply_dft: bool = True) -> torch.Tensor:
    """fourier.py:90-124: |X_k|^2 of the ortho rFFT, shape (batch_size, max_len // 2 + 1, n_channels).  `apply_dft=False`: `x` is already
    the packed spectrum.  This is the front-end of the spectral metric that follows the sampler (metrics.py:76-84)."""
    assert x.dim() == 3, f"expected (batch_size, max_len, n_channels), got {tuple(x.shape)}"
    lib = _lib.load()
    dev = _device_for(x)
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, L, Cc = xd.shape
    out = torch.empty(B, L // 2 + 1, Cc, device=dev, dtype=torch.float32)
    if B == 0:
        return out.to(x.device)
    scratch = torch.empty_like(xd) if apply_dft else None
    with torch.cuda.device(dev):
        _lib.check(lib.fd_spectral_density(C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()),
                                           None if scratch is None else C.c_void_p(scratch.data_ptr()), B, L, Cc, int(bool(apply_dft)),
                                           dev.index, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.to(x.device)
