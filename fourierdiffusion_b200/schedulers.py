"""Host mirror of the reference's SDE schedulers (src/fdiff/schedulers/sde.py) for the sampling path.

Same class names, constructor arguments and attributes as the reference (`noise_scaling`, `eps`, `G`, `timesteps`,
`step_size`, `beta_0/beta_1`, `sigma_min/sigma_max`), so a `ScoreModule` built with these is indistinguishable to the
sampler from one built with the reference's.  `prior_sampling` and `step` run in the CUDA library (fd_prior / fd_step);
`marginal_prob` / `add_noise` (the loss's forward perturbation) go through fd_perturb.
"""
from __future__ import annotations

import abc
import ctypes as C
import math
from collections import namedtuple
from typing import Optional

import torch

from . import _lib

SamplingOutput = namedtuple("SamplingOutput", ["prev_sample"])


class SDE(abc.ABC):
    def __init__(self, fourier_noise_scaling: bool = False, eps: float = 1e-5):
        super().__init__()
        self.noise_scaling = fourier_noise_scaling
        self.eps = eps
        self.G: Optional[torch.Tensor] = None
        self._engines: dict = {}

    @property
    def T(self) -> float:
        return 1.0

    # -- host-side state (sde.py:42-64): tiny, computed with the same torch ops so the values are bit-identical ------------
    def set_noise_scaling(self, max_len: int) -> None:
        G = torch.ones(max_len)
        if self.noise_scaling:
            G = 1 / (math.sqrt(2)) * G
            G[0] *= math.sqrt(2)
            if max_len % 2 == 0:
                G[max_len // 2] *= math.sqrt(2)
        self.G = G
        self._engines.clear()

    @property
    def G_matrix(self) -> torch.Tensor:
        assert self.G is not None
        return torch.diag(self.G)

    def set_timesteps(self, num_diffusion_steps: int) -> None:
        self.timesteps = torch.linspace(1.0, self.eps, num_diffusion_steps)
        self.step_size = self.timesteps[0] - self.timesteps[1]

    # -- device work ---------------------------------------------------------------------------------------------------
    def _engine(self, max_len: int, n_channels: int, device: torch.device):
        """A scheduler-only fd_handle (no score network) for stand-alone prior_sampling()/step() calls."""
        from .engine import Engine, scheduler_params

        if not torch.cuda.is_available():
            raise _lib.FdError("no CUDA device visible: fourierdiffusion_b200 has no CPU fallback")
        device = torch.device(device)
        if device.type != "cuda" or device.index is None:  # always an INDEXED device: the handle, its streams and tensors must agree
            device = torch.device("cuda", torch.cuda.current_device())
        key = (max_len, n_channels, device.index)
        eng = self._engines.get(key)
        if eng is None:
            kind, p0, p1 = scheduler_params(self)
            cfg = _lib.FdConfig(struct_size=C.sizeof(_lib.FdConfig), device=device.index, model_kind=_lib.FD_MODEL_TRANSFORMER,
                                max_len=max_len, n_channels=n_channels, d_model=1, n_head=1, num_layers=0, d_ff=1, sched_kind=kind,
                                sched_p0=p0, sched_p1=p1, fourier_noise_scaling=int(bool(self.noise_scaling)), math_mode=_lib.FD_MATH_FP32)
            eng = Engine(cfg, device)
            if self.G is None:
                self.set_noise_scaling(max_len)
            assert self.G is not None and self.G.shape[0] == max_len
            eng.set_weight("noise_scheduler.G", self.G)
            self._engines[key] = eng
        return eng

    def prior_sampling(self, shape: tuple[int, ...]) -> torch.Tensor:
        """G ⊙ z, z ~ N(0, I) drawn from torch's CPU generator exactly like the reference (sde.py:79-87); returned on the CPU."""
        z = torch.randn(*shape)
        eng = self._engine(shape[1], shape[2], torch.device("cuda"))
        return eng.prior(z).cpu()

    @abc.abstractmethod
    def step(self, model_output: torch.Tensor, timestep: float, sample: torch.Tensor) -> SamplingOutput: ...

    def _step(self, model_output: torch.Tensor, timestep: float, sample: torch.Tensor) -> SamplingOutput:
        assert self.G is not None
        assert self.step_size > 0
        z = torch.randn_like(sample)  # sde.py:238 / :155
        eng = self._engine(sample.shape[1], sample.shape[2], sample.device)
        out = eng.step(sample, model_output, z, float(timestep), float(self.step_size))
        return SamplingOutput(prev_sample=out.to(sample.device))

    def marginal_prob(self, x: torch.Tensor, t: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """(mean, std) of the perturbation kernel p(x_t | x_0): mean (batch, max_len, n_channels), std (batch, max_len) — sde.py:108-123
        (VE), :187-210 (VP).  Evaluated by fd_perturb (zero noise gives the mean, the per-series scalar times G the std); returned on
        x's device."""
        if self.G is None:
            self.set_noise_scaling(x.shape[1])
        assert self.G is not None
        eng = self._engine(x.shape[1], x.shape[2], x.device)
        mean, std_scalar = eng.perturb(x, t, torch.zeros_like(x))
        std = std_scalar.view(-1, 1) * self.G.to(std_scalar.device)
        return mean.to(x.device), std.to(x.device)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """mean(x0, t) + noise, the noise already scaled by the caller (sde.py:66-77)."""
        mean, _ = self.marginal_prob(original_samples, timesteps)
        return mean + noise


class VEScheduler(SDE):
    def __init__(self, sigma_min: float = 0.01, sigma_max: float = 50.0, fourier_noise_scaling: bool = False, eps: float = 1e-5):
        super().__init__(fourier_noise_scaling=fourier_noise_scaling, eps=eps)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max

    def step(self, model_output: torch.Tensor, timestep: float, sample: torch.Tensor) -> SamplingOutput:
        return self._step(model_output, timestep, sample)


class VPScheduler(SDE):
    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, fourier_noise_scaling: bool = False, eps: float = 1e-5):
        super().__init__(fourier_noise_scaling=fourier_noise_scaling, eps=eps)
        self.beta_0 = beta_min
        self.beta_1 = beta_max

    def get_beta(self, timestep: float) -> float:
        return self.beta_0 + timestep * (self.beta_1 - self.beta_0)

    def step(self, model_output: torch.Tensor, timestep: float, sample: torch.Tensor) -> SamplingOutput:
        return self._step(model_output, timestep, sample)
