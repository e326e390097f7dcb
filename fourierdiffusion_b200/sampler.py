"""Drop-in replacement for the reference's `DiffusionSampler` (src/fdiff/sampling/sampler.py:11-122).

Selected from the reference's host shell by one config line (cmd/conf/sampler/default.yaml:1):

    _target_: fourierdiffusion_b200.sampler.DiffusionSampler

Same constructor `(score_model, sample_batch_size)` and `sample(num_samples, num_diffusion_steps=None)` contract: returns
a new CPU fp32 tensor `(n, max_len, n_channels)` in the model's domain, `n` following the reference's batch rule
(`max(1, n // bs)` batches, remainder dropped, sampler.py:63,75-78).  The whole inner loop — prior, N x (score network +
scheduler update with fresh noise) — runs inside libfdiff_b200 (fd_sample_host); torch is plumbing.

`score_model` is duck-typed: either this package's host mirror (`score_models.ScoreModule` ...) or the reference's own
module; only `.state_dict()`, `.noise_scheduler`, `.n_channels`, `.max_len`, `.d_model`, `.backbone`,
`.num_training_steps` and `.eval()` are read.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, distributed
from .batch import DiffusableBatch
from .engine import Engine, scheduler_params


class DiffusionSampler:
    def __init__(self, score_model, sample_batch_size: int, seed: Optional[int] = None, math_mode: Optional[int] = None,
                 device=None) -> None:
        self.score_model = score_model
        self.noise_scheduler = score_model.noise_scheduler
        self.sample_batch_size = sample_batch_size
        self.n_channels = score_model.n_channels
        self.max_len = score_model.max_len
        # extras (all optional, so the reference's two construction sites — cmd/sample.py:59-60, callbacks.py:41-46 — work)
        self.seed = seed              # None: a fresh 63-bit seed is drawn from torch's global CPU generator per sample() call
        self.math_mode = math_mode    # None: library default (TF32 tensor-core path where a specialised kernel exists)
        self._device = device
        self._engine: Optional[Engine] = None
        self._engine_sig = None
        scheduler_params(self.noise_scheduler)  # raises NotImplementedError("Scheduler not recognized.") like sampler.py:118-119

    # -- engine ---------------------------------------------------------------------------------------------------------
    def _weights_signature(self):
        return tuple((k, v.data_ptr(), v._version) for k, v in self.score_model.state_dict(keep_vars=True).items())

    def engine(self) -> Engine:
        sig = self._weights_signature()
        if self._engine is None or sig != self._engine_sig:
            if hasattr(self.score_model, "engine") and self._device is None:
                self._engine = self.score_model.engine(math_mode=self.math_mode)  # share the handle with forward()
            else:
                self._engine = Engine.for_score_model(self.score_model, self._device, self.math_mode)
            self._engine_sig = sig
        return self._engine

    # -- the reference's methods --------------------------------------------------------------------------------------
    def reverse_diffusion_step(self, batch: DiffusableBatch) -> torch.Tensor:
        """One reverse step with noise drawn from torch's generator like the reference (sampler.py:24-43)."""
        X = batch.X
        timesteps = batch.timesteps
        assert timesteps is not None and timesteps.size(0) == len(batch)
        assert torch.min(timesteps) == torch.max(timesteps)
        eng = self.engine()
        t = timesteps[0].item()
        score = eng.score(X, float(t))
        z = torch.randn_like(X)  # sde.py:238
        assert self.noise_scheduler.step_size > 0
        out = eng.step(X, score, z, float(t), float(self.noise_scheduler.step_size))
        return out.to(X.device)

    def sample_prior(self, batch_size: int) -> torch.Tensor:
        """G ⊙ z with z from torch's CPU generator (sde.py:79-87 via sampler.py:111-122)."""
        eng = self.engine()
        z = torch.randn(batch_size, self.max_len, self.n_channels)
        return eng.prior(z)

    def _num_steps(self, num_diffusion_steps: Optional[int]) -> int:
        n = self.score_model.num_training_steps if num_diffusion_steps is None else num_diffusion_steps  # sampler.py:52-56
        assert float(n) == int(n) and int(n) >= 2, f"num_diffusion_steps must be an integer >= 2, got {n}"
        return int(n)

    def sample(self, num_samples: int, num_diffusion_steps: Optional[int] = None, *, prior_z: Optional[torch.Tensor] = None,
               noise: Optional[torch.Tensor] = None, return_device: bool = False) -> torch.Tensor:
        """Generate series.  Keyword-only extras (not in the reference): `prior_z` (n, L, C) and `noise` (N, n, L, C) inject
        the randomness for parity runs; `return_device=True` keeps the result on the GPU (no `X.cpu()`)."""
        self.score_model.eval()
        n_steps = self._num_steps(num_diffusion_steps)
        sch = self.noise_scheduler
        sch.set_timesteps(n_steps)  # sde.py:62-64: the timestep grid and step size stay host-side torch values
        if getattr(sch, "G", None) is None:
            sch.set_noise_scaling(self.max_len)
        eng = self.engine()

        bs = self.sample_batch_size
        num_batches = max(1, num_samples // bs)  # sampler.py:63
        sizes = [min(num_samples - b * bs, bs) for b in range(num_batches)]  # sampler.py:75-78
        n_total = sum(sizes)
        seed = self.seed if self.seed is not None else int(torch.randint(0, 2**62, (1,)).item())
        rank, ws = distributed.world()
        lo, hi = distributed.shard_range(n_total, rank, ws)

        ts = sch.timesteps.detach().float().cpu().contiguous()
        dt = float(sch.step_size)
        dev = eng.device
        if ws == 1 and not return_device:
            out = torch.empty(n_total, self.max_len, self.n_channels, dtype=torch.float32, pin_memory=True)
        else:
            out = torch.empty(hi - lo, self.max_len, self.n_channels, dtype=torch.float32, device=dev)

        # chunk the rank's index range at the reference's batch boundaries (results do not depend on the chunking)
        start = lo
        while start < hi:
            stop = min(hi, (start // bs + 1) * bs)
            b = stop - start
            pz = None if prior_z is None else prior_z[start:stop]
            nz = None if noise is None else noise[:, start:stop]
            if ws == 1 and not return_device:
                eng.sample_host(b, ts, dt, seed=seed, first_series=start, prior_z=pz, noise=nz, out=out[start:stop])
            else:
                out[start - lo : stop - lo] = eng.sample(b, ts, dt, seed=seed, first_series=start, prior_z=pz, noise=nz)
            start = stop
        if ws > 1:
            out = distributed.all_gather_series(out, n_total)  # the one collective of the path
        if return_device:
            return out
        if out.device.type == "cuda":
            host = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(out)
            torch.cuda.current_stream(dev).synchronize()
            out = host
        return out


    def sample_time_domain(self, num_samples: int, num_diffusion_steps: Optional[int] = None, feature_mean: Optional[torch.Tensor] = None,
                           feature_std: Optional[torch.Tensor] = None, *, prior_z: Optional[torch.Tensor] = None,
                           noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`sample` followed by the two steps the reference's runner applies to its result (cmd/sample.py:76-82): de-standardise with the
        (max_len, n_channels) statistics of the DFT'd training set (datamodules.py:52-53,153-161) and `idft` — here fused into one kernel
        that reads the samples where the sampler left them on the GPU.  Returns the generated TIME-DOMAIN series as a CPU fp32 tensor."""
        from .fourier import idft

        X = self.sample(num_samples, num_diffusion_steps, prior_z=prior_z, noise=noise, return_device=True)
        Y = idft(X, mean=feature_mean, std=feature_std)
        host = torch.empty(Y.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(Y)
        torch.cuda.current_stream(Y.device).synchronize()
        return host


Sampler = DiffusionSampler
