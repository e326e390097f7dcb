"""Host mirror of the reference's score modules (src/fdiff/models/score_models.py, transformer.py) for sampling.

The classes keep the reference's names, constructor signatures, attributes and `state_dict` keys,
so checkpoints load with `load_state_dict` and, built under the same `torch.manual_seed`, they hold the very same
weights.  torch.nn modules are used purely as PARAMETER CONTAINERS: `forward` never runs them — it hands the batch to the
CUDA library (fd_score / fd_score_t).  The evaluation loss (`validation_step`) runs in the library too (fd_sde_loss); the training step,
optimisers and Lightning hooks are out of scope.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .batch import DiffusableBatch
from .losses import get_sde_loss_fn
from .schedulers import SDE


class PositionalEncoding(nn.Module):
    """Container for the learnable positional table (transformer.py:8-15); the add happens inside the embed kernel."""

    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        self.embedding = nn.Embedding(num_embeddings=max_len, embedding_dim=d_model, max_norm=math.sqrt(d_model))


class GaussianFourierProjection(nn.Module):
    """Container for the Gaussian random time features (transformer.py:62-75); evaluated by the time-embedding kernel."""

    def __init__(self, d_model: int, scale: float = 30.0):
        super().__init__()
        self.d_model = d_model
        self.W = nn.Parameter(torch.randn((d_model + 1) // 2) * scale, requires_grad=False)
        self.dense = nn.Linear(d_model, d_model)


class ScoreModule(nn.Module):
    """Transformer-encoder score network (score_models.py:22-94)."""

    def __init__(self, n_channels: int, max_len: int, noise_scheduler: SDE, fourier_noise_scaling: bool = True, d_model: int = 60,
                 num_layers: int = 3, n_head: int = 12, num_training_steps: int = 1000, lr_max: float = 1e-3,
                 likelihood_weighting: bool = False) -> None:
        super().__init__()
        self.max_len = max_len
        self.n_channels = n_channels
        self.noise_scheduler = noise_scheduler
        self.num_warmup_steps = num_training_steps // 10
        self.num_training_steps = num_training_steps
        self.lr_max = lr_max
        self.d_model = d_model
        self.scale_noise = fourier_noise_scaling
        self.likelihood_weighting = likelihood_weighting
        self.training_loss_fn, self.validation_loss_fn = self.set_loss_fn()  # score_models.py:50
        if not hasattr(noise_scheduler, "noise_scaling"):
            raise NotImplementedError(f"Scheduler {noise_scheduler} not implemented yet, cannot set time encoder.")
        # parameter containers, created in the reference's order so that a shared seed gives shared weights
        self.pos_encoder: Optional[PositionalEncoding] = PositionalEncoding(d_model=d_model, max_len=self.max_len)
        self.time_encoder = GaussianFourierProjection(d_model=d_model)
        self.embedder = nn.Linear(in_features=n_channels, out_features=d_model)
        self.unembedder = nn.Linear(in_features=d_model, out_features=n_channels)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, batch_first=True)
        self.backbone = nn.TransformerEncoder(encoder_layer=layer, num_layers=num_layers, enable_nested_tensor=False)
        self._engines: dict = {}
        self.math_mode: Optional[int] = None  # None -> library default (TF32 tensor-core path where available)

    # -- engine cache ------------------------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device=None, math_mode: Optional[int] = None):
        """The fd_handle holding this module's weights on `device` (built on first use, rebuilt when weights change)."""
        from .engine import Engine

        if not torch.cuda.is_available():
            raise _lib.FdError("no CUDA device visible: fourierdiffusion_b200 has no CPU fallback")
        if device is None or torch.device(device).type != "cuda":
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        mm = self.math_mode if math_mode is None else math_mode
        key = (device.index, mm)
        sig = self._signature()
        hit = self._engines.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        eng = Engine.for_score_model(self, device, mm)
        self._engines[key] = (sig, eng)
        return eng

    # -- forward = score ------------------------------------------------------------------------------------------------
    def forward(self, batch: DiffusableBatch) -> torch.Tensor:
        X = batch.X
        assert X.size()[1:] == (self.max_len, self.n_channels), (
            f"X has wrong shape, should be {(X.size(0), self.max_len, self.n_channels)}, but is {X.size()}"
        )
        timesteps = batch.timesteps
        assert timesteps is not None and timesteps.size(0) == len(batch)
        eng = self.engine(X.device)
        tcpu = timesteps.detach().float().cpu()
        if bool((tcpu == tcpu[0]).all()):
            out = eng.score(X, float(tcpu[0]))  # the sampler's case: the persistent-stack path keyed on one time-embedding row
        else:  # per-series times (training / validation style batches): a (batch, d_model) time-embedding table, same kernels
            out = eng.score_t(X, timesteps)
        return out.to(X.device)

    # -- evaluation loss (score_models.py:110-121, :132-152) -----------------------------------------------------------
    def set_loss_fn(self):
        if not isinstance(self.noise_scheduler, SDE) and not hasattr(self.noise_scheduler, "noise_scaling"):
            raise NotImplementedError(f"Scheduler {self.noise_scheduler} not implemented yet, cannot set loss function.")
        return (get_sde_loss_fn(scheduler=self.noise_scheduler, train=True, likelihood_weighting=self.likelihood_weighting),
                get_sde_loss_fn(scheduler=self.noise_scheduler, train=False, likelihood_weighting=self.likelihood_weighting))

    def validation_step(self, batch: DiffusableBatch, batch_idx: int = 0, dataloader_idx: int = 0) -> torch.Tensor:
        """The validation loss of one batch (the value the reference logs as `val/loss`, score_models.py:110-121)."""
        return self.validation_loss_fn(self, batch)


class LSTMScoreModule(ScoreModule):
    """LSTM score network (score_models.py:249-317): 10 residual single-layer LSTMs, no positional table."""

    def __init__(self, n_channels: int, max_len: int, noise_scheduler: SDE, fourier_noise_scaling: bool = True, d_model: int = 72,
                 num_layers: int = 3, num_training_steps: int = 1000, lr_max: float = 1e-3, likelihood_weighting: bool = False) -> None:
        super().__init__(n_channels=n_channels, max_len=max_len, noise_scheduler=noise_scheduler,
                         fourier_noise_scaling=fourier_noise_scaling, d_model=d_model, num_layers=num_layers, n_head=1,
                         num_training_steps=num_training_steps, lr_max=lr_max, likelihood_weighting=likelihood_weighting)
        self.backbone = nn.ModuleList(  # type: ignore[assignment]
            [nn.LSTM(input_size=d_model, hidden_size=d_model, batch_first=True, bidirectional=False) for _ in range(num_layers)]
        )
        self.pos_encoder = None


class MLPScoreModule(ScoreModule):
    """MLP score network (score_models.py:169-246): flatten, embed, residual Linear-ReLU-Linear blocks, unembed."""

    def __init__(self, n_channels: int, max_len: int, noise_scheduler: SDE, fourier_noise_scaling: bool = True, d_model: int = 72,
                 d_mlp: int = 512, num_layers: int = 3, num_training_steps: int = 1000, lr_max: float = 1e-3,
                 likelihood_weighting: bool = False) -> None:
        super().__init__(n_channels=n_channels, max_len=max_len, noise_scheduler=noise_scheduler,
                         fourier_noise_scaling=fourier_noise_scaling, d_model=d_model, num_layers=num_layers, n_head=1,
                         num_training_steps=num_training_steps, lr_max=lr_max, likelihood_weighting=likelihood_weighting)
        self.embedder = nn.Linear(in_features=max_len * n_channels, out_features=d_model)
        self.unembedder = nn.Linear(in_features=d_model, out_features=max_len * n_channels)
        # same child indices as torchvision.ops.MLP(in, [d_mlp, d_model], dropout): 0 Linear, 1 ReLU, 2 Dropout, 3 Linear, 4 Dropout
        self.backbone = nn.ModuleList(  # type: ignore[assignment]
            [nn.Sequential(nn.Linear(d_model, d_mlp), nn.ReLU(), nn.Dropout(0.1), nn.Linear(d_mlp, d_model), nn.Dropout(0.1))
             for _ in range(num_layers)]
        )
        self.pos_encoder = None
