"""Engine: one libfdiff_b200 handle bound to one score module on one GPU.

Host-side plumbing only — reads the (reference-layout) `state_dict` and scheduler scalars of a score module, uploads
them through the C ABI and exposes the per-phase and whole-loop entry points on torch tensors.  torch is used for device
memory, streams and pinned host buffers; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Optional

import torch

from . import _lib
from ._lib import FdConfig, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def renorm_fixed_point(table: torch.Tensor, max_norm: float) -> torch.Tensor:
    """Fixed point of nn.Embedding(max_norm)'s in-place row renormalisation (reference transformer.py:13-15): rows with
    ||row||_2 > max_norm are scaled by max_norm / (norm + 1e-7) on every lookup until none moves."""
    table = table.detach().clone().float().cpu()
    idx = torch.arange(table.shape[0])
    for _ in range(16):
        before = table.clone()
        torch.embedding_renorm_(table, idx, max_norm, 2.0)
        if torch.equal(before, table):
            break
    return table


def model_kind_of(score_model) -> int:
    """Kernel family from the module's class, like the reference's get_model_type (utils/extraction.py:58-76)."""
    names = {c.__name__ for c in type(score_model).__mro__}
    if "LSTMScoreModule" in names:
        return _lib.FD_MODEL_LSTM
    if "MLPScoreModule" in names:
        return _lib.FD_MODEL_MLP
    if "ScoreModule" in names:
        return _lib.FD_MODEL_TRANSFORMER
    raise NotImplementedError(f"Score model {type(score_model).__name__} not recognized.")


def scheduler_params(noise_scheduler):
    """(kind, p0, p1) from a reference-style scheduler object (sde.py:93-106 VE, :171-185 VP)."""
    names = {c.__name__ for c in type(noise_scheduler).__mro__}
    if "VPScheduler" in names or hasattr(noise_scheduler, "beta_0"):
        return _lib.FD_SCHED_VP, float(noise_scheduler.beta_0), float(noise_scheduler.beta_1)
    if "VEScheduler" in names or hasattr(noise_scheduler, "sigma_min"):
        return _lib.FD_SCHED_VE, float(noise_scheduler.sigma_min), float(noise_scheduler.sigma_max)
    raise NotImplementedError("Scheduler not recognized.")  # sampler.py:118-119


def extract_score_model(score_model):
    """Everything the library needs from a score module, read duck-typed — works on this package's host mirror and on
    the reference's own `ScoreModule` / `LSTMScoreModule` / `MLPScoreModule` alike (a build-container test checks the latter).
    Pure host function, no GPU: returns (fd_config fields, {state_dict key: fp32 CPU tensor}); the positional table comes back at the
    fixed point of nn.Embedding(max_norm)'s renormalisation and the scheduler's G vector under the key "noise_scheduler.G"."""
    sched = score_model.noise_scheduler
    kind = model_kind_of(score_model)
    skind, p0, p1 = scheduler_params(sched)
    sd = {k: v.detach() for k, v in score_model.state_dict().items()}
    D = int(score_model.d_model)
    n_head, d_ff, num_layers = 1, 0, 0
    if kind == _lib.FD_MODEL_TRANSFORMER:
        layer0 = score_model.backbone.layers[0]
        n_head = int(layer0.self_attn.num_heads)
        d_ff = int(layer0.linear1.out_features)
        num_layers = len(score_model.backbone.layers)
    elif kind == _lib.FD_MODEL_LSTM:
        num_layers = len(score_model.backbone)
    else:
        num_layers = len(score_model.backbone)
        d_ff = int(sd["backbone.0.0.weight"].shape[0])
    fields = dict(
        model_kind=kind,
        max_len=int(score_model.max_len),
        n_channels=int(score_model.n_channels),
        d_model=D,
        n_head=n_head,
        num_layers=num_layers,
        d_ff=d_ff,
        sched_kind=skind,
        sched_p0=p0,
        sched_p1=p1,
        fourier_noise_scaling=int(bool(sched.noise_scaling)),
    )
    if getattr(sched, "G", None) is None:
        sched.set_noise_scaling(int(score_model.max_len))  # the reference's tests call this by hand (test_sampling.py:28-29)
    weights = {"noise_scheduler.G": sched.G.detach().float().cpu().contiguous()}
    for name, tensor in sd.items():
        if name == "pos_encoder.embedding.weight":
            tensor = renorm_fixed_point(tensor, math.sqrt(D))
        weights[name] = tensor.detach().to(device="cpu", dtype=torch.float32).contiguous()
    return fields, weights


class Engine:
    """Owns an fd_handle.  Build with `Engine.for_score_model(model, device)`."""

    def __init__(self, cfg: FdConfig, device: torch.device):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FdError("no CUDA device visible: fourierdiffusion_b200 runs only on a B200 (no CPU fallback)")
        self.device = torch.device(device)
        self.cfg = cfg
        handle = C.c_void_p()
        check(self.lib.fd_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self._finalizer = weakref.finalize(self, self.lib.fd_destroy, handle)
        self.L, self.C, self.D = cfg.max_len, cfg.n_channels, cfg.d_model

    # ---- construction --------------------------------------------------------------------------------------------
    @classmethod
    def for_score_model(cls, score_model, device=None, math_mode: Optional[int] = None) -> "Engine":
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.FdError("fourierdiffusion_b200 needs a CUDA device (no CPU fallback)")
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        device = torch.device("cuda", dev_index)
        fields, weights = extract_score_model(score_model)
        cfg = FdConfig(struct_size=C.sizeof(FdConfig), device=dev_index,
                       math_mode=_lib.FD_MATH_TF32 if math_mode is None else int(math_mode), **fields)
        eng = cls(cfg, device)
        for name, tensor in weights.items():
            eng.set_weight(name, tensor)
        eng.finalize()
        return eng

    def set_weight(self, name: str, tensor: torch.Tensor) -> None:
        t = tensor.detach().to(device="cpu", dtype=torch.float32).contiguous()
        check(self.lib.fd_set_weight(self._h, name.encode(), _ptr(t), t.numel()))

    def finalize(self) -> None:
        check(self.lib.fd_finalize_weights(self._h))

    @property
    def active_path(self) -> str:
        return {0: "generic-fp32", 1: "tf32-tensor-core", 2: "lstm-f16-warp-mma"}[self.lib.fd_active_path(self._h)]

    @property
    def launch_count(self) -> int:
        return int(self.lib.fd_launch_count(self._h))

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def _check_x(self, x: torch.Tensor) -> None:
        assert tuple(x.shape[1:]) == (self.L, self.C), (
            f"X has wrong shape, should be {(x.shape[0], self.L, self.C)}, but is {tuple(x.shape)}"
        )  # score_models.py:69-72

    # ---- per-phase entry points ---------------------------------------------------------------------------------------
    def score(self, x: torch.Tensor, t: float) -> torch.Tensor:
        self._check_x(x)
        xd = self._dev(x)
        out = torch.empty_like(xd)
        with torch.cuda.device(self.device):
            check(self.lib.fd_score(self._h, _ptr(xd), float(t), _ptr(out), xd.shape[0], _stream_ptr(self.device)))
        return out

    def _dev_t(self, timesteps: torch.Tensor, batch: int) -> torch.Tensor:
        td = self._dev(timesteps).reshape(-1)
        assert td.numel() == batch, f"timesteps has {td.numel()} entries for a batch of {batch}"
        return td

    def score_t(self, x: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """Score with one diffusion time per series (training / validation style batches, losses.py:58-62)."""
        self._check_x(x)
        xd = self._dev(x)
        td = self._dev_t(timesteps, xd.shape[0])
        out = torch.empty_like(xd)
        with torch.cuda.device(self.device):
            check(self.lib.fd_score_t(self._h, _ptr(xd), _ptr(td), _ptr(out), xd.shape[0], _stream_ptr(self.device)))
        return out

    def perturb(self, x0: torch.Tensor, timesteps: torch.Tensor, z: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """(mean(x0, t) + diag(std(t)) z, per-series std scalar): marginal_prob + add_noise as the loss uses them (losses.py:67-84)."""
        self._check_x(x0)
        xd, zd = self._dev(x0), self._dev(z)
        assert zd.shape == xd.shape
        td = self._dev_t(timesteps, xd.shape[0])
        out = torch.empty_like(xd)
        std_scalar = torch.empty(xd.shape[0], device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.fd_perturb(self._h, _ptr(xd), _ptr(td), _ptr(zd), _ptr(out), _ptr(std_scalar), xd.shape[0], _stream_ptr(self.device)))
        return out, std_scalar

    def sde_loss(self, x0: torch.Tensor, timesteps: torch.Tensor, z: torch.Tensor, likelihood_weighting: bool = False,
                 reduce_mean: bool = True) -> tuple[torch.Tensor, torch.Tensor]:
        """(loss, per-series losses) of get_sde_loss_fn(train=False) for supplied times and normals (losses.py:39-125), all on the device."""
        self._check_x(x0)
        xd, zd = self._dev(x0), self._dev(z)
        assert zd.shape == xd.shape
        td = self._dev_t(timesteps, xd.shape[0])
        losses = torch.empty(xd.shape[0], device=self.device, dtype=torch.float32)
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.fd_sde_loss(self._h, _ptr(xd), _ptr(td), _ptr(zd), int(bool(likelihood_weighting)), int(bool(reduce_mean)),
                                       _ptr(losses), _ptr(loss), xd.shape[0], _stream_ptr(self.device)))
        return loss, losses

    def step(self, x: torch.Tensor, score: torch.Tensor, z: torch.Tensor, t: float, step_size: float) -> torch.Tensor:
        self._check_x(x)
        xd, sd_, zd = self._dev(x), self._dev(score), self._dev(z)
        out = torch.empty_like(xd)
        with torch.cuda.device(self.device):
            check(self.lib.fd_step(self._h, _ptr(xd), _ptr(sd_), _ptr(zd), float(t), float(step_size), _ptr(out), xd.shape[0],
                                   _stream_ptr(self.device)))
        return out

    def prior(self, z: torch.Tensor) -> torch.Tensor:
        self._check_x(z)
        zd = self._dev(z)
        out = torch.empty_like(zd)
        with torch.cuda.device(self.device):
            check(self.lib.fd_prior(self._h, _ptr(zd), _ptr(out), zd.shape[0], _stream_ptr(self.device)))
        return out

    def ffn_block(self, layer: int, h: torch.Tensor) -> torch.Tensor:
        """LN2(h + FFN(h)) of encoder layer `layer` on (n_tokens, d_model) activations (per-phase parity entry point)."""
        hd = self._dev(h).clone()
        assert hd.dim() == 2 and hd.shape[1] == self.D
        with torch.cuda.device(self.device):
            check(self.lib.fd_ffn_block(self._h, layer, _ptr(hd), hd.shape[0], _stream_ptr(self.device)))
        return hd

    def attention_block(self, layer: int, h: torch.Tensor) -> torch.Tensor:
        """LN1(h + out_proj(MHA(h))) of encoder layer `layer` on (batch, max_len, d_model) activations (per-phase parity entry point)."""
        hd = self._dev(h).clone()
        assert hd.dim() == 3 and tuple(hd.shape[1:]) == (self.L, self.D)
        with torch.cuda.device(self.device):
            check(self.lib.fd_attention_block(self._h, layer, _ptr(hd), hd.shape[0], _stream_ptr(self.device)))
        return hd

    def encoder_stack(self, h: torch.Tensor) -> torch.Tensor:
        """backbone(h): every encoder layer on (batch, max_len, d_model) activations (score_models.py:87; per-phase parity entry point)."""
        hd = self._dev(h).clone()
        assert hd.dim() == 3 and tuple(hd.shape[1:]) == (self.L, self.D)
        with torch.cuda.device(self.device):
            check(self.lib.fd_encoder_stack(self._h, _ptr(hd), hd.shape[0], _stream_ptr(self.device)))
        return hd

    def normal(self, batch: int, seed: int, first_series: int = 0, draw: int = 0) -> torch.Tensor:
        out = torch.empty(batch, self.L, self.C, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.fd_normal(self._h, seed, first_series, draw, _ptr(out), batch, _stream_ptr(self.device)))
        return out

    # ---- the hot loop ----------------------------------------------------------------------------------------------
    def sample(self, batch: int, timesteps: torch.Tensor, step_size: float, seed: int = 0, first_series: int = 0,
               prior_z: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None, n_run: Optional[int] = None) -> torch.Tensor:
        """Device-resident variant: returns a CUDA tensor (batch, L, C); asynchronous on the current stream."""
        ts = timesteps.detach().to(device="cpu", dtype=torch.float32).contiguous()
        n_run = ts.numel() if n_run is None else int(n_run)
        pz = None if prior_z is None else self._dev(prior_z)
        nz = None if noise is None else self._dev(noise)
        if pz is not None:
            assert tuple(pz.shape) == (batch, self.L, self.C), f"prior_z has shape {tuple(pz.shape)}, expected {(batch, self.L, self.C)}"
        if nz is not None:
            assert nz.dim() == 4 and nz.shape[0] >= n_run and tuple(nz.shape[1:]) == (batch, self.L, self.C), (
                f"noise has shape {tuple(nz.shape)}, expected (>= {n_run}, {batch}, {self.L}, {self.C})")
        out = torch.empty(batch, self.L, self.C, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.fd_sample(self._h, batch, n_run, _ptr(ts), float(step_size), seed, first_series, _ptr(pz), _ptr(nz),
                                     _ptr(out), _stream_ptr(self.device)))
        self._keepalive = (ts, pz, nz)
        return out

    def sample_host(self, batch: int, timesteps: torch.Tensor, step_size: float, seed: int = 0, first_series: int = 0,
                    prior_z: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None, n_run: Optional[int] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end variant with HOST buffers (the `X.cpu()` of sampler.py:107 included); returns a CPU tensor."""
        ts = timesteps.detach().to(device="cpu", dtype=torch.float32).contiguous()
        n_run = ts.numel() if n_run is None else int(n_run)
        pz = None if prior_z is None else prior_z.detach().to(device="cpu", dtype=torch.float32).contiguous()
        nz = None if noise is None else noise.detach().to(device="cpu", dtype=torch.float32).contiguous()
        # fd_sample_host copies batch*L*C (prior) and n_run*batch*L*C (noise) floats from these buffers: check before handing out raw pointers
        if pz is not None:
            assert tuple(pz.shape) == (batch, self.L, self.C), f"prior_z has shape {tuple(pz.shape)}, expected {(batch, self.L, self.C)}"
        if nz is not None:
            assert nz.dim() == 4 and nz.shape[0] >= n_run and tuple(nz.shape[1:]) == (batch, self.L, self.C), (
                f"noise has shape {tuple(nz.shape)}, expected (>= {n_run}, {batch}, {self.L}, {self.C})")
            nz = nz[:n_run].contiguous()
        if out is None:
            out = torch.empty(batch, self.L, self.C, dtype=torch.float32, pin_memory=True)
        assert out.device.type == "cpu" and out.is_contiguous() and tuple(out.shape) == (batch, self.L, self.C)
        with torch.cuda.device(self.device):
            check(self.lib.fd_sample_host(self._h, batch, n_run, _ptr(ts), float(step_size), seed, first_series, _ptr(pz), _ptr(nz),
                                          _ptr(out), _stream_ptr(self.device)))
        return out

    def set_option(self, name: str, value: int) -> None:
        """Tuning knobs of the handle, e.g. ("attn_bounded_softmax", 0) forces the exact two-pass softmax (include/fdiff_b200.h)."""
        check(self.lib.fd_set_option(self._h, name.encode(), int(value)))

    def stack_stats(self):
        """Per-CTA cycle counters of the persistent encoder-stack kernel accumulated since the last call (option "stack_debug" must be
        on): int64 array (n_ctas, 64), see include/fdiff_b200.h::fd_debug_stack_stats.  Synchronises the device."""
        import numpy as np

        buf = np.zeros((1024, 64), dtype=np.int64)
        n = self.lib.fd_debug_stack_stats(self._h, C.c_void_p(buf.ctypes.data), 1024)
        return buf[:n]

    # ---- profiling -------------------------------------------------------------------------------------------------
    def profile_enable(self, every_n_steps: int) -> None:
        check(self.lib.fd_profile_enable(self._h, int(every_n_steps)))

    def profile(self, family: str) -> tuple[float, int]:
        return float(self.lib.fd_profile_ms(self._h, family.encode())), int(self.lib.fd_profile_launches(self._h, family.encode()))
