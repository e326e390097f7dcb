"""CPU oracle for the fdiff sampling hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement (elementary fp32 tensor ops on the CPU) of the reference algorithm that the CUDA path
in `fourierdiffusion_b200/csrc/` replaces.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module, and only as the checker / the timed CPU baseline.
Nothing under `fourierdiffusion_b200/` imports it: the product path fails loudly without its CUDA library.

Parity status: PINNED.  `tests/test_oracle_pinned.py` checks every function below against the UNMODIFIED
reference (imported from /root/reference through `oracle/ref_loader.py`) when that tree is present, and
`tests/golden/*.npz` (generated from the reference by `tests/golden/make_golden.py`) pins it everywhere else.
The arithmetic itself lives in the third-party dependency `torch` (unpinned in the reference's pyproject.toml:38;
this image: 2.11.0+cu128, CPU/MKL); its call sites are cited per function below.

All citations are relative to /root/reference/ (commit e60d532c).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# Scheduler side (src/fdiff/schedulers/sde.py)
# --------------------------------------------------------------------------------------------------------------
def g_vector(max_len: int, fourier_noise_scaling: bool) -> Tensor:
    """Diagonal of the diffusion scaling G.  sde.py:42-60 (`SDE.set_noise_scaling`)."""
    G = torch.ones(max_len)
    if fourier_noise_scaling:
        G = 1 / (math.sqrt(2)) * G
        G[0] *= math.sqrt(2)
        if max_len % 2 == 0:
            G[max_len // 2] *= math.sqrt(2)
    return G


def make_timesteps(num_diffusion_steps: int, eps: float = 1e-5) -> tuple[Tensor, Tensor]:
    """`linspace(1, eps, N)` fp32 and the constant step size t0 - t1 (0-dim fp32).  sde.py:62-64."""
    ts = torch.linspace(1.0, eps, num_diffusion_steps)
    return ts, ts[0] - ts[1]


def prior_from_noise(z: Tensor, G: Tensor, sigma_max: Optional[float] = None) -> Tensor:
    """x_T = G ⊙ z (dense `G_matrix @ z` in the reference, sde.py:79-87); VE multiplies by sigma_max (sde.py:125-127)."""
    x = G.view(1, -1, 1) * z
    if sigma_max is not None:
        x = sigma_max * x
    return x


def vp_beta(t: float, beta_0: float, beta_1: float) -> float:
    """sde.py:212-213 (`get_beta`), python float64."""
    return beta_0 + t * (beta_1 - beta_0)


def ve_sqrt_derivative(t: float, sigma_min: float, sigma_max: float) -> float:
    """sde.py:142-146, python float64."""
    return sigma_min * math.sqrt(2 * math.log(sigma_max / sigma_min)) * (sigma_max / sigma_min) ** t


def vp_step(x: Tensor, score: Tensor, z: Tensor, t: float, G: Tensor, dt: Tensor, beta_0: float, beta_1: float) -> Tensor:
    """One reverse VP-SDE step with the noise `z` supplied by the caller.  sde.py:215-246.

    Elementwise closed form of the reference's dense diag-matmul expression, evaluated in the same operation
    order (bit-identical on CPU, see tests/test_oracle_pinned.py):
        d_l   = fl32(sqrt(beta)) * G_l
        drift = fl32(-0.5*beta) * x - (d_l*d_l) * score
        x'    = x - drift*dt + sqrt(dt) * (d_l * z)
    """
    beta = vp_beta(t, beta_0, beta_1)
    d = (math.sqrt(beta) * G).view(1, -1, 1)
    drift = -0.5 * beta * x - (d * d) * score
    return x - drift * dt + torch.sqrt(dt) * (d * z)


def ve_step(x: Tensor, score: Tensor, z: Tensor, t: float, G: Tensor, dt: Tensor, sigma_min: float, sigma_max: float) -> Tensor:
    """One reverse VE-SDE step with supplied noise.  sde.py:129-165."""
    sd = ve_sqrt_derivative(t, sigma_min, sigma_max)
    d = (sd * G).view(1, -1, 1)
    drift = -((d * d) * score)
    return x - drift * dt + torch.sqrt(dt) * (d * z)


# --------------------------------------------------------------------------------------------------------------
# Embeddings (src/fdiff/models/transformer.py)
# --------------------------------------------------------------------------------------------------------------
def renorm_positional_table(E: Tensor, max_norm: float, max_passes: int = 16) -> Tensor:
    """Fixed point of `nn.Embedding(max_norm=sqrt(D))`'s in-place renorm (transformer.py:13-15,27).

    torch's `embedding_renorm_` rescales every looked-up row with ||row||_2 > max_norm by max_norm/(norm+1e-7) on
    EVERY forward; after <=3 passes no row moves any more.  The reference mutates its weights this way, so the
    table the score network really uses is this fixed point.
    """
    E = E.clone()
    for _ in range(max_passes):
        norms = E.norm(p=2, dim=1)
        mask = norms > max_norm
        if not bool(mask.any()):
            break
        scale = torch.where(mask, max_norm / (norms + 1e-7), torch.ones_like(norms))
        E = E * scale[:, None]
    return E


def time_embedding(t: Tensor, W: Tensor, dense_w: Tensor, dense_b: Tensor, d_model: int) -> Tensor:
    """GaussianFourierProjection: dense(cat(sin, cos)(t*W*2*pi)[:D]).  transformer.py:77-91.  t: (B,) -> (B, D)."""
    proj = t[:, None] * W[None, :] * 2 * np.pi
    emb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)[:, :d_model]
    return emb @ dense_w.t() + dense_b


# --------------------------------------------------------------------------------------------------------------
# Score networks (src/fdiff/models/score_models.py)
# --------------------------------------------------------------------------------------------------------------
def _layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def transformer_score(sd: Dict[str, Tensor], x: Tensor, t: Tensor, n_head: int, pos_table: Optional[Tensor] = None) -> Tensor:
    """`ScoreModule.forward` (score_models.py:67-94) in eval mode, written out.

    embedder (:78) -> + positional table (:81, transformer.py:26-28) -> + time encoding (:84) ->
    num_layers x nn.TransformerEncoderLayer (post-LN, ReLU, ff=2048, eps=1e-5; :57-62,87) -> unembedder (:90).
    Packed in_proj rows are [q | k | v]; head j owns rows j*dh:(j+1)*dh of each block; q is scaled by 1/sqrt(dh).
    `pos_table` must be the renormalised table (see `renorm_positional_table`); default: taken from `sd` as is.
    """
    B, L, C = x.shape
    D = sd["embedder.weight"].shape[0]
    dh = D // n_head
    E = sd["pos_encoder.embedding.weight"] if pos_table is None else pos_table
    temb = time_embedding(t, sd["time_encoder.W"], sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"], D)
    h = x @ sd["embedder.weight"].t() + sd["embedder.bias"]
    h = h + E[:L][None]
    h = h + temb[:, None, :]
    i = 0
    while f"backbone.layers.{i}.linear1.weight" in sd:
        p = f"backbone.layers.{i}."
        qkv = h @ sd[p + "self_attn.in_proj_weight"].t() + sd[p + "self_attn.in_proj_bias"]
        q, k, v = qkv.split(D, dim=-1)
        q = q.view(B, L, n_head, dh).transpose(1, 2) / math.sqrt(dh)
        k = k.view(B, L, n_head, dh).transpose(1, 2)
        v = v.view(B, L, n_head, dh).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        s = s - s.max(dim=-1, keepdim=True).values
        pr = torch.exp(s)
        pr = pr / pr.sum(dim=-1, keepdim=True)
        o = (pr @ v).transpose(1, 2).reshape(B, L, D)
        o = o @ sd[p + "self_attn.out_proj.weight"].t() + sd[p + "self_attn.out_proj.bias"]
        h = _layer_norm(h + o, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        f = torch.relu(h @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
        f = f @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"]
        h = _layer_norm(h + f, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        i += 1
    return h @ sd["unembedder.weight"].t() + sd["unembedder.bias"]


def transformer_score_aten(sd: Dict[str, Tensor], x: Tensor, t: Tensor, n_head: int, pos_table: Optional[Tensor] = None) -> Tensor:
    """Same function as `transformer_score`, but each encoder layer is evaluated by the very ATen operator the reference
    dispatches to in eval + no_grad (`aten::_transformer_encoder_layer_fwd`, the nn.TransformerEncoderLayer fast path reached
    from score_models.py:87).  This is what the reference's CPU arithmetic costs, so it is the variant that
    `bench.py` TIMES as the CPU baseline; the written-out `transformer_score` above is the variant parity is checked with.
    Both are pinned to the same golden vectors (tests/test_oracle_golden.py)."""
    B, L, C = x.shape
    D = sd["embedder.weight"].shape[0]
    E = sd["pos_encoder.embedding.weight"] if pos_table is None else pos_table
    temb = time_embedding(t, sd["time_encoder.W"], sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"], D)
    h = x @ sd["embedder.weight"].t() + sd["embedder.bias"]
    h = h + E[:L][None]
    h = h + temb[:, None, :]
    i = 0
    while f"backbone.layers.{i}.linear1.weight" in sd:
        p = f"backbone.layers.{i}."
        h = torch._transformer_encoder_layer_fwd(
            h, D, n_head, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"], sd[p + "self_attn.out_proj.weight"],
            sd[p + "self_attn.out_proj.bias"], False, False, 1e-5, sd[p + "norm1.weight"], sd[p + "norm1.bias"], sd[p + "norm2.weight"],
            sd[p + "norm2.bias"], sd[p + "linear1.weight"], sd[p + "linear1.bias"], sd[p + "linear2.weight"], sd[p + "linear2.bias"],
            None, None)
        i += 1
    return h @ sd["unembedder.weight"].t() + sd["unembedder.bias"]


def lstm_score(sd: Dict[str, Tensor], x: Tensor, t: Tensor) -> Tensor:
    """`LSTMScoreModule.forward` (score_models.py:292-317): embedder -> + time enc (no positional table, :287) ->
    layers x (u <- u + LSTM_i(u)), each a single-layer nn.LSTM with zero initial state, gate order i,f,g,o -> unembedder."""
    B, L, C = x.shape
    D = sd["embedder.weight"].shape[0]
    temb = time_embedding(t, sd["time_encoder.W"], sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"], D)
    u = x @ sd["embedder.weight"].t() + sd["embedder.bias"] + temb[:, None, :]
    i = 0
    while f"backbone.{i}.weight_ih_l0" in sd:
        p = f"backbone.{i}."
        wih, whh = sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"]
        bih, bhh = sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"]
        hs = torch.zeros(B, D)
        cs = torch.zeros(B, D)
        xin = u @ wih.t() + bih
        ys = []
        for s in range(L):
            g = xin[:, s] + hs @ whh.t() + bhh
            gi, gf, gg, go = g.split(D, dim=-1)
            cs = torch.sigmoid(gf) * cs + torch.sigmoid(gi) * torch.tanh(gg)
            hs = torch.sigmoid(go) * torch.tanh(cs)
            ys.append(hs)
        u = u + torch.stack(ys, dim=1)
        i += 1
    return u @ sd["unembedder.weight"].t() + sd["unembedder.bias"]


def mlp_score(sd: Dict[str, Tensor], x: Tensor, t: Tensor) -> Tensor:
    """`MLPScoreModule.forward` (score_models.py:215-246), eval mode: flatten `b t c -> b (t c)`, embedder, + time enc,
    layers x (u <- u + W2 relu(W1 u + b1) + b2) [torchvision.ops.MLP: keys `backbone.{i}.0` / `.3`], unembedder, unflatten."""
    B, L, C = x.shape
    D = sd["embedder.weight"].shape[0]
    temb = time_embedding(t, sd["time_encoder.W"], sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"], D)
    u = x.reshape(B, L * C) @ sd["embedder.weight"].t() + sd["embedder.bias"] + temb
    i = 0
    while f"backbone.{i}.0.weight" in sd:
        p = f"backbone.{i}."
        f = torch.relu(u @ sd[p + "0.weight"].t() + sd[p + "0.bias"])
        u = u + (f @ sd[p + "3.weight"].t() + sd[p + "3.bias"])
        i += 1
    out = u @ sd["unembedder.weight"].t() + sd["unembedder.bias"]
    return out.reshape(B, L, C)


# --------------------------------------------------------------------------------------------------------------
# Fourier utilities (src/fdiff/utils/fourier.py)
# --------------------------------------------------------------------------------------------------------------
def dft(x: Tensor) -> Tensor:
    """Ortho rFFT along dim 1, packed real: [Re X_0..X_{L//2} | Im X_1..X_{ceil(L/2)-1}] (length L).  fourier.py:8-45."""
    L = x.shape[1]
    X = np.fft.rfft(x.detach().cpu().numpy().astype(np.float64), axis=1, norm="ortho")
    re = X.real
    im = X.imag[:, 1:]
    if L % 2 == 0:
        im = im[:, :-1]
    return torch.from_numpy(np.concatenate([re, im], axis=1).astype(np.float32))


def idft(x: Tensor) -> Tensor:
    """Inverse of `dft`: rebuild the half spectrum (Im X_0 = 0, Im X_{L/2} = 0 for even L), ortho irFFT.  fourier.py:48-87."""
    L = x.shape[1]
    n_real = math.ceil((L + 1) / 2)
    a = x.detach().cpu().numpy().astype(np.float64)
    re = a[:, :n_real]
    im = np.zeros_like(re)
    n_im = a.shape[1] - n_real
    im[:, 1 : 1 + n_im] = a[:, n_real:]
    out = np.fft.irfft(re + 1j * im, n=L, axis=1, norm="ortho")
    return torch.from_numpy(out.astype(np.float32))


# --------------------------------------------------------------------------------------------------------------
# Sampler (src/fdiff/sampling/sampler.py)
# --------------------------------------------------------------------------------------------------------------
def spectral_density(x: Tensor, apply_dft: bool = True) -> Tensor:
    """fourier.py:90-124: squared modulus of the ortho rFFT bins k = 0 .. L // 2 from the packed layout (Im X_0 and, for even L,
    Im X_{L/2} are zero and not stored) -> (batch, L // 2 + 1, n_channels)."""
    L = x.shape[1]
    p = dft(x) if apply_dft else x
    n_real = L // 2 + 1  # == ceil((L + 1) / 2), fourier.py:104
    re = p[:, :n_real, :]
    im = torch.zeros_like(re)
    n_imag = L - n_real
    im[:, 1 : 1 + n_imag, :] = p[:, n_real:, :]
    return re**2 + im**2


@dataclass
class SchedulerSpec:
    """The scalar state of an fdiff scheduler that the path reads (sde.py:17-24,93-106,171-185)."""

    kind: str = "vp"  # "vp" | "ve"
    beta_0: float = 0.1
    beta_1: float = 20.0
    sigma_min: float = 0.01
    sigma_max: float = 50.0
    eps: float = 1e-5
    fourier_noise_scaling: bool = False


@dataclass
class ModelSpec:
    kind: str = "transformer"  # "transformer" | "lstm" | "mlp"
    n_head: int = 12
    sd: Dict[str, Tensor] = field(default_factory=dict)
    pos_table: Optional[Tensor] = None  # renormalised positional table (transformer only)


def score(model: ModelSpec, x: Tensor, t: Tensor, aten_layers: bool = False) -> Tensor:
    if model.kind == "transformer":
        fn = transformer_score_aten if aten_layers else transformer_score
        return fn(model.sd, x, t, model.n_head, model.pos_table)
    if model.kind == "lstm":
        return lstm_score(model.sd, x, t)
    if model.kind == "mlp":
        return mlp_score(model.sd, x, t)
    raise NotImplementedError(model.kind)


def scheduler_step(sch: SchedulerSpec, x: Tensor, s: Tensor, z: Tensor, t: float, G: Tensor, dt: Tensor) -> Tensor:
    if sch.kind == "vp":
        return vp_step(x, s, z, t, G, dt, sch.beta_0, sch.beta_1)
    if sch.kind == "ve":
        return ve_step(x, s, z, t, G, dt, sch.sigma_min, sch.sigma_max)
    raise NotImplementedError(sch.kind)


def marginal_prob(sch: SchedulerSpec, x: Tensor, t: Tensor, G: Tensor) -> tuple[Tensor, Tensor]:
    """(mean, std) of the perturbation kernel p(x_t | x_0) at per-series times t (B,): sde.py:108-123 (VE), :187-210 (VP)."""
    if sch.kind == "ve":
        sigma_min = torch.tensor(sch.sigma_min).type_as(t)
        sigma_max = torch.tensor(sch.sigma_max).type_as(t)
        std = (sigma_min * (sigma_max / sigma_min) ** t).view(-1, 1) * G
        return x, std
    if sch.kind == "vp":
        log_mean_coeff = -0.25 * t**2 * (sch.beta_1 - sch.beta_0) - 0.5 * t * sch.beta_0
        mean = torch.exp(log_mean_coeff[(...,) + (None,) * len(x.shape[1:])]) * x
        std = torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff.view(-1, 1))) * G
        return mean, std
    raise NotImplementedError(sch.kind)


def sde_loss(model: ModelSpec, sch: SchedulerSpec, x0: Tensor, t: Tensor, z: Tensor, G: Tensor, likelihood_weighting: bool = False,
             reduce_mean: bool = True) -> Dict[str, Tensor]:
    """Evaluation forward of get_sde_loss_fn for given times t (B,) and normals z (losses.py:39-125): returns the scalar loss, the
    per-series losses, the perturbed batch and the score.  The diag_embed matmuls of losses.py:70-79 are restated as row scalings
    (every other term of those contractions is an exact zero)."""
    mean, std = marginal_prob(sch, x0, t, G)  # losses.py:67, std (B, L)
    var = std**2
    noise = std.unsqueeze(-1) * z  # losses.py:73
    target_noise = (1 / std).unsqueeze(-1) * z  # losses.py:76-78
    x_noisy = mean + noise  # losses.py:82-84, sde.py:66-77
    s = score(model, x_noisy, t)  # losses.py:89
    if not likelihood_weighting:
        weighting_factor = 1.0 / torch.sum(1.0 / var, dim=1)  # losses.py:95
        losses = weighting_factor.view(-1, 1, 1) * torch.square(s + target_noise)  # losses.py:99-101
    else:
        losses = torch.square(std.unsqueeze(-1) * (s + target_noise))  # losses.py:110-117
    flat = losses.reshape(losses.shape[0], -1)
    per_series = torch.mean(flat, dim=-1) if reduce_mean else 0.5 * torch.sum(flat, dim=-1)  # losses.py:34-38
    return {"loss": torch.mean(per_series), "losses": per_series, "x_noisy": x_noisy, "score": s}


def sample_trajectory(
    model: ModelSpec,
    sch: SchedulerSpec,
    prior_z: Tensor,
    noise: Tensor,
    num_diffusion_steps: int,
    first_steps: Optional[int] = None,
) -> Tensor:
    """One batch of `DiffusionSampler.sample` (sampler.py:80-104) with injected randomness.

    prior_z: (B, L, C) standard normal draws for the prior (sde.py:85); noise: (n_steps, B, L, C) the per-step
    `randn_like` draws (sde.py:238).  `first_steps` truncates the loop (the timestep grid is still that of
    `num_diffusion_steps`), used for bounded CPU timing and short golden trajectories.
    """
    B, L, C = prior_z.shape
    G = g_vector(L, sch.fourier_noise_scaling)
    ts, dt = make_timesteps(num_diffusion_steps, sch.eps)
    x = prior_from_noise(prior_z, G, sch.sigma_max if sch.kind == "ve" else None)
    n = num_diffusion_steps if first_steps is None else first_steps
    for i in range(n):
        t = ts[i]
        tvec = torch.full((B,), t.item(), dtype=torch.float32)  # sampler.py:91-99
        s = score(model, x, tvec)
        x = scheduler_step(sch, x, s, noise[i], t.item(), G, dt)
    return x


def num_returned_samples(num_samples: int, sample_batch_size: int) -> int:
    """How many series `DiffusionSampler.sample` really returns: the remainder batch is dropped.  sampler.py:63,75-78."""
    num_batches = max(1, num_samples // sample_batch_size)
    total = 0
    for b in range(num_batches):
        total += min(num_samples - b * sample_batch_size, sample_batch_size)
    return total


# --------------------------------------------------------------------------------------------------------------
# Duck-typed extraction from a score module (reference `ScoreModule` or the repo's host mirror)
# --------------------------------------------------------------------------------------------------------------
def model_spec_from_module(module) -> ModelSpec:
    """Read weights from any object with the reference's attribute/`state_dict` layout."""
    sd = {k: v.detach().cpu().float().clone() for k, v in module.state_dict().items()}
    if any(k.startswith("backbone.layers.") for k in sd):
        n_head = module.backbone.layers[0].self_attn.num_heads
        D = sd["embedder.weight"].shape[0]
        pos = renorm_positional_table(sd["pos_encoder.embedding.weight"], math.sqrt(D))
        return ModelSpec(kind="transformer", n_head=n_head, sd=sd, pos_table=pos)
    if any(k.endswith("weight_ih_l0") for k in sd):
        return ModelSpec(kind="lstm", n_head=1, sd=sd)
    if "backbone.0.0.weight" in sd:
        return ModelSpec(kind="mlp", n_head=1, sd=sd)
    raise NotImplementedError("unrecognised score module layout")


def scheduler_spec_from_object(s) -> SchedulerSpec:
    """Read the scalars of a reference-style scheduler object (sde.py:93-106 VE, :171-185 VP)."""
    if hasattr(s, "beta_0"):
        return SchedulerSpec(kind="vp", beta_0=float(s.beta_0), beta_1=float(s.beta_1), eps=float(s.eps),
                             fourier_noise_scaling=bool(s.noise_scaling))
    if hasattr(s, "sigma_min"):
        return SchedulerSpec(kind="ve", sigma_min=float(s.sigma_min), sigma_max=float(s.sigma_max), eps=float(s.eps),
                             fourier_noise_scaling=bool(s.noise_scaling))
    raise NotImplementedError("Scheduler not recognized.")
