"""Case definitions shared by `make_golden.py` (generator; needs /root/reference) and the parity tests.

Every case is reproducible from seeds alone: weights come from constructing the score module under
`torch.manual_seed(weight_seed)` (CPU mt19937 — identical on every box), inputs/noise from
`torch.Generator().manual_seed(noise_seed)`.  The golden files therefore hold only outputs and fp64 weight
checksums (to prove the host mirror built the same weights as the reference did).
"""
from __future__ import annotations

import torch

WEIGHT_SEED = 42  # cmd/conf/train.yaml:1, cmd/conf/sample.yaml:5
NOISE_SEED = 1234

# name -> dict(model=..., L, C, kwargs, scheduler, fourier)
SCORE_CASES = {
    "tiny_vp": dict(model="transformer", L=20, C=3, B=3, kw=dict(d_model=8, n_head=4, num_layers=2), sched="vp", fourier=False),
    "classdefault_ve": dict(model="transformer", L=50, C=3, B=2, kw=dict(), sched="ve", fourier=True),
    "cfg2_vp": dict(model="transformer", L=256, C=12, B=2, kw=dict(d_model=72, n_head=12, num_layers=10), sched="vp", fourier=True),
    "nasdaq_vp": dict(model="transformer", L=252, C=5, B=2, kw=dict(d_model=72, n_head=12, num_layers=10), sched="vp", fourier=True),
    "ecg_vp": dict(model="transformer", L=187, C=1, B=3, kw=dict(d_model=72, n_head=12, num_layers=10), sched="vp", fourier=True),
    # US-Droughts length (datamodules.py:530-532): beyond the fused attention kernel's 256-key tile -> streaming attention kernels
    "droughts_vp": dict(model="transformer", L=365, C=7, B=2, kw=dict(d_model=72, n_head=12, num_layers=10), sched="vp", fourier=True),
    "mimic_lstm_vp": dict(model="lstm", L=24, C=40, B=4, kw=dict(d_model=72, num_layers=10), sched="vp", fourier=True),
    "lstm_small_ve": dict(model="lstm", L=30, C=4, B=2, kw=dict(d_model=16, num_layers=2), sched="ve", fourier=False),
    "mlp_vp": dict(model="mlp", L=20, C=3, B=3, kw=dict(d_model=72, d_mlp=128, num_layers=3), sched="vp", fourier=True),
}
SCORE_TIMES = (1.0, 0.37, 1e-5)

# trajectories: (case name, num_diffusion_steps grid, steps actually run)
TRAJ_CASES = {
    "tiny_vp": (10, 10),
    "classdefault_ve": (10, 10),
    "cfg2_vp": (1000, 50),
    "ecg_vp": (50, 50),
    "droughts_vp": (1000, 10),
    "mimic_lstm_vp": (1000, 20),
    "mlp_vp": (10, 10),
}

# full-length drift check on the headline config: cfg2_vp, 1000 of 1000 steps, states kept after these many steps
LONG_TRAJ_CASE, LONG_TRAJ_STEPS, LONG_TRAJ_MARKS = "cfg2_vp", 1000, (50, 200, 500, 1000)


def long_traj_noise():
    """(prior_z, noise[1000]) of the 1000-step trajectory; 24.6 MB, regenerated from the seed instead of being stored."""
    c = SCORE_CASES[LONG_TRAJ_CASE]
    g = torch.Generator().manual_seed(NOISE_SEED + 7)
    prior_z = torch.randn(c["B"], c["L"], c["C"], generator=g)
    noise = torch.randn(LONG_TRAJ_STEPS, c["B"], c["L"], c["C"], generator=g)
    return prior_z, noise


# evaluation loss (losses.py:39-125): score cases run through get_sde_loss_fn(train=False) with seeded per-series times and normals
LOSS_CASES = ("tiny_vp", "classdefault_ve", "cfg2_vp", "droughts_vp", "mimic_lstm_vp", "lstm_small_ve", "mlp_vp")


def loss_inputs(name: str):
    """(x0, t, z): data, per-series diffusion times in [eps, 1) and standard normals of a loss case."""
    c = SCORE_CASES[name]
    g = torch.Generator().manual_seed(NOISE_SEED + 11)
    x0 = torch.randn(c["B"], c["L"], c["C"], generator=g)
    t = torch.rand(c["B"], generator=g) * (1.0 - 1e-5) + 1e-5
    z = torch.randn(c["B"], c["L"], c["C"], generator=g)
    return x0, t, z


DFT_LENGTHS = (8, 7, 24, 100, 101, 187, 251, 252, 256, 365, 1024, 4096)
DFT_B, DFT_C = 3, 2

SCHED_KW = {
    "vp": dict(beta_min=0.1, beta_max=20.0),  # cmd/conf/score_model/noise_scheduler/vpsde.yaml:2-5
    "ve": dict(sigma_min=0.01, sigma_max=2.0),  # .../vesde.yaml:2-5
}


def case_inputs(name: str):
    """(x, prior_z) for a score case, from the shared noise seed."""
    c = SCORE_CASES[name]
    g = torch.Generator().manual_seed(NOISE_SEED)
    x = torch.randn(c["B"], c["L"], c["C"], generator=g)
    return x


def traj_noise(name: str):
    c = SCORE_CASES[name]
    grid, run = TRAJ_CASES[name]
    g = torch.Generator().manual_seed(NOISE_SEED + 1)
    prior_z = torch.randn(c["B"], c["L"], c["C"], generator=g)
    noise = torch.randn(run, c["B"], c["L"], c["C"], generator=g)
    return prior_z, noise


def dft_input(L: int):
    g = torch.Generator().manual_seed(NOISE_SEED + L)
    return torch.randn(DFT_B, L, DFT_C, generator=g)


def weight_checksums(state_dict) -> dict:
    return {k: float(v.detach().double().sum()) for k, v in state_dict.items()}
