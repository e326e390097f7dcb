"""CPU: the oracle (oracle/fdiff_oracle.py) against the golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden.py), and the host mirror's weights against the reference's weight checksums."""
from __future__ import annotations

import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, build_mirror_model, cases, load_golden, rel_err
from oracle import fdiff_oracle as O

SCORE_TOL = 5e-6  # fp32 restatement vs torch's fused encoder fast path (measured <= 8e-7; headroom for MKL thread counts)


@pytest.mark.parametrize("name", list(cases.SCORE_CASES))
def test_mirror_weights_equal_reference(name):
    from fourierdiffusion_b200.engine import renorm_fixed_point

    m, _ = build_mirror_model(name)
    g = load_golden(name)
    sd = m.state_dict()
    assert set(sd) == set(g["checksums"])
    for k, want in g["checksums"].items():
        t = sd[k]
        if k == "pos_encoder.embedding.weight":  # the reference renormalises this table in place on every forward
            t = renorm_fixed_point(t, math.sqrt(m.d_model))
        assert float(t.double().sum()) == pytest.approx(want, rel=1e-12, abs=1e-12), k


@pytest.mark.parametrize("name", list(cases.SCORE_CASES))
def test_oracle_score_step_prior(name):
    m, sch = build_mirror_model(name)
    g = load_golden(name)
    spec = O.model_spec_from_module(m)
    sspec = O.scheduler_spec_from_object(sch)
    c = cases.SCORE_CASES[name]
    x = cases.case_inputs(name)
    with torch.no_grad():
        for i, t in enumerate(cases.SCORE_TIMES):
            s = O.score(spec, x, torch.full((c["B"],), t, dtype=torch.float32))
            assert rel_err(s, g[f"score_{i}"]) < SCORE_TOL, (name, t)
            if spec.kind == "transformer":  # the timed-baseline variant (ATen fused encoder layer, what the reference dispatches to)
                s2 = O.score(spec, x, torch.full((c["B"],), t, dtype=torch.float32), aten_layers=True)
                assert rel_err(s2, g[f"score_{i}"]) < SCORE_TOL, (name, t)
        # scheduler step and prior: bit-exact
        G = O.g_vector(c["L"], sspec.fourier_noise_scaling)
        ts, dt = O.make_timesteps(1000, sspec.eps)
        gen = torch.Generator().manual_seed(cases.NOISE_SEED + 2)
        z = torch.randn(*x.shape, generator=gen)
        out = O.scheduler_step(sspec, x, torch.from_numpy(g["score_1"]), z, 0.5, G, dt)
        assert np.array_equal(out.numpy(), g["step_t0.5"]), name
        pr = O.prior_from_noise(z, G, sspec.sigma_max if sspec.kind == "ve" else None)
        assert np.array_equal(pr.numpy(), g["prior"]), name


@pytest.mark.parametrize("name", list(cases.TRAJ_CASES))
def test_oracle_trajectory(name):
    m, sch = build_mirror_model(name)
    g = load_golden(name)
    grid, run = cases.TRAJ_CASES[name]
    prior_z, noise = cases.traj_noise(name)
    with torch.no_grad():
        traj = O.sample_trajectory(O.model_spec_from_module(m), O.scheduler_spec_from_object(sch), prior_z, noise, grid, run)
    assert rel_err(traj, g["traj"]) < 2e-5, name


@pytest.mark.parametrize("L", cases.DFT_LENGTHS)
def test_oracle_dft_idft(L):
    g = np.load(os.path.join(GOLDEN, "fourier.npz"))
    x = cases.dft_input(L)
    assert rel_err(O.dft(x), g[f"dft_{L}"]) < 2e-6
    assert rel_err(O.idft(x), g[f"idft_{L}"]) < 2e-6
    # the reference's own known-answer test: idft(dft(x)) = x and dft(idft(x)) = x, atol 1e-5 (tests/test_utils.py:37-51)
    assert torch.allclose(O.idft(O.dft(x)), x, atol=1e-5)
    assert torch.allclose(O.dft(O.idft(x)), x, atol=1e-5)
    # spectral_density (fourier.py:90-124), from the time series and from the packed spectrum
    assert O.spectral_density(x).shape == (cases.DFT_B, L // 2 + 1, cases.DFT_C)
    assert rel_err(O.spectral_density(x), g[f"spec_{L}"]) < 5e-6
    assert rel_err(O.spectral_density(O.dft(x), apply_dft=False), g[f"specpacked_{L}"]) < 5e-6


def test_returned_sample_count_rule():
    # sampler.py:63,75-78: max(1, n // bs) batches, remainder dropped
    assert O.num_returned_samples(48, 12) == 48
    assert O.num_returned_samples(10, 4) == 8
    assert O.num_returned_samples(3, 4) == 3
    assert O.num_returned_samples(10000, 200) == 10000


def test_oracle_long_trajectory_first_200_steps():
    """The 1000-step golden trajectory of the headline configuration (tests/golden/traj1000_cfg2.npz, reference sampler with injected
    noise): the oracle restatement is checked over its first 200 steps here (the CUDA path runs all 1000 in tests/test_gpu_parity.py)."""
    name = cases.LONG_TRAJ_CASE
    m, sch = build_mirror_model(name)
    g = np.load(os.path.join(GOLDEN, "traj1000_cfg2.npz"))
    prior_z, noise = cases.long_traj_noise()
    with torch.no_grad():
        traj = O.sample_trajectory(O.model_spec_from_module(m), O.scheduler_spec_from_object(sch), prior_z, noise[:200], cases.LONG_TRAJ_STEPS, 200)
    assert rel_err(traj, g["x_200"]) < 2e-5
