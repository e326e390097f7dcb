"""Evaluation loss (forward half): get_sde_loss_fn(train=False) of src/fdiff/utils/losses.py:39-125.

CPU: the oracle's restatement against the golden vectors of the UNMODIFIED reference (tests/golden/loss.npz, made by make_golden.py
through the reference's own loss factory) and, in the build container, against the reference imported live.
GPU: fd_perturb / fd_score_t / fd_sde_loss and the host mirror (`marginal_prob`, `add_noise`, `get_sde_loss_fn`, `validation_step`)
through the C ABI against both.

Tolerances (conftest.rel_err norm unless a scalar):
  * perturbed batch, mean, std                 : 2e-6 (expf / powf of the device vs the host libm, otherwise the same fp32 operations)
  * score at per-series times                  : the suite's score tolerances (2e-5 fp32 path, 2e-3 tensor-core path)
  * loss (scalar, relative)                    : 1e-4 fp32 path, 5e-3 tensor-core path (a mean of squares of score + target; measured
                                                 <= 3e-6 / <= 6e-4)
"""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, build_mirror_model, cases, rel_err
from oracle import fdiff_oracle as O
from oracle import ref_loader

FP32, TF32 = 0, 1
SCORE_TOL = {FP32: 2e-5, TF32: 2e-3}
LOSS_TOL = {FP32: 1e-4, TF32: 5e-3}
MODES = [(lw, rm) for lw in (False, True) for rm in (True, False)]


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "loss.npz"))
    return {k: g[k] for k in g.files}


def _oracle_specs(name):
    m, sch = build_mirror_model(name)
    return m, sch, O.model_spec_from_module(m), O.scheduler_spec_from_object(sch), O.g_vector(cases.SCORE_CASES[name]["L"], bool(sch.noise_scaling))


# ---- CPU: oracle pinned ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", cases.LOSS_CASES)
def test_oracle_loss_matches_golden(name, gold):
    m, sch, spec, sspec, G = _oracle_specs(name)
    x0, t, z = cases.loss_inputs(name)
    with torch.no_grad():
        mean, std = O.marginal_prob(sspec, x0, t, G)
        assert np.array_equal(mean.numpy(), gold[f"{name}_mean"]) and np.array_equal(std.numpy(), gold[f"{name}_std"])
        for lw, rm in MODES:
            r = O.sde_loss(spec, sspec, x0, t, z, G, likelihood_weighting=lw, reduce_mean=rm)
            want = float(gold[f"{name}_loss_lw{int(lw)}_rm{int(rm)}"])
            assert float(r["loss"]) == pytest.approx(want, rel=2e-5), (name, lw, rm)
        assert rel_err(r["x_noisy"], gold[f"{name}_x_noisy"]) < 1e-7
        assert rel_err(r["score"], gold[f"{name}_score"]) < 5e-6


@pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("name", ["tiny_vp", "classdefault_ve", "lstm_small_ve", "mlp_vp"])
def test_oracle_loss_against_reference_rng_path(name):
    """The reference draws the times (torch.rand) and the normals (torch.randn_like) itself: same seed, same draws, same loss."""
    R = ref_loader.load_reference()
    c = cases.SCORE_CASES[name]
    torch.manual_seed(cases.WEIGHT_SEED)
    Sched = {"vp": R.VPScheduler, "ve": R.VEScheduler}[c["sched"]]
    sch = Sched(fourier_noise_scaling=c["fourier"], **cases.SCHED_KW[c["sched"]])
    Model = {"transformer": R.ScoreModule, "lstm": R.LSTMScoreModule, "mlp": R.MLPScoreModule}[c["model"]]
    m = Model(n_channels=c["C"], max_len=c["L"], noise_scheduler=sch, fourier_noise_scaling=c["fourier"], **c["kw"]).eval()
    x0, _, _ = cases.loss_inputs(name)
    with torch.no_grad():
        for _ in range(3):
            m(R.DiffusableBatch(X=x0, y=None, timesteps=torch.full((c["B"],), 0.5)))
        torch.manual_seed(77)
        want = m.validation_loss_fn(m, R.DiffusableBatch(X=x0, y=None))
        torch.manual_seed(77)
        t = torch.rand(c["B"]) * (sch.T - sch.eps) + sch.eps
        z = torch.randn_like(x0)
        r = O.sde_loss(O.model_spec_from_module(m), O.scheduler_spec_from_object(sch), x0, t, z, O.g_vector(c["L"], c["fourier"]))
    assert float(r["loss"]) == pytest.approx(float(want), rel=2e-5)


def test_training_loss_is_refused():
    """No dropout / backward in the library: the mirror says so instead of returning a forward-only number."""
    import fourierdiffusion_b200 as fd

    m, sch = build_mirror_model("tiny_vp")
    x0, t, _ = cases.loss_inputs("tiny_vp")
    with pytest.raises(NotImplementedError):
        m.training_loss_fn(m, fd.DiffusableBatch(X=x0, y=None, timesteps=t))


# ---- GPU ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.LOSS_CASES)
def test_perturb_matches_golden(name, gold):
    m, sch = build_mirror_model(name)
    eng = m.engine(math_mode=FP32)
    x0, t, z = cases.loss_inputs(name)
    out, std_scalar = eng.perturb(x0, t, z)
    assert rel_err(out, gold[f"{name}_x_noisy"]) < 2e-6
    # the scheduler mirror's own methods (sde.py:66-77, :108-123, :187-210) on a scheduler-only handle
    mean, std = sch.marginal_prob(x0, t)
    assert tuple(std.shape) == (x0.shape[0], x0.shape[1])
    assert rel_err(mean, gold[f"{name}_mean"]) < 2e-6 and rel_err(std, gold[f"{name}_std"]) < 2e-6
    noise = std.unsqueeze(-1) * z
    assert rel_err(sch.add_noise(original_samples=x0, noise=noise, timesteps=t), gold[f"{name}_x_noisy"]) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [FP32, TF32])
@pytest.mark.parametrize("name", cases.LOSS_CASES)
def test_score_per_series_times_matches_golden(name, mode, gold):
    import fourierdiffusion_b200 as fd

    m, sch = build_mirror_model(name)
    m.math_mode = mode
    eng = m.engine(math_mode=mode)
    _, t, _ = cases.loss_inputs(name)
    xn = torch.from_numpy(gold[f"{name}_x_noisy"])
    s = eng.score_t(xn, t).cpu()
    assert rel_err(s, gold[f"{name}_score"]) < SCORE_TOL[mode], (name, eng.active_path)
    # the module's forward takes the same route for a batch whose times differ (score_models.py:67-94)
    s2 = m(fd.DiffusableBatch(X=xn, y=None, timesteps=t))
    assert torch.equal(s2.cpu(), s)
    # and a batch that shares one time gives what fd_score gives
    t1 = torch.full_like(t, 0.37)
    a, b = eng.score_t(xn, t1).cpu(), eng.score(xn, 0.37).cpu()
    assert rel_err(a, b) < (1e-6 if mode == FP32 else SCORE_TOL[mode])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [FP32, TF32])
@pytest.mark.parametrize("name", cases.LOSS_CASES)
def test_sde_loss_matches_golden(name, mode, gold):
    m, sch = build_mirror_model(name)
    eng = m.engine(math_mode=mode)
    x0, t, z = cases.loss_inputs(name)
    _, _, spec, sspec, G = _oracle_specs(name)
    for lw, rm in MODES:
        loss, losses = eng.sde_loss(x0, t, z, likelihood_weighting=lw, reduce_mean=rm)
        want = float(gold[f"{name}_loss_lw{int(lw)}_rm{int(rm)}"])
        assert float(loss) == pytest.approx(want, rel=LOSS_TOL[mode]), (name, lw, rm, eng.active_path)
        with torch.no_grad():
            r = O.sde_loss(spec, sspec, x0, t, z, G, likelihood_weighting=lw, reduce_mean=rm)
        assert rel_err(losses, r["losses"]) < LOSS_TOL[mode]
        assert float(loss) == pytest.approx(float(losses.double().mean()), rel=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_vp", "cfg2_vp", "mimic_lstm_vp"])
def test_validation_step_draws_like_the_reference(name, gold):
    """timesteps=None: the mirror consumes torch's generator in the reference's order (rand, then randn_like; losses.py:58-65)."""
    import fourierdiffusion_b200 as fd

    m, sch = build_mirror_model(name)
    m.math_mode = FP32
    x0, _, _ = cases.loss_inputs(name)
    torch.manual_seed(77)
    got = m.validation_step(fd.DiffusableBatch(X=x0, y=None), 0)
    torch.manual_seed(77)
    t = torch.rand(x0.shape[0]) * (sch.T - sch.eps) + sch.eps
    z = torch.randn_like(x0)
    _, _, spec, sspec, G = _oracle_specs(name)
    with torch.no_grad():
        want = O.sde_loss(spec, sspec, x0, t, z, G)["loss"]
    assert got.dim() == 0 and float(got) == pytest.approx(float(want), rel=LOSS_TOL[FP32])
    # the golden loss, with the batch carrying its times
    _, tg, zg = cases.loss_inputs(name)
    torch.manual_seed(5)
    z_drawn = torch.randn_like(x0)
    torch.manual_seed(5)
    got2 = m.validation_loss_fn(m, fd.DiffusableBatch(X=x0, y=None, timesteps=tg))
    with torch.no_grad():
        want2 = O.sde_loss(spec, sspec, x0, tg, z_drawn, G)["loss"]
    assert float(got2) == pytest.approx(float(want2), rel=LOSS_TOL[FP32])


@pytest.mark.gpu
def test_sde_loss_full_size_properties():
    """cfg 2 at its bench batch (256 series): per-series losses do not depend on the batch they are computed in, and the scalar is
    their mean — the size-independent checks at a size the oracle does not run at."""
    m, sch = build_mirror_model("cfg2_vp")
    eng = m.engine(math_mode=TF32)
    g = torch.Generator().manual_seed(9)
    B, L, C = 256, 256, 12
    x0 = torch.randn(B, L, C, generator=g)
    t = torch.rand(B, generator=g) * (1 - 1e-5) + 1e-5
    z = torch.randn(B, L, C, generator=g)
    loss, losses = eng.sde_loss(x0, t, z)
    assert torch.isfinite(losses).all() and float(loss) == pytest.approx(float(losses.double().mean()), rel=1e-6)
    idx = torch.tensor([0, 17, 100, 255])
    _, sub = eng.sde_loss(x0[idx], t[idx], z[idx])
    assert rel_err(sub, losses[idx.to(losses.device)]) < 5e-3
    # un-trained network, score ~ O(1): the un-weighted loss at small t is dominated by |z / std|^2 * w = mean(z^2) / L-ish; sanity bound
    assert 0.0 < float(loss) < 10.0
