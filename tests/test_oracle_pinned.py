"""CPU, build container only: the oracle against the UNMODIFIED reference imported from /root/reference (skipped where that
tree is absent, e.g. on the GPU box — there the committed golden vectors pin it, tests/test_oracle_golden.py)."""
from __future__ import annotations

import pytest
import torch

from conftest import cases, rel_err
from oracle import fdiff_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load_reference()


def _ref_model(R, name):
    c = cases.SCORE_CASES[name]
    torch.manual_seed(cases.WEIGHT_SEED)
    Sched = {"vp": R.VPScheduler, "ve": R.VEScheduler}[c["sched"]]
    sch = Sched(fourier_noise_scaling=c["fourier"], **cases.SCHED_KW[c["sched"]])
    Model = {"transformer": R.ScoreModule, "lstm": R.LSTMScoreModule, "mlp": R.MLPScoreModule}[c["model"]]
    m = Model(n_channels=c["C"], max_len=c["L"], noise_scheduler=sch, fourier_noise_scaling=c["fourier"], **c["kw"]).eval()
    sch.set_noise_scaling(c["L"])
    return m, sch


@pytest.mark.parametrize("name", ["tiny_vp", "classdefault_ve", "ecg_vp", "mimic_lstm_vp", "mlp_vp"])
def test_score_against_reference(R, name):
    m, sch = _ref_model(R, name)
    c = cases.SCORE_CASES[name]
    x = torch.randn(c["B"], c["L"], c["C"], generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        for _ in range(3):
            m(R.DiffusableBatch(X=x, y=None, timesteps=torch.full((c["B"],), 0.5)))
        spec = O.model_spec_from_module(m)
        for t in (0.9, 0.01):
            tv = torch.full((c["B"],), t, dtype=torch.float32)
            want = m(R.DiffusableBatch(X=x, y=None, timesteps=tv))
            assert rel_err(O.score(spec, x, tv), want) < 5e-6


@pytest.mark.parametrize("kind", ["vp", "ve"])
@pytest.mark.parametrize("L", [24, 187, 256])
def test_step_and_prior_bit_exact(R, kind, L):
    Sched = {"vp": R.VPScheduler, "ve": R.VEScheduler}[kind]
    sch = Sched(fourier_noise_scaling=True, **cases.SCHED_KW[kind])
    sch.set_noise_scaling(L)
    sch.set_timesteps(1000)
    spec = O.scheduler_spec_from_object(sch)
    g = torch.Generator().manual_seed(3)
    x, s, z = (torch.randn(2, L, 3, generator=g) for _ in range(3))
    G = O.g_vector(L, True)
    assert torch.equal(G, sch.G)
    ts, dt = O.make_timesteps(1000, sch.eps)
    assert torch.equal(ts, sch.timesteps) and torch.equal(dt, sch.step_size)
    for t in (1.0, 0.5, 1e-5):
        real_randn_like = torch.randn_like
        torch.randn_like = lambda *a, **k: z.clone()
        try:
            want = sch.step(s, t, x).prev_sample
        finally:
            torch.randn_like = real_randn_like
        assert torch.equal(O.scheduler_step(spec, x, s, z, t, G, dt), want)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: z.clone()
    try:
        want = sch.prior_sampling((2, L, 3))
    finally:
        torch.randn = real_randn
    assert torch.equal(O.prior_from_noise(z, G, spec.sigma_max if kind == "ve" else None), want)


@pytest.mark.parametrize("L", [7, 8, 100, 101, 252])
def test_dft_idft_against_reference(R, L):
    x = torch.randn(3, L, 2, generator=torch.Generator().manual_seed(L))
    assert rel_err(O.dft(x), R.dft(x)) < 2e-6
    assert rel_err(O.idft(x), R.idft(x)) < 2e-6


def test_renorm_fixed_point_matches_embedding(R):
    import math

    torch.manual_seed(0)
    pe = R.transformer.PositionalEncoding(d_model=72, max_len=256)
    x = torch.zeros(1, 256, 72)
    for _ in range(5):
        pe(x)
    raw = torch.manual_seed(0) and None
    torch.manual_seed(0)
    pe2 = R.transformer.PositionalEncoding(d_model=72, max_len=256)
    fixed = O.renorm_positional_table(pe2.embedding.weight.detach(), math.sqrt(72))
    assert rel_err(fixed, pe.embedding.weight.detach()) < 1e-6


@pytest.mark.parametrize("name", ["tiny_vp", "classdefault_ve", "cfg2_vp", "mimic_lstm_vp", "mlp_vp"])
def test_extraction_reads_the_real_reference_modules(R, name):
    """`extract_score_model` (what Engine.for_score_model uploads) is duck-typed; run it on the reference's OWN ScoreModule /
    LSTMScoreModule / MLPScoreModule objects and on this package's host mirror built from the same seed: identical config fields,
    identical keys, shapes and values — so the drop-in reads a real checkpointed reference module exactly like the mirror."""
    from conftest import build_mirror_model
    from fourierdiffusion_b200.engine import extract_score_model

    ref_m, _ = _ref_model(R, name)
    mir_m, _ = build_mirror_model(name)
    f_ref, w_ref = extract_score_model(ref_m)
    f_mir, w_mir = extract_score_model(mir_m)
    assert f_ref == f_mir
    assert set(w_ref) == set(w_mir)
    for k in w_ref:
        assert w_ref[k].dtype == torch.float32 and w_ref[k].is_contiguous() and w_ref[k].device.type == "cpu"
        assert w_ref[k].shape == w_mir[k].shape, k
        assert torch.equal(w_ref[k], w_mir[k]), k
    c = cases.SCORE_CASES[name]
    assert (f_ref["max_len"], f_ref["n_channels"]) == (c["L"], c["C"])
    assert f_ref["model_kind"] == {"transformer": 0, "lstm": 1, "mlp": 2}[c["model"]]
    assert f_ref["sched_kind"] == {"vp": 0, "ve": 1}[c["sched"]] and f_ref["fourier_noise_scaling"] == int(c["fourier"])
    if c["model"] == "transformer":
        assert f_ref["d_ff"] == 2048 and f_ref["n_head"] == c["kw"].get("n_head", 12)
    # the positional table is uploaded at the fixed point the reference reaches after a few forwards (transformer.py:13-15)
    if "pos_encoder.embedding.weight" in w_ref:
        x = torch.zeros(1, c["L"], c["C"])
        with torch.no_grad():
            for _ in range(4):
                ref_m(R.DiffusableBatch(X=x, y=None, timesteps=torch.full((1,), 0.5)))
        assert rel_err(w_ref["pos_encoder.embedding.weight"], ref_m.pos_encoder.embedding.weight.detach()) < 1e-6
