"""Wasserstein metrics (reference src/fdiff/utils/wasserstein.py:95-199, src/fdiff/sampling/metrics.py:102-220).

CPU: the numpy oracle (oracle/wasserstein_oracle.py) against closed forms and against the ground truths of the reference's own metric
tests (tests/test_metrics.py:18-83; POT itself is not installed here).  GPU: the CUDA path (csrc/fd_wass.cu through
fourierdiffusion_b200.wasserstein / .metrics) against the oracle, tolerance 1e-6 relative (fp32 projections, fp64 transport sums).
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wasserstein_oracle as WO  # noqa: E402

TOL = 1e-6


# ---- the oracle ---------------------------------------------------------------------------------------------------------------------
def test_oracle_emd2_1d_closed_forms():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=64), rng.normal(size=64) + 0.3
    assert abs(WO.emd2_1d(a, b) - np.mean((np.sort(a) - np.sort(b)) ** 2)) < 1e-14  # equal sizes: order statistics are matched
    assert abs(WO.emd2_1d([0.0, 1.0], [0.5]) - 0.25) < 1e-15                          # one-point target
    assert abs(WO.emd2_1d(a, a + 2.5) - 6.25) < 1e-12                                  # pure shift
    # unequal sizes: integral over the merged quantile grid
    a, b = np.sort(rng.normal(size=7)), np.sort(rng.normal(size=5))
    tot = sum(max(0, min((i + 1) * 5, (j + 1) * 7) - max(i * 5, j * 7)) * (a[i] - b[j]) ** 2 for i in range(7) for j in range(5)) / 35.0
    assert abs(WO.emd2_1d(a, b) - tot) < 1e-14


@pytest.mark.parametrize("shift", [0.0, 0.1, 1.0])
def test_oracle_reference_metric_ground_truths(shift):
    """tests/test_metrics.py:56-83 (marginal) and :18-53 (sliced) of the reference, with the oracle in place of fdiff + POT."""
    np.random.seed(42)
    d1 = np.random.rand(1000, 2, 1).reshape(1000, -1)
    d2 = (np.random.rand(1000, 2, 1) + shift).reshape(1000, -1)
    wd = WO.WassersteinDistances(d1, d2, seed=42)
    marg = wd.marginal_distances()
    assert abs(marg.mean() - shift) <= 0.1 and abs(marg.max() - shift) <= 0.1
    sl = wd.sliced_distances(20)
    assert sl.mean() <= sl.max() and np.all(sl <= np.sqrt(2.0) * (shift + 0.1))  # a projection cannot exceed the shift vector's length


# ---- the CUDA path ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,m,d", [(1000, 1000, 2), (777, 1234, 5), (5000, 1, 3), (1, 9, 4), (4097, 4096, 2), (20000, 9000, 3), (64, 64, 1)])
@pytest.mark.parametrize("norm", ["none", "standardise"])
def test_gpu_wasserstein_matches_oracle(n, m, d, norm):
    from fourierdiffusion_b200.wasserstein import WassersteinDistances

    rng = np.random.default_rng(n + m + d)
    x = rng.normal(size=(n, d)).astype(np.float32) * 1.7 + 0.2
    y = (rng.normal(size=(m, d)) ** 3).astype(np.float32) - 0.4
    if norm == "standardise" and n == 1:
        pytest.skip("np.std of a single sample is 0: the reference divides by zero")
    K = 6
    got = WassersteinDistances(x, y, normalisation=norm, seed=3)
    want = WO.WassersteinDistances(x.astype(np.float64), y.astype(np.float64), normalisation=norm, seed=3)
    gs, ws = got.sliced_distances(K), want.sliced_distances(K)
    assert gs.shape == (K,) and np.allclose(gs, ws, rtol=TOL, atol=1e-7), (gs, ws)
    gm, wm = got.marginal_distances(), want.marginal_distances()
    assert gm.shape == (d,) and np.allclose(gm, wm, rtol=TOL, atol=1e-7), (gm, wm)
    assert abs(got.feature_distance(d - 1) - wm[d - 1]) <= TOL * max(1.0, wm[d - 1])
    direction = want.random_direction(d)
    assert abs(got.directional_distance(direction) - want.directional_distance(direction)) <= TOL * 10


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0.0, 0.1, 1.0])
def test_reference_metric_tests_ported(shift):
    """The reference's tests/test_metrics.py:18-83 on the GPU metrics (its POT cross-check is replaced by the oracle)."""
    from fourierdiffusion_b200.metrics import MarginalWasserstein, SlicedWasserstein

    np.random.seed(42)
    dataset1 = np.random.rand(1000, 2, 1)
    dataset2 = np.random.rand(1000, 2, 1) + shift
    sw = SlicedWasserstein(original_samples=dataset1, random_seed=42, num_directions=1000, save_all_distances=True)
    metrics = sw(dataset2)
    assert abs(metrics["sliced_wasserstein_mean"] - np.mean(metrics["sliced_wasserstein_all"])) <= 1e-5
    assert metrics["sliced_wasserstein_mean"] <= metrics["sliced_wasserstein_max"]
    want = WO.WassersteinDistances(dataset1.reshape(1000, -1), dataset2.reshape(1000, -1), seed=42).sliced_distances(40)
    assert np.allclose(metrics["sliced_wasserstein_all"][:40], want, rtol=TOL, atol=1e-7)
    mw = MarginalWasserstein(original_samples=dataset1, random_seed=42, save_all_distances=True)
    metrics = mw(dataset2)
    assert abs(metrics["marginal_wasserstein_mean"] - np.mean(metrics["marginal_wasserstein_all"])) <= 1e-5
    assert metrics["marginal_wasserstein_mean"] <= metrics["marginal_wasserstein_max"]
    assert abs(metrics["marginal_wasserstein_mean"] - shift) <= 0.1 and abs(metrics["marginal_wasserstein_max"] - shift) <= 0.1
    base = mw.baseline_metrics  # two folds of the original samples; a generator that outputs the average sample (m = 1)
    assert set(base) == {"marginal_wasserstein_mean_self", "marginal_wasserstein_max_self", "marginal_wasserstein_mean_dummy",
                         "marginal_wasserstein_max_dummy"}
    flat = dataset1.reshape(1000, -1)
    dummy = WO.WassersteinDistances(flat, flat.mean(axis=0, keepdims=True)).marginal_distances()
    assert abs(base["marginal_wasserstein_mean_dummy"] - dummy.mean()) <= 1e-6


@pytest.mark.gpu
def test_metric_collection_like_cmd_sample():
    """cmd/sample.py:85 -> metrics.py:29-99: time- and frequency-domain metrics + the spectral-density marginals, all on the GPU."""
    from functools import partial

    from fourierdiffusion_b200.metrics import MarginalWasserstein, MetricCollection, SlicedWasserstein
    from oracle import fdiff_oracle as O

    g = torch.Generator().manual_seed(0)
    orig = torch.randn(96, 24, 3, generator=g)
    other = torch.randn(80, 24, 3, generator=g) * 1.2 + 0.1
    mc = MetricCollection(metrics=[partial(SlicedWasserstein, random_seed=7, num_directions=16), partial(MarginalWasserstein, random_seed=7)],
                          original_samples=orig, include_baselines=True, include_spectral_density=True)
    res = mc(other)
    assert list(res) == sorted(res)
    flat = lambda t: t.reshape(t.shape[0], -1).double().numpy()
    want_t = WO.WassersteinDistances(flat(orig), flat(other), seed=7).sliced_distances(16)
    want_f = WO.WassersteinDistances(flat(O.dft(orig)), flat(O.dft(other)), seed=7).marginal_distances()
    assert abs(res["time_sliced_wasserstein_mean"] - want_t.mean()) <= 1e-5 * want_t.mean()
    assert abs(res["freq_marginal_wasserstein_max"] - want_f.max()) <= 1e-5 * want_f.max()
    spec = lambda t: flat(O.spectral_density(t))
    want_s = WO.WassersteinDistances(spec(orig), spec(other), seed=42).marginal_distances()
    assert abs(res["spectral_marginal_wasserstein_mean"] - want_s.mean()) <= 1e-5 * want_s.mean()
    assert "time_sliced_wasserstein_mean_self" in res and "freq_marginal_wasserstein_max_dummy" in res
