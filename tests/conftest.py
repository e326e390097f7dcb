"""Shared test plumbing.  Markers: `gpu` = needs a B200 (run on the GPU box with `-m gpu`); everything else runs on CPU."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402  (tests/golden/cases.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")
    # a fresh checkout has no libfdiff_b200.so yet (built artefacts are git-ignored): build it once, nvcc cross-compiles without a GPU
    lib = os.path.join(ROOT, "fourierdiffusion_b200", "libfdiff_b200.so")
    if not os.path.exists(lib) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        import __graft_entry__

        __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def build_mirror_model(name: str):
    """Host-mirror score module + scheduler for a golden case: same seed, same construction order as the reference, so
    the weights are identical to the ones the golden outputs were produced with (checked via fp64 checksums)."""
    import fourierdiffusion_b200 as fd

    c = cases.SCORE_CASES[name]
    torch.manual_seed(cases.WEIGHT_SEED)
    Sched = {"vp": fd.VPScheduler, "ve": fd.VEScheduler}[c["sched"]]
    sch = Sched(fourier_noise_scaling=c["fourier"], **cases.SCHED_KW[c["sched"]])
    Model = {"transformer": fd.ScoreModule, "lstm": fd.LSTMScoreModule, "mlp": fd.MLPScoreModule}[c["model"]]
    m = Model(n_channels=c["C"], max_len=c["L"], noise_scheduler=sch, fourier_noise_scaling=c["fourier"], **c["kw"]).eval()
    sch.set_noise_scaling(c["L"])
    return m, sch


def load_golden(name: str):
    g = np.load(os.path.join(GOLDEN, f"score_{name}.npz"))
    out = {k: g[k] for k in g.files if k != "checksums"}
    out["checksums"] = json.loads(str(g["checksums"]))
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| — every tolerance in the suite is stated in this norm (SURVEY.md §7 hard part 4)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
