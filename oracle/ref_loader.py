"""Loader for the UNMODIFIED reference (`/root/reference/src/fdiff`) — TEST INFRASTRUCTURE ONLY.

Used to (1) pin `oracle/fdiff_oracle.py` against the real reference and (2) generate the golden
vectors under `tests/golden/` (script: `tests/golden/make_golden.py`). `/root/reference` exists only in
the build container, never on the GPU box, so nothing in `-m gpu` tests, `smoke()` or `bench.py`
may call `load_reference()`; tests that use it are skipped when the tree is absent.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

REFERENCE_SRC = os.environ.get("FDIFF_REFERENCE_SRC", "/root/reference/src")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "fdiff"))


def load_reference() -> SimpleNamespace:
    """Import the reference modules on the hot path and return them in a namespace."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    try:
        import pytorch_lightning  # noqa: F401  (real one, if ever installed)
    except ImportError:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import tqdm as _tqdm_mod  # silence the progress bars of sampler.py:67-89
    from functools import partialmethod

    _tqdm_mod.tqdm.__init__ = partialmethod(_tqdm_mod.tqdm.__init__, disable=True)

    from fdiff.models import score_models, transformer
    from fdiff.sampling import sampler
    from fdiff.schedulers import sde
    from fdiff.utils import dataclasses as dc
    from fdiff.utils import fourier, losses

    return SimpleNamespace(
        score_models=score_models,
        transformer=transformer,
        sampler=sampler,
        sde=sde,
        dataclasses=dc,
        fourier=fourier,
        losses=losses,
        ScoreModule=score_models.ScoreModule,
        LSTMScoreModule=score_models.LSTMScoreModule,
        MLPScoreModule=score_models.MLPScoreModule,
        DiffusionSampler=sampler.DiffusionSampler,
        VPScheduler=sde.VPScheduler,
        VEScheduler=sde.VEScheduler,
        DiffusableBatch=dc.DiffusableBatch,
        dft=fourier.dft,
        idft=fourier.idft,
    )
