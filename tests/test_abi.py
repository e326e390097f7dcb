"""CPU: the C-ABI library loads, exports every symbol include/fdiff_b200.h declares, and fails loudly without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from fourierdiffusion_b200 import _lib

HEADER = os.path.join(ROOT, "include", "fdiff_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in fdiff_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding table and header disagree"
    assert lib.fd_abi_version() == _lib.FD_ABI_VERSION


def test_config_struct_layout_matches_header():
    # int32 x10, double x2, int32 x2 -> 64 bytes with natural alignment; fd_create rejects any other struct_size
    assert C.sizeof(_lib.FdConfig) == 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import fourierdiffusion_b200 as fd

    lib = _lib.load()
    cfg = _lib.FdConfig(struct_size=C.sizeof(_lib.FdConfig), device=0, model_kind=0, max_len=8, n_channels=2, d_model=8, n_head=2,
                        num_layers=1, d_ff=16, sched_kind=0, sched_p0=0.1, sched_p1=20.0, fourier_noise_scaling=1, math_mode=0)
    h = C.c_void_p()
    assert lib.fd_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.fd_last_error()
    sch = fd.VPScheduler()
    m = fd.ScoreModule(n_channels=2, max_len=8, noise_scheduler=sch, d_model=8, n_head=2, num_layers=1)
    with pytest.raises(_lib.FdError):
        fd.DiffusionSampler(score_model=m, sample_batch_size=2).sample(2, 3)
    with pytest.raises(_lib.FdError):
        fd.dft(torch.zeros(1, 8, 2))
    with pytest.raises(_lib.FdError):
        m(fd.DiffusableBatch(X=torch.zeros(1, 8, 2), timesteps=torch.ones(1)))


def test_bad_arguments_are_rejected():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.FdConfig(struct_size=4)
    assert lib.fd_create(C.byref(cfg), C.byref(h)) != 0
    assert b"size mismatch" in lib.fd_last_error()
    cfg = _lib.FdConfig(struct_size=C.sizeof(_lib.FdConfig), sched_kind=7, max_len=8, n_channels=1, d_model=8, n_head=2, d_ff=4)
    assert lib.fd_create(C.byref(cfg), C.byref(h)) != 0
    assert b"Scheduler not recognized" in lib.fd_last_error()
    assert lib.fd_destroy(None) == 0


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fourierdiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f"{f} mentions the oracle"
