"""DiffusionDataset on the GPU (first half; reference src/fdiff/dataloaders/datamodules.py:23-65): DFT, per-feature
statistics and standardisation against the reference's own torch expressions evaluated on the CPU (the oracle for this row IS the
reference call: `X.mean(dim=0)`, `X.std(dim=0)`, `(x - mean) / std`)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import rel_err  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1000, 24, 40), (37, 256, 12), (2, 7, 3), (5000, 187, 1)])
@pytest.mark.parametrize("fourier", [False, True])
def test_diffusion_dataset_matches_reference_expressions(shape, fourier):
    from fourierdiffusion_b200.datasets import DiffusionDataset, standardise
    from oracle import fdiff_oracle as O

    g = torch.Generator().manual_seed(sum(shape))
    X = torch.randn(*shape, generator=g) * 3.0 + 1.5
    ds = DiffusionDataset(X, fourier_transform=fourier, standardize=True)
    Xr = O.dft(X) if fourier else X  # datamodules.py:42-43
    mean, std = Xr.mean(dim=0), Xr.std(dim=0)  # datamodules.py:52-53
    assert len(ds) == shape[0]
    assert rel_err(ds.feature_mean, mean) < 5e-6 and rel_err(ds.feature_std, std) < 5e-6
    want = (Xr - mean) / std  # datamodules.py:62
    for i in (0, shape[0] - 1):
        assert rel_err(ds[i]["X"], want[i]) < 2e-5
    # standardise / de-standardise are exact inverses up to rounding, and the inverse matches cmd/sample.py:76-78
    back = standardise(ds.standardized(), ds.feature_mean, ds.feature_std, inverse=True)
    assert rel_err(back, ds.X) < 1e-5
    assert torch.equal(standardise(Xr, mean, std, inverse=True), Xr * std + mean)


@pytest.mark.gpu
def test_reference_statistics_with_x_ref_and_labels():
    from fourierdiffusion_b200.datasets import DiffusionDataset

    X, X_ref, y = torch.randn(20, 16, 2), torch.randn(64, 16, 2) * 2.0, torch.arange(20)
    ds = DiffusionDataset(X, y=y, fourier_transform=False, standardize=False, X_ref=X_ref)
    assert rel_err(ds.feature_std, X_ref.std(dim=0)) < 5e-6
    item = ds[3]
    assert torch.equal(item["X"], X[3]) and int(item["y"]) == 3
