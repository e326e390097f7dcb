"""fourierdiffusion_b200 — B200-native sampling hot path of JonathanCrabbe/FourierDiffusion ("fdiff").

Host-side mirror of the reference interface for the path (same names, arguments and error behaviour):

    reference                                         here
    fdiff.sampling.sampler.DiffusionSampler      ->   fourierdiffusion_b200.sampler.DiffusionSampler  (alias Sampler)
    fdiff.models.score_models.ScoreModule etc.   ->   fourierdiffusion_b200.score_models.*
    fdiff.schedulers.sde.VPScheduler/VEScheduler ->   fourierdiffusion_b200.schedulers.*
    fdiff.utils.fourier.dft / idft / spectral_density -> fourierdiffusion_b200.fourier.dft / idft / spectral_density
    fdiff.utils.dataclasses.DiffusableBatch      ->   fourierdiffusion_b200.batch.DiffusableBatch
    fdiff.dataloaders.datamodules.DiffusionDataset ->  fourierdiffusion_b200.datasets.DiffusionDataset (DFT, feature statistics, standardisation)
    fdiff.utils.losses.get_sde_loss_fn (train=False)  ->  fourierdiffusion_b200.losses.get_sde_loss_fn  (ScoreModule.validation_step)
    fdiff.utils.wasserstein.WassersteinDistances ->   fourierdiffusion_b200.wasserstein.WassersteinDistances
    fdiff.sampling.metrics.SlicedWasserstein / MarginalWasserstein / MetricCollection -> fourierdiffusion_b200.metrics.*

All arithmetic runs in libfdiff_b200.so (hand-written sm_100a CUDA behind the C ABI of include/fdiff_b200.h).  There is
no CPU / PyTorch fallback: without the library or without a B200 every compute entry point raises.
"""
from .batch import DiffusableBatch
from .datasets import DiffusionDataset
from .fourier import dft, idft, spectral_density
from .losses import get_sde_loss_fn
from .metrics import MarginalWasserstein, MetricCollection, SlicedWasserstein
from .sampler import DiffusionSampler, Sampler
from .schedulers import SDE, SamplingOutput, VEScheduler, VPScheduler
from .score_models import LSTMScoreModule, MLPScoreModule, ScoreModule
from .wasserstein import WassersteinDistances

__all__ = [
    "DiffusableBatch", "DiffusionSampler", "Sampler", "SDE", "SamplingOutput", "VEScheduler", "VPScheduler",
    "ScoreModule", "LSTMScoreModule", "MLPScoreModule", "dft", "idft", "spectral_density",
    "SlicedWasserstein", "MarginalWasserstein", "MetricCollection", "WassersteinDistances", "DiffusionDataset", "get_sde_loss_fn",
]
