"""Sample-quality metrics with the reference's interface (src/fdiff/sampling/metrics.py:13-199): the frequency-domain copies of the
samples, the spectral densities and every Wasserstein distance are computed by the CUDA library.  cmd/sample.py:85 builds a
`MetricCollection` through Hydra (`cmd_conf/.../metrics`); pointing those `_target_`s at this module keeps the run on the GPU after the
sampler has finished."""
from __future__ import annotations

from abc import ABC, abstractmethod
from functools import partial
from typing import Any, Optional

import numpy as np
import torch

from .fourier import dft, spectral_density
from .wasserstein import WassersteinDistances


def check_flat_array(x: torch.Tensor | np.ndarray) -> np.ndarray:
    """utils/tensors.py:5-24."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if x.ndim > 2:
        x = x.reshape(x.shape[0], -1)
    assert isinstance(x, np.ndarray), f"x must be a numpy array or a torch tensor. Got {type(x)}"
    assert x.ndim == 2, f"x must be a 2d array. Got {x.ndim}d array."
    return x


class Metric(ABC):
    """metrics.py:13-26."""

    def __init__(self, original_samples: np.ndarray | torch.Tensor) -> None:
        self.original_samples = check_flat_array(original_samples)

    @abstractmethod
    def __call__(self, other_samples: np.ndarray | torch.Tensor) -> dict[str, Any]: ...

    @property
    @abstractmethod
    def name(self) -> str: ...

    @property
    def baseline_metrics(self) -> dict[str, float]:
        return {}


class _WassersteinMetric(Metric):
    prefix = ""

    def __init__(self, original_samples, random_seed: int, save_all_distances: bool = False) -> None:
        super().__init__(original_samples=original_samples)
        self.random_seed = random_seed
        self.save_all_distances = save_all_distances

    def _distances(self, wd: WassersteinDistances) -> np.ndarray:
        raise NotImplementedError

    def __call__(self, other_samples) -> dict[str, Any]:
        wd = WassersteinDistances(original_data=self.original_samples, other_data=check_flat_array(other_samples), seed=self.random_seed)
        distances = self._distances(wd)
        metrics: dict[str, Any] = {f"{self.prefix}_mean": float(np.mean(distances)), f"{self.prefix}_max": float(np.max(distances))}
        if self.save_all_distances:
            metrics[f"{self.prefix}_all"] = distances.tolist()
        return metrics

    @property
    def baseline_metrics(self) -> dict[str, float]:
        """Distances between two folds of the original samples, and to a generator that only outputs the average sample.
        metrics.py:130-160, 187-216."""
        n_samples = self.original_samples.shape[0]
        wd_self = WassersteinDistances(original_data=self.original_samples[: n_samples // 2], other_data=self.original_samples[n_samples // 2:],
                                       seed=self.random_seed)
        distances_self = self._distances(wd_self)
        avg_sample = np.mean(self.original_samples, axis=0, keepdims=True)
        wd_dummy = WassersteinDistances(original_data=self.original_samples, other_data=avg_sample, seed=self.random_seed)
        distances_dummy = self._distances(wd_dummy)
        return {
            f"{self.prefix}_mean_self": float(np.mean(distances_self)),
            f"{self.prefix}_max_self": float(np.max(distances_self)),
            f"{self.prefix}_mean_dummy": float(np.mean(distances_dummy)),
            f"{self.prefix}_max_dummy": float(np.max(distances_dummy)),
        }

    @property
    def name(self) -> str:
        return self.prefix


class SlicedWasserstein(_WassersteinMetric):
    """metrics.py:102-164."""
    prefix = "sliced_wasserstein"

    def __init__(self, original_samples, random_seed: int, num_directions: int, save_all_distances: bool = False) -> None:
        super().__init__(original_samples=original_samples, random_seed=random_seed, save_all_distances=save_all_distances)
        self.num_directions = num_directions

    def _distances(self, wd: WassersteinDistances) -> np.ndarray:
        return wd.sliced_distances(self.num_directions)


class MarginalWasserstein(_WassersteinMetric):
    """metrics.py:167-220."""
    prefix = "marginal_wasserstein"

    def _distances(self, wd: WassersteinDistances) -> np.ndarray:
        return wd.marginal_distances()


class MetricCollection:
    """metrics.py:29-99: every metric in the time and in the frequency domain (+ optionally the marginal distances of the spectral density)."""

    def __init__(self, metrics: list, original_samples: Optional[np.ndarray | torch.Tensor] = None, include_baselines: bool = True,
                 include_spectral_density: bool = False) -> None:
        metrics_time: list[Metric] = []
        metrics_freq: list[Metric] = []
        original_samples_freq = dft(torch.as_tensor(original_samples)) if original_samples is not None else None
        for metric in metrics:
            if isinstance(metric, partial):  # partially instantiated by Hydra: bind the original samples
                assert original_samples is not None, f"Original samples must be provided for metric {metric} to be instantiated."
                metrics_time.append(metric(original_samples=original_samples))
                metrics_freq.append(metric(original_samples=original_samples_freq))
        self.metrics_time = metrics_time
        self.metrics_freq = metrics_freq
        self.include_baselines = include_baselines
        self.metric_spectral = (
            MarginalWasserstein(original_samples=spectral_density(torch.as_tensor(original_samples)), random_seed=42, save_all_distances=True)
            if include_spectral_density else None)

    def __call__(self, other_samples: np.ndarray | torch.Tensor) -> dict[str, Any]:
        other = torch.as_tensor(other_samples)
        metric_dict: dict[str, Any] = {}
        other_samples_freq = dft(other)
        for metric_time, metric_freq in zip(self.metrics_time, self.metrics_freq):
            metric_dict.update({f"time_{k}": v for k, v in metric_time(other).items()})
            metric_dict.update({f"freq_{k}": v for k, v in metric_freq(other_samples_freq).items()})
        if self.include_baselines:
            metric_dict.update(self.baseline_metrics)
        if self.metric_spectral is not None:
            metric_dict.update({f"spectral_{k}": v for k, v in self.metric_spectral(spectral_density(other)).items()})
        return dict(sorted(metric_dict.items(), key=lambda item: item[0]))

    @property
    def baseline_metrics(self) -> dict[str, float]:
        metric_dict = {}
        for metric_time, metric_freq in zip(self.metrics_time, self.metrics_freq):
            metric_dict.update({f"time_{k}": v for k, v in metric_time.baseline_metrics.items()})
            metric_dict.update({f"freq_{k}": v for k, v in metric_freq.baseline_metrics.items()})
        return metric_dict
