"""TEST INFRASTRUCTURE — CPU restatement of the Wasserstein metrics of the hot path's evaluation stage.  Never imported by the product.

Restates, in float64 numpy:
  * POT's `ot.emd2_1d(x_a, x_b)` with uniform weights and the default metric 'sqeuclidean' — the published algorithm of POT 0.9
    (ot/lp/emd_wrap.pyx `emd_1d_sorted`: sort both samples, then move mass greedily between the two sorted sequences); POT is a
    dependency of the reference (`pyproject.toml:50`, unpinned) that is NOT installed in this image, so this port is pinned by the
    reference's own metric tests instead (tests/test_metrics.py:18-83: ground truths of shifted uniform samples, mean/max consistency)
    and by closed forms (equal sizes: mean squared difference of the order statistics; one-point target; pure shift);
  * `WassersteinDistances` (src/fdiff/utils/wasserstein.py:12-199) on top of it, line by line.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def emd2_1d(x_a: np.ndarray, x_b: np.ndarray) -> float:
    """Optimal transport cost between two 1-D samples with uniform weights, ground cost |x - y|^2 (ot.emd2_1d defaults)."""
    u = np.sort(np.asarray(x_a, dtype=np.float64))
    v = np.sort(np.asarray(x_b, dtype=np.float64))
    n, m = u.shape[0], v.shape[0]
    w_i, w_j = 1.0 / n, 1.0 / m
    i = j = 0
    cost = 0.0
    while True:  # emd_1d_sorted: ship min(remaining mass) between the current pair, advance the exhausted side
        if w_i < w_j or j == m - 1:
            cost += (u[i] - v[j]) ** 2 * w_i
            i += 1
            if i == n:
                break
            w_j -= w_i
            w_i = 1.0 / n
        else:
            cost += (u[i] - v[j]) ** 2 * w_j
            j += 1
            if j == m:
                break
            w_i -= w_j
            w_j = 1.0 / m
    return float(cost)


class WassersteinDistances:
    """wasserstein.py:12-199."""

    def __init__(self, original_data: np.ndarray, other_data: np.ndarray, normalisation: Optional[str] = "none", seed: Optional[int] = None):
        self.original_data = original_data
        self.other_data = other_data
        self.normalisation = normalisation
        self.rng = np.random.default_rng(seed)

    def random_direction(self, dim: int) -> np.ndarray:  # :44-61
        vector = self.rng.normal(size=dim)
        return vector / np.linalg.norm(vector)

    def get_random_directions(self, n_directions: int) -> list:  # :63-76
        return [self.random_direction(self.original_data.shape[1]) for _ in range(n_directions)]

    def _normalise(self, orig, other):  # :149-158
        if self.normalisation == "none":
            return orig, other
        if self.normalisation == "standardise":
            sd = np.std(orig)
            return orig / sd, other / sd
        raise ValueError(f"Unrecognised normalisation type: {self.normalisation}")

    def feature_distance(self, feature: int) -> float:  # :95-117
        o, t = self._normalise(self.original_data[:, feature], self.other_data[:, feature])
        return float(np.sqrt(emd2_1d(o, t)))

    def directional_distance(self, direction: np.ndarray) -> float:  # :119-142
        o, t = self._normalise(self.original_data @ direction, self.other_data @ direction)
        return float(np.sqrt(emd2_1d(o, t)))

    def sliced_distances(self, num_directions: int) -> np.ndarray:  # :160-179
        return np.array([self.directional_distance(d) for d in self.get_random_directions(num_directions)])

    def marginal_distances(self) -> np.ndarray:  # :181-199
        return np.array([self.feature_distance(f) for f in range(self.original_data.shape[1])])
