"""`WassersteinDistances` with the reference's interface (src/fdiff/utils/wasserstein.py:12-199), computed by the CUDA library: the
projections, the per-direction sorts and the 1-D optimal transport (POT's `ot.emd2_1d` in the reference) run on the GPU (csrc/fd_wass.cu);
the random directions come from the same `numpy.random.default_rng(seed)` stream as the reference's, so a seed gives the same directions.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

_MAX_DIRS_PER_CALL = 4096  # bounds the projection workspace (n_dirs x (n + m) floats)


class WassersteinDistances:
    """Wasserstein-2 distances between two (n, d) data sets along one-dimensional projections.  wasserstein.py:12-42."""

    def __init__(self, original_data: np.ndarray, other_data: np.ndarray, normalisation: Optional[str] = "none",
                 seed: Optional[int] = None, device: Optional[torch.device] = None) -> None:
        self.original_data = original_data
        self.other_data = other_data
        self.normalisation = normalisation
        self.rng = np.random.default_rng(seed)
        self._device = device

    # ---- directions: host side, bit-identical to the reference (wasserstein.py:44-93) ----
    def random_direction(self, dim: int) -> np.ndarray:
        vector = self.rng.normal(size=dim)
        return vector / np.linalg.norm(vector)

    def get_random_directions(self, n_directions: int) -> list[np.ndarray]:
        dimension = self.original_data.shape[1]
        return [self.random_direction(dimension) for _ in range(n_directions)]

    def get_marginal_directions(self) -> list[np.ndarray]:
        dimension = self.original_data.shape[1]
        return [np.identity(dimension)[i] for i in range(dimension)]

    # ---- distances: GPU ----
    def _dev(self) -> torch.device:
        if not torch.cuda.is_available():
            raise _lib.FdError("no CUDA device visible: fourierdiffusion_b200 has no CPU fallback")
        return self._device if self._device is not None else torch.device("cuda", torch.cuda.current_device())

    def _standardise_flag(self) -> int:
        if self.normalisation == "none":
            return 0
        if self.normalisation == "standardise":
            return 1
        raise ValueError(f"Unrecognised normalisation type: {self.normalisation}")  # wasserstein.py:158

    def _distances(self, directions: Optional[np.ndarray]) -> np.ndarray:
        lib = _lib.load()
        dev = self._dev()
        x = torch.as_tensor(np.ascontiguousarray(self.original_data, dtype=np.float32)).to(dev)
        y = torch.as_tensor(np.ascontiguousarray(self.other_data, dtype=np.float32)).to(dev)
        assert x.dim() == 2 and y.dim() == 2 and x.shape[1] == y.shape[1], (tuple(x.shape), tuple(y.shape))
        n, d = x.shape
        m = y.shape[0]
        flag = self._standardise_flag()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        outs = []
        with torch.cuda.device(dev):
            if directions is None:  # marginal: the standard basis, all d features in one call
                out = torch.empty(d, dtype=torch.float64, device=dev)
                _lib.check(lib.fd_wasserstein(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), None, n, m, d, d, flag,
                                              C.c_void_p(out.data_ptr()), dev.index, stream))
                outs.append(out)
            else:
                dirs = torch.as_tensor(np.ascontiguousarray(directions, dtype=np.float64)).to(dev)
                for k0 in range(0, dirs.shape[0], _MAX_DIRS_PER_CALL):
                    blk = dirs[k0:k0 + _MAX_DIRS_PER_CALL].contiguous()
                    out = torch.empty(blk.shape[0], dtype=torch.float64, device=dev)
                    _lib.check(lib.fd_wasserstein(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(blk.data_ptr()), n, m, d,
                                                  blk.shape[0], flag, C.c_void_p(out.data_ptr()), dev.index, stream))
                    outs.append(out)
        return torch.cat(outs).cpu().numpy()

    def feature_distance(self, feature: int) -> float:
        """wasserstein.py:95-117."""
        e = np.zeros((1, self.original_data.shape[1]))
        e[0, feature] = 1.0
        return float(self._distances(e)[0])

    def directional_distance(self, direction: np.ndarray) -> float:
        """wasserstein.py:119-142 (returns the W_2 distance, like the reference's code; its docstring says W_2^2)."""
        return float(self._distances(np.asarray(direction, dtype=np.float64)[None, :])[0])

    def sliced_distances(self, num_directions: int) -> np.ndarray:
        """wasserstein.py:160-179: one distance per random direction (all directions in one batched GPU pass)."""
        return self._distances(np.stack(self.get_random_directions(num_directions)))

    def marginal_distances(self) -> np.ndarray:
        """wasserstein.py:181-199: one distance per feature."""
        return self._distances(None)
