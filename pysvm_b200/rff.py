"""Random Fourier features for the RBF kernel (reference ``pysvm/rff.py``)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from sklearn.base import TransformerMixin

from . import _native as N
from .solver import default_device, to_device_matrix


def rff_transform_device(X: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """sqrt(2/D) * cos(X w^T + b) as a CUDA matrix (rff.py:19-20): float64 like the reference (b2svm_rff_transform), or
    float32 for prediction-only pipelines (b2svm_rff_transform_f32: tensor cores, float32 X with at most 96 features)."""
    lib = N.load()
    X = X.contiguous()
    D, d = w.shape
    assert X.shape[1] == d
    torch.cuda.synchronize(X.device)
    if dtype == torch.float32:
        if X.dtype != torch.float32:
            raise TypeError("float32 features need a float32 X")
        ldo = (D + 3) // 4 * 4
        out = torch.empty((X.shape[0], ldo), dtype=torch.float32, device=X.device)
        N.check(lib.b2svm_rff_transform_f32(X.device.index or 0, None, X.data_ptr(), X.shape[0], X.stride(0), d, w.data_ptr(),
                                            b.data_ptr(), D, out.data_ptr(), ldo))
        return out if ldo == D else out[:, :D]
    out = torch.empty((X.shape[0], D), dtype=torch.float64, device=X.device)
    N.check(lib.b2svm_rff_transform(X.device.index or 0, None, N.F64 if X.dtype == torch.float64 else N.F32,
                                    X.data_ptr(), X.shape[0], X.stride(0), d, w.data_ptr(), b.data_ptr(), D,
                                    out.data_ptr()))
    return out


def rff_decision_device(Xq: torch.Tensor, w: torch.Tensor, b: torch.Tensor, wz: torch.Tensor) -> torch.Tensor:
    """z(x_q) . wz for every query row without materialising z (b2svm_rff_decision)."""
    lib = N.load()
    Xq = Xq.contiguous()
    D, d = w.shape
    out = torch.empty(Xq.shape[0], dtype=torch.float64, device=Xq.device)
    torch.cuda.synchronize(Xq.device)
    N.check(lib.b2svm_rff_decision(Xq.device.index or 0, None, N.F64 if Xq.dtype == torch.float64 else N.F32,
                                   Xq.data_ptr(), Xq.shape[0], Xq.stride(0), d, w.data_ptr(), b.data_ptr(), D,
                                   wz.contiguous().data_ptr(), out.data_ptr()))
    return out


class NormalRFF(TransformerMixin):
    """Same constructor / methods / RNG consumption as the reference (rff.py:5-23): ``fit`` draws
    ``w = sqrt(2*gamma) * randn(D, d)`` then ``b ~ U(0, 2*pi)`` from NumPy's global RNG on the host;
    ``transform`` runs on the device and returns a float64 NumPy array."""

    def __init__(self, gamma=1, D=1000) -> None:
        super().__init__()
        self.gamma = gamma
        self.D = D

    def fit(self, X):
        self.n_features = np.array(X).shape[1]
        self.w = np.sqrt(self.gamma * 2) * np.random.randn(self.D, self.n_features)
        self.b = np.random.uniform(0, 2 * np.pi, self.D)
        self._w_dev = self._b_dev = None
        return self

    def _device_params(self, device):
        if self._w_dev is None or self._w_dev.device != device:
            self._w_dev = torch.from_numpy(np.ascontiguousarray(self.w)).to(device)
            self._b_dev = torch.from_numpy(np.ascontiguousarray(self.b)).to(device)
        return self._w_dev, self._b_dev

    def transform_device(self, X, dtype=torch.float64) -> torch.Tensor:
        device = X.device if isinstance(X, torch.Tensor) else default_device()
        Xd = to_device_matrix(X, device)
        w, b = self._device_params(device)
        return rff_transform_device(Xd, w, b, dtype=dtype)

    def transform(self, X):
        return self.transform_device(X).cpu().numpy()

    def fit_transform(self, X):
        return self.fit(X).transform(X)
