"""Seeded synthetic inputs of BASELINE.json's configs: the shapes bench.py and the full-size
scripts run on.  NumPy's Generator is deterministic across platforms, so the GPU box regenerates the very arrays
the parity fixtures were made from (tests/test_abi_cpu.py compares these with the generators the tests use)."""
from __future__ import annotations

import numpy as np


def classification(n: int = 200_000, d: int = 128, seed: int = 0, sep: float = 1.0):
    """configs[2] (C3): two overlapping Gaussians, float32 rows, labels in {0, 1}."""
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.5).astype(int)
    mu = rng.standard_normal(d)
    mu /= np.linalg.norm(mu)
    shift = (2 * y[:, None] - 1) * sep * mu[None, :].astype(np.float32)
    X = (rng.standard_normal((n, d)).astype(np.float32) + shift).astype(np.float32)
    return X, y


def regression(n: int = 100_000, d: int = 64, seed: int = 0):
    """configs[3] (C4): z = sin(X w) + noise, float32."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    w = rng.standard_normal(d).astype(np.float32) / np.sqrt(d)
    z = (np.sin(X @ w) + 0.1 * rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    return X, z


def oneclass(n: int = 1_000_000, d: int = 32, seed: int = 1):
    """configs[4] (C5): standard normal rows, float32."""
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
