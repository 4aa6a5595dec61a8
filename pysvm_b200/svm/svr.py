"""Regression estimators with PySVM's API (reference ``pysvm/svm/svr.py``): 2l-variable problems
whose variables t and t+l share training row t ("mirror" mode of the device solver)."""
from __future__ import annotations

import numpy as np
import torch
from sklearn.metrics import r2_score

from ..solver import KernelSpec, NuSolverWithCache, QColumns, SolverWithCache, weighted_rows
from .svc import BiKernelSVC, BiLinearSVC, _Fitted


def _svr_vectors(z, l, eps):
    y = np.empty(2 * l)
    y[:l], y[l:] = 1., -1.
    p = np.ones(2 * l) * eps       # svr.py:161-163
    p[:l] -= z
    p[l:] += z
    return y, p


class LinearSVR(BiLinearSVC):
    """svr.py:8-107."""
    _name = "LinearSVR"

    def __init__(self, C: float = 1., eps: float = 0., max_iter: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256) -> None:
        super().__init__(C, max_iter, tol, cache_size)
        self.eps = eps

    def fit(self, X, y):
        X, z = np.array(X), np.array(y)
        l, self.n_features = X.shape
        y, p = _svr_vectors(z, l, self.eps)
        func = QColumns(X, y, KernelSpec("linear"), mirror=True)
        solver = SolverWithCache(p, y, self.C, self.tol, self.cache_size, func=func, max_iter_hint=self.max_iter)
        self._report(solver.solve(self.max_iter))
        coef = solver.alpha[:l] - solver.alpha[l:]
        w = weighted_rows(func.X, coef).cpu().numpy()      # svr.py:89-90, telescoped (b2svm_weighted_rows)
        self.coef_ = (w, solver.calculate_rho())
        self._fitted = _Fitted(solver, func, func.kernel, coef)
        return self

    def predict(self, X):
        return self.decision_function(X)

    def score(self, X, y):
        return r2_score(y, self.predict(X))


class KernelSVR(BiKernelSVC):
    """epsilon-SVR (svr.py:110-204)."""
    _name = "KernelSVR"

    def __init__(self, C: int = 1., eps: float = 0., kernel: str = 'rbf', degree: float = 3, gamma: float = 'scale',
                 coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256) -> None:
        super().__init__(C, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size)
        self.eps = eps

    def fit(self, X, y):
        X, z = np.array(X), np.array(y)
        l, self.n_features = X.shape
        from .. import sharded
        if sharded.active_for(l) and not self.rff:
            from ..solver import host_std
            group = sharded._ENABLED["group"]
            spec, _ = self.register_kernel(host_std(X))
            solver, shard, stats = sharded.fit_svr_sharded(X, z, self.C, self.eps, spec, self.tol, self.max_iter, group)
            self._report(stats)
            nl = shard.n_local
            self._fitted = sharded.ShardedFit(solver, shard, solver.alpha[:nl] - solver.alpha[nl:], True, group)
            self.alpha_ = sharded.gather_variables(solver.alpha.cpu().numpy(), shard, True, group)
            self.rho_ = solver.calculate_rho()
            return self
        y, p = _svr_vectors(z, l, self.eps)
        func, spec, rff = self._columns(X, y, mirror=True)
        solver = SolverWithCache(p, y, self.C, self.tol, self.cache_size, func=func, max_iter_hint=self.max_iter)
        self._report(solver.solve(self.max_iter))
        self._finish(solver, func, spec, rff, X, solver.alpha[:l] - solver.alpha[l:])
        return self

    def predict(self, X):
        return self.decision_function(np.array(X))

    def score(self, X, y):
        return r2_score(y, self.predict(X))


class NuSVR(KernelSVR):
    """nu-SVR (svr.py:207-307): p = [-z, z], e'alpha = C*l*nu."""
    _name = "NuSVR"

    def __init__(self, C: float = 1., nu: float = 0.5, kernel: str = 'rbf', degree: float = 3,
                 gamma: float = 'scale', coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000,
                 tol: float = 1e-5, cache_size: int = 256) -> None:
        super().__init__(C, 0, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size)
        self.nu = nu

    def fit(self, X, y):
        X, z = np.array(X), np.array(y)
        l, self.n_features = X.shape
        from .. import sharded
        if sharded.active_for(l) and not self.rff:
            from ..solver import host_std
            group = sharded._ENABLED["group"]
            spec, _ = self.register_kernel(host_std(X))
            solver, shard, stats = sharded.fit_nusvr_sharded(X, z, self.C, self.nu, spec, self.tol, self.max_iter, group)
            self._report(stats)
            nl = shard.n_local
            self._fitted = sharded.ShardedFit(solver, shard, solver.alpha[:nl] - solver.alpha[nl:], True, group)
            self.alpha_ = sharded.gather_variables(solver.alpha.cpu().numpy(), shard, True, group)
            self.rho_, self.b_ = solver.calculate_rho_b()
            return self
        y = np.empty(2 * l)
        y[:l], y[l:] = 1, -1
        p = np.empty(2 * l)
        p[:l], p[l:] = -z, z
        func, spec, rff = self._columns(X, y, mirror=True)
        solver = NuSolverWithCache(p, y, self.C * l * self.nu, self.C, func, self.tol, self.cache_size, max_iter_hint=self.max_iter)
        self._report(solver.solve(self.max_iter))
        self._finish(solver, func, spec, rff, X, solver.alpha[:l] - solver.alpha[l:])
        self.rho_, self.b_ = solver.calculate_rho_b()
        return self

    def _finish(self, solver, func, spec, rff, X, coef):
        from ..solver import to_device_matrix
        dense = getattr(self, "_X_dense", None)
        self._fitted = _Fitted(solver, func, spec if dense is not None else func.kernel, coef, rff=rff,
                               X_raw=to_device_matrix(X[:1], func.device) if rff is not None else None,
                               Z=getattr(self, "_Z", None), X_train=dense)
        self.alpha_ = solver.alpha.cpu().numpy()

    def decision_function(self, x):
        return self._fitted.kernel_sum(x) + self.b_
