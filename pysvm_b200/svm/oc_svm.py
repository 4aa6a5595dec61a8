"""One-class SVM with PySVM's API (reference ``pysvm/svm/oc_svm.py``)."""
from __future__ import annotations

import numpy as np

from ..solver import SolverWithCache
from .svc import BiNuSVC


def oneclass_initial_alpha(nu: float, l: int) -> np.ndarray:
    """oc_svm.py:76-83, including the leaked loop variable: with n = int(nu*l), alpha[:n-1] = 1,
    alpha[n-1] = nu*l - n, alpha[n] = 1, alpha[n+1:] = 0 (sum = nu*l).  n == 0 raises, as there."""
    n = int(nu * l)
    if n == 0:
        raise UnboundLocalError("int(nu*l) == 0: the reference fails here too (oc_svm.py:80)")
    alpha = np.ones(l)
    if n < l:
        alpha[n - 1] = nu * l - n
    alpha[n + 1:] = 0
    return alpha


class OneClassSVM(BiNuSVC):
    _name = "OneClassSVM"

    def __init__(self, nu: float = 0.5, kernel: str = 'rbf', degree: float = 3, gamma: float = 'scale',
                 coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256) -> None:
        super().__init__(nu, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size)

    def fit(self, X):
        X = np.array(X)
        l, self.n_features = X.shape
        from .. import sharded
        if sharded.active_for(X.shape[0]) and not self.rff:
            return self._fit_oneclass_sharded(X, sharded)
        p = np.zeros(l)
        y = np.ones(l)
        func, spec, rff = self._columns(X, y)
        solver = SolverWithCache(p, y, 1, self.tol, self.cache_size, func=func, max_iter_hint=self.max_iter)
        solver.alpha = oneclass_initial_alpha(self.nu, l)          # oc_svm.py:76-83, 93
        solver._engine.init_gradient(None)                         # oc_svm.py:94-95: g = -K alpha
        self._report(solver.solve(self.max_iter))
        self._finish(solver, func, spec, rff, X, solver.alpha.clone())
        self.rho_ = solver.calculate_rho()
        return self

    def _fit_oneclass_sharded(self, X, sharded):
        """The same fit, row-sharded over the GPUs of the current process group (sharded.enable())."""
        from ..solver import host_std
        group = sharded._ENABLED["group"]
        spec, _ = self.register_kernel(host_std(X))
        solver, shard, stats = sharded.fit_oneclass_sharded(X, self.nu, spec, self.tol, self.max_iter, group)
        self._report(stats)
        self._fitted = sharded.ShardedFit(solver, shard, solver.alpha.clone(), False, group)
        self.alpha_ = sharded.gather_variables(solver.alpha.cpu().numpy(), shard, False, group)
        self.support_ = np.flatnonzero(self.alpha_ > 0)
        self.rho_ = solver.calculate_rho()
        return self

    def decision_function(self, x):
        return self._fitted.kernel_sum(x) - self.rho_

    def predict(self, X):
        pred = np.sign(self.decision_function(X))
        pred[pred == 0] = -1
        return pred

    def score(self, X, y):
        raise NotImplementedError
