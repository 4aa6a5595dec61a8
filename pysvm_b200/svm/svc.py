"""Classification estimators with PySVM's API (reference ``pysvm/svm/svc.py``), solved on the GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
from sklearn.base import BaseEstimator
from sklearn.metrics import accuracy_score
from sklearn.multiclass import OneVsOneClassifier, OneVsRestClassifier

from .._native import B2SVMError
from ..rff import NormalRFF, rff_decision_device
from ..solver import (KernelSpec, host_std, NuSolverWithCache, QColumns, SolverWithCache, default_device, kernel_matrix,
                      kernel_matvec, solve_batched, to_device_matrix, weighted_rows)


# ---- batching of small independent sub-problems (one-vs-one / one-vs-rest) --------------------------------
_PENDING = None                  # list of (estimator, solver, finish) while a multiclass wrapper is fitting
BATCH_MAX_ROWS = int(os.environ.get("B2SVM_BATCH_MAX_ROWS", "4096"))


class deferred_solves:
    """While active, ``Bi*.fit`` on a small problem uploads its data and creates its device solver but
    leaves the SMO loop to ``run()``, which solves all collected sub-problems in one persistent launch
    (one CTA each).  sklearn's OneVsOne / OneVsRest code stays in charge of everything else."""

    def __enter__(self):
        global _PENDING
        self._saved, _PENDING = _PENDING, []
        self.items = _PENDING
        return self

    def __exit__(self, *exc):
        global _PENDING
        _PENDING = self._saved

    def run(self, max_iter):
        if not self.items:
            return
        from .. import sharded
        if sharded.active():
            return self._run_across_gpus(max_iter, sharded)
        stats = solve_batched([it[1] for it in self.items], max_iter)
        for (est, solver, finish), st in zip(self.items, stats):
            finish(st)

    def _run_across_gpus(self, max_iter, sharded):
        """The independent sub-problems of sklearn.multiclass scheduled over the GPUs of the process group with no
        communication during the solves: longest-processing-time assignment by row count (the same on
        every rank), one batched launch per GPU, then one all-gather of (stats, alpha, gradient) so that every rank
        ends with every fitted sub-estimator."""
        import torch.distributed as dist
        from .. import _native as N
        group = sharded._ENABLED["group"]
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        sizes = [int(it[1].alpha.shape[0]) for it in self.items]
        owner = sharded.schedule_subproblems(sizes, world)
        mine = [k for k in range(len(sizes)) if owner[k] == rank]
        stats = solve_batched([self.items[k][1] for k in mine], max_iter) if mine else []
        nstat = C.sizeof(N.SolveStats) // 8
        results = {}
        for k, st in zip(mine, stats):
            solver = self.items[k][1]
            results[k] = (np.frombuffer(bytes(st), dtype=np.float64), solver.alpha.cpu().numpy(),
                          solver.neg_y_grad.cpu().numpy())
        everything = sharded.exchange_subproblem_results(results, sizes, owner, nstat, group)
        for k, (est, solver, finish) in enumerate(self.items):
            words, alpha, grad = everything[k]
            st = N.SolveStats.from_buffer_copy(words.tobytes())
            if owner[k] != rank:
                solver.alpha = alpha
                solver.neg_y_grad = grad
                solver.last_stats = st
            finish(st)


class _Fitted:
    """What the reference keeps alive inside the ``decision_function`` closure."""

    def __init__(self, solver, func, spec, coef, rff=None, X_raw=None, Z=None, X_train=None):
        self.solver, self.func, self.spec, self.coef, self.rff, self.X_raw = solver, func, spec, coef, rff, X_raw
        self.Z = Z            # random Fourier features of the training rows (rff=True)
        self.X_train = X_train if X_train is not None else func.X     # (dense-Q mode: func.X is the kernel matrix)
        self._wz = None
        solver.release_cache()        # the fit is over: the solver stays (rho, state), its column cache does not

    def kernel_sum(self, x) -> np.ndarray:
        """coef . K(X_train, x) as float64 NumPy (svc.py:269-272 without the n x m matrix)."""
        dev = self.func.device
        if self.rff is not None:
            # K(a, b) = z(a).z(b)  =>  coef.K(X, x) = z(x).(Z^T coef)
            Xq = to_device_matrix(x if isinstance(x, torch.Tensor) else np.asarray(x), dev).to(self.X_raw.dtype)
            if self._wz is None:
                self._wz = weighted_rows(self.Z, self.coef)           # Z^T coef (b2svm_weighted_rows)
            w, b = self.rff._device_params(dev)
            return rff_decision_device(Xq, w, b, self._wz).cpu().numpy()
        Xq = to_device_matrix(x if isinstance(x, torch.Tensor) else np.asarray(x), dev).to(self.X_train.dtype)
        return kernel_matvec(self.spec, self.X_train, self.coef, Xq).cpu().numpy()


class BiLinearSVC(BaseEstimator):
    """Binary linear C-SVC (reference svc.py:10-94): the linear kernel on the same device solver,
    with the primal weights recovered as w = sum_i alpha_i y_i x_i."""

    def __init__(self, C: float = 1., max_iter: int = 1000, tol: float = 1e-5, cache_size: int = 256) -> None:
        super().__init__()
        self.C = C
        self.max_iter = max_iter
        self.tol = tol
        self.cache_size = cache_size

    _name = "LinearSVC"

    def _report(self, stats):
        self.n_iter_ = int(stats.n_iter)
        self.solve_stats_ = stats.as_dict()
        if stats.n_iter >= self.max_iter:      # the for-loop ran out (svc.py:77-79, 265-267)
            print("{} not coverage with {} iterations".format(self._name, self.max_iter))

    def fit(self, X, y):
        X, y = np.array(X), np.array(y, dtype=float)
        y[y != 1] = -1
        l, self.n_features = X.shape
        p = -np.ones(l)
        func = QColumns(X, y, KernelSpec("linear"))
        solver = SolverWithCache(p, y, self.C, self.tol, self.cache_size, func=func, max_iter_hint=self.max_iter)

        def finish(stats):
            self._report(stats)
            # svc.py:76 accumulates w += delta_i y_i x_i + delta_j y_j x_j; the sum telescopes to X^T (alpha*y)
            w = weighted_rows(func.X, solver.alpha * func.y).cpu().numpy()
            self.coef_ = (w, solver.calculate_rho())
            self._fitted = _Fitted(solver, func, func.kernel, solver.alpha * func.y)

        if _PENDING is not None and l <= BATCH_MAX_ROWS:
            _PENDING.append((self, solver, finish))
        else:
            finish(solver.solve(self.max_iter))
        return self

    def decision_function(self, X) -> np.ndarray:
        return np.matmul(self.coef_[0], np.array(X).T) - self.coef_[-1]

    def predict(self, X) -> np.ndarray:
        return (self.decision_function(np.array(X)) >= 0).astype(int)

    def score(self, X, y) -> float:
        return accuracy_score(y, self.predict(X))


class LinearSVC(BiLinearSVC):
    """Multiclass wrapper (svc.py:97-158).  ``n_jobs`` is accepted but sub-problems are not farmed
    out to worker processes: they run on the GPU of this process."""

    def __init__(self, C: float = 1., max_iter: int = 1000, tol: float = 1e-5, cache_size: int = 256,
                 multiclass: str = "ovr", n_jobs=None) -> None:
        super().__init__(C, max_iter, tol, cache_size)
        self.multiclass = multiclass
        self.n_jobs = n_jobs
        params = {"estimator": BiLinearSVC(C, max_iter, tol, cache_size), "n_jobs": None}
        self.multiclass_model = {"ovo": OneVsOneClassifier(**params), "ovr": OneVsRestClassifier(**params)}[multiclass]

    def fit(self, X, y):
        # sklearn builds the sub-problems (svc.py:134-149); their SMO loops run together on the GPU
        with deferred_solves() as batch:
            self.multiclass_model.fit(X, y)
        batch.run(self.max_iter)
        return self

    def decision_function(self, X):
        return self.multiclass_model.decision_function(X)

    def predict(self, X):
        return self.multiclass_model.predict(X)

    def score(self, X, y):
        return self.multiclass_model.score(X, y)


class BiKernelSVC(BiLinearSVC):
    """Binary kernel C-SVC (svc.py:161-279)."""
    _name = "KernelSVC"

    def __init__(self, C: float = 1., kernel: str = 'rbf', degree: float = 3, gamma: str = 'scale', coef0: float = 0,
                 max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256, wss: str = "mvp") -> None:
        """``wss`` is an additive, opt-in keyword (not in the reference): "mvp" = the reference's maximal violating
        pair (default, parity path); "wss3" = LIBSVM's second-order working-set selection."""
        super().__init__(C, max_iter, tol, cache_size)
        self.wss = wss
        self.kernel = kernel
        self.gamma = gamma
        self.degree = degree
        self.coef0 = coef0
        self.rff = rff
        self.D = D

    def register_kernel(self, std: float):
        """svc.py:210-241.  Returns (KernelSpec, NormalRFF or None).  With ``rff=True`` the RBF kernel
        is replaced by z(a).z(b): the features are computed once and the solver runs a linear kernel on
        them (the reference recomputes both transforms in every kernel call)."""
        if type(self.gamma) == str:
            gamma = {'scale': 1 / (self.n_features * std), 'auto': 1 / self.n_features}[self.gamma]
        else:
            gamma = self.gamma
        rff = None
        if self.rff:          # drawn whatever the kernel is, exactly like the reference (svc.py:225-226)
            rff = NormalRFF(gamma, self.D).fit(np.ones((1, self.n_features)))
        self.gamma_ = gamma
        spec = KernelSpec(self.kernel, gamma, self.coef0, self.degree)
        if self.kernel not in ("linear", "poly", "rbf", "sigmoid"):
            raise KeyError(self.kernel)
        return spec, (rff if self.kernel == "rbf" else None)

    def _columns(self, X, y, mirror=False):
        """Device Q-column source for this estimator's kernel (the reference's ``func``)."""
        spec, rff = self.register_kernel(host_std(X))
        self._Z = None
        self._X_dense = None
        if rff is not None:
            Z = rff.transform_device(to_device_matrix(X, default_device()))       # n x D float64, like rff.py:20
            self._Z = Z
            if Z.shape[1] <= 256:
                return QColumns(Z, y, KernelSpec("linear"), mirror=mirror), spec, rff
            # wider feature maps (default D = 1000): the kernel z(a).z(b) is handed over as a precomputed
            # n x n matrix (b2svm_kernel_matrix), which the solver reads like a completely filled column cache
            n = Z.shape[0]
            self._check_dense(n, 8)
            K = kernel_matrix(KernelSpec("linear"), Z)
            return QColumns(K, y, KernelSpec("precomputed"), mirror=mirror, cache_bytes=0), spec, rff
        if self.cache_size == 0:
            # the reference's dense-Q mode (svc.py:251-253, 419-421; oc_svm.py:85-89): the whole kernel matrix is built
            # once -- on the device (b2svm_kernel_matrix) -- and every SMO iteration reads two of its columns
            Xd = to_device_matrix(X, default_device())
            self._check_dense(Xd.shape[0], Xd.element_size())
            K = kernel_matrix(spec, Xd)
            self._X_dense = Xd
            return QColumns(K, y, KernelSpec("precomputed"), mirror=mirror, cache_bytes=0), spec, None
        return QColumns(X, y, spec, mirror=mirror), spec, None

    @staticmethod
    def _check_dense(n, esz):
        free, _ = torch.cuda.mem_get_info(default_device())
        if n * n * esz > 0.6 * free:
            raise MemoryError(f"the dense {n} x {n} kernel matrix needs {n * n * esz / 2**30:.1f} GiB of device memory; "
                              f"use the column cache (cache_size > 0) for this many rows")

    def fit(self, X, y):
        X, y = np.asarray(X), np.array(y, dtype=float)     # X is never written: no host copy needed
        y[y != 1] = -1
        l, self.n_features = X.shape
        p = -np.ones(l)
        from .. import sharded
        if sharded.active_for(l) and not self.rff and not (_PENDING is not None and l <= BATCH_MAX_ROWS):
            return self._fit_sharded(X, y, sharded)
        func, spec, rff = self._columns(X, y)
        solver = SolverWithCache(p, y, self.C, self.tol, self.cache_size, func=func, max_iter_hint=self.max_iter)

        def finish(stats):
            self._report(stats)
            self._finish(solver, func, spec, rff, X, solver.alpha * func.y)

        wss = getattr(self, "wss", "mvp")
        if _PENDING is not None and l <= BATCH_MAX_ROWS and wss == "mvp":
            _PENDING.append((self, solver, finish))       # solved together with its sibling sub-problems
        else:
            finish(solver.solve(self.max_iter, wss=wss))
        return self

    def _finish(self, solver, func, spec, rff, X, coef):
        dense = getattr(self, "_X_dense", None)
        self._fitted = _Fitted(solver, func, spec if dense is not None else func.kernel, coef, rff=rff,
                               X_raw=to_device_matrix(X[:1], func.device) if rff is not None else None,
                               Z=getattr(self, "_Z", None), X_train=dense)
        try:
            self.rho_ = solver.calculate_rho()
        except B2SVMError:
            # e.g. a single-class problem: the reference's fit() succeeds and calculate_rho() only raises
            # (ValueError: min of an empty set, solver.py:175) once decision_function is called
            self.rho_ = None
        a = solver.alpha.cpu().numpy()
        self.alpha_ = a
        self.support_ = np.flatnonzero(a > 0)

    def _rho(self):
        if self.rho_ is None:
            raise ValueError("calculate_rho: zero-size bound set (no free support vector and an empty bound class)")
        return self.rho_

    def _fit_sharded(self, X, y, sharded):
        """The same fit, row-sharded over the GPUs of the current process group (sharded.enable())."""
        group = sharded._ENABLED["group"]
        spec, _ = self.register_kernel(host_std(X))
        solver, shard, stats = sharded.fit_csvc_sharded(X, (y == 1).astype(int), self.C, spec, self.tol,
                                                        self.max_iter, group)
        self._report(stats)
        self._fitted = sharded.ShardedFit(solver, shard, solver.alpha * solver._engine.func.y, False, group)
        self.rho_ = solver.calculate_rho()
        self.alpha_ = sharded.gather_variables(solver.alpha.cpu().numpy(), shard, False, group)
        self.support_ = np.flatnonzero(self.alpha_ > 0)
        return self

    def decision_function(self, x) -> np.ndarray:
        return self._fitted.kernel_sum(x) - self._rho()

    def predict(self, X) -> np.ndarray:
        return super().predict(X)

    def score(self, X, y) -> float:
        return super().score(X, y)


class KernelSVC(LinearSVC, BiKernelSVC):
    """Multiclass kernel C-SVC (svc.py:282-354): sklearn's OvO / OvR wrappers around BiKernelSVC."""

    def __init__(self, C: float = 1., kernel: str = 'rbf', degree: float = 3, gamma: float = 'scale',
                 coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256, multiclass: str = "ovr", n_jobs: int = None) -> None:
        LinearSVC.__init__(self, C, max_iter, tol, cache_size, multiclass, n_jobs)
        self.kernel = kernel
        self.gamma = gamma
        self.degree = degree
        self.coef0 = coef0
        self.rff = rff
        self.D = D
        params = {"estimator": BiKernelSVC(C, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size),
                  "n_jobs": None}
        self.multiclass_model = {"ovo": OneVsOneClassifier(**params), "ovr": OneVsRestClassifier(**params)}[multiclass]

    def fit(self, X, y):
        return LinearSVC.fit(self, X, y)

    def decision_function(self, X):
        return LinearSVC.decision_function(self, X)

    def predict(self, X):
        return LinearSVC.predict(self, X)

    def score(self, X, y):
        return LinearSVC.score(self, X, y)


class BiNuSVC(BiKernelSVC):
    """Binary nu-SVC (svc.py:357-446): p = 0, C = 1, e'alpha = nu*l."""
    _name = "NuSVC"

    def __init__(self, nu: float = 0.5, kernel: str = 'rbf', degree: float = 3, gamma: float = 'scale',
                 coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256) -> None:
        super().__init__(1, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size)
        self.nu = nu

    def fit(self, X, y):
        X, y = np.array(X), np.array(y, dtype=float)
        y[y != 1] = -1
        l, self.n_features = X.shape
        from .. import sharded
        if sharded.active_for(l) and not self.rff and not (_PENDING is not None and l <= BATCH_MAX_ROWS):
            group = sharded._ENABLED["group"]
            spec, _ = self.register_kernel(host_std(X))
            solver, shard, stats = sharded.fit_nusvc_sharded(X, (y == 1).astype(int), self.nu, spec, self.tol, self.max_iter, group)
            self._report(stats)
            self._fitted = sharded.ShardedFit(solver, shard, solver.alpha * solver._engine.func.y, False, group)
            self.alpha_ = sharded.gather_variables(solver.alpha.cpu().numpy(), shard, False, group)
            self.support_ = np.flatnonzero(self.alpha_ > 0)
            self.rho_, self.b_ = solver.calculate_rho_b()
            return self
        p = np.zeros(l)
        func, spec, rff = self._columns(X, y)
        solver = NuSolverWithCache(p, y, self.nu * l, self.C, func, self.tol, self.cache_size, max_iter_hint=self.max_iter)

        def finish(stats):
            self._report(stats)
            self._finish(solver, func, spec, rff, X, solver.alpha * func.y)
            self.rho_, self.b_ = solver.calculate_rho_b()

        if _PENDING is not None and l <= BATCH_MAX_ROWS:
            _PENDING.append((self, solver, finish))
        else:
            finish(solver.solve(self.max_iter))
        return self

    def decision_function(self, x) -> np.ndarray:
        rho, b = self.rho_, self.b_
        return self._fitted.kernel_sum(x) / rho + b / rho

    def _finish(self, solver, func, spec, rff, X, coef):
        dense = getattr(self, "_X_dense", None)
        self._fitted = _Fitted(solver, func, spec if dense is not None else func.kernel, coef, rff=rff,
                               X_raw=to_device_matrix(X[:1], func.device) if rff is not None else None,
                               Z=getattr(self, "_Z", None), X_train=dense)
        a = solver.alpha.cpu().numpy()
        self.alpha_ = a
        self.support_ = np.flatnonzero(a > 0)


class NuSVC(KernelSVC, BiNuSVC):
    """Multiclass nu-SVC (svc.py:449-514)."""

    def __init__(self, nu: float = 0.5, kernel: str = 'rbf', degree: float = 3, gamma: float = 'scale',
                 coef0: float = 0., max_iter: int = 1000, rff: bool = False, D: int = 1000, tol: float = 1e-5,
                 cache_size: int = 256, multiclass: str = "ovr", n_jobs: int = None) -> None:
        KernelSVC.__init__(self, 1, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size, multiclass,
                           n_jobs)
        self.nu = nu
        params = {"estimator": BiNuSVC(nu, kernel, degree, gamma, coef0, max_iter, rff, D, tol, cache_size),
                  "n_jobs": None}
        self.multiclass_model = {"ovo": OneVsOneClassifier(**params), "ovr": OneVsRestClassifier(**params)}[multiclass]
