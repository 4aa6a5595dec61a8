"""CPU oracle for the SMO solve path of PySVM -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm (Kaslanarian/PySVM 0.2,
``pysvm/solver.py``, ``pysvm/svm/{svc,svr,oc_svm}.py``, ``pysvm/rff.py``).  It is the checker
the CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; nothing under
``pysvm_b200/`` does, and the product has no CPU fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference (imported
from ``/root/reference`` in the build container) and stores its answers under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks that this restatement reproduces them (iteration counts and
working-pair traces exactly, alpha / rho / decision values to 1e-12).  The floating-point
operation order of the reference is kept on purpose (same NumPy calls in the same order) so the
agreement is bit-level on the same NumPy/OpenBLAS build.

The layout is deliberately different from the reference (one flat state object + free functions
instead of a class hierarchy with callbacks) -- it restates *what* is computed, citing the
reference line for every rule.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# kernels  (reference: pysvm/svm/svc.py:210-241, pysvm/rff.py:12-20)
# --------------------------------------------------------------------------------------


@dataclass
class KernelSpec:
    kind: str = "rbf"            # linear | poly | rbf | sigmoid
    gamma: float = 1.0
    coef0: float = 0.0
    degree: float = 3
    rff_w: Optional[np.ndarray] = None   # (D, d) float64, already scaled by sqrt(2*gamma)
    rff_b: Optional[np.ndarray] = None   # (D,)


def resolve_gamma(gamma, n_features: int, std: float) -> float:
    """svc.py:217-223 -- 'scale' uses the *standard deviation* (not the variance)."""
    if isinstance(gamma, str):
        if gamma == "scale":
            return 1 / (n_features * std)
        if gamma == "auto":
            return 1 / n_features
        raise KeyError(gamma)
    return gamma


def draw_rff(gamma: float, D: int, n_features: int) -> Tuple[np.ndarray, np.ndarray]:
    """rff.py:14-16 -- global NumPy RNG, randn(D, d) first, then uniform(0, 2pi, D)."""
    w = np.sqrt(gamma * 2) * np.random.randn(D, n_features)
    b = np.random.uniform(0, 2 * np.pi, D)
    return w, b


def rff_features(X: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """rff.py:19-20."""
    D = w.shape[0]
    return np.sqrt(2 / D) * np.cos(np.matmul(X, w.T) + b)


def make_kernel(spec: KernelSpec) -> Callable[[np.ndarray, np.ndarray], np.ndarray]:
    """Dense kernel block K(A, B) with the reference's operation order (svc.py:225-241)."""
    g, c0, deg = spec.gamma, spec.coef0, spec.degree
    if spec.kind == "linear":
        return lambda A, B: np.matmul(A, B.T)
    if spec.kind == "poly":
        return lambda A, B: (g * np.matmul(A, B.T) + c0) ** deg
    if spec.kind == "sigmoid":
        return lambda A, B: np.tanh(g * np.matmul(A, B.T) + c0)
    if spec.kind == "rbf":
        if spec.rff_w is not None:
            w, b = spec.rff_w, spec.rff_b
            return lambda A, B: np.matmul(rff_features(A, w, b), rff_features(B, w, b).T)

        def rbf(A, B):
            sq = (A ** 2).sum(1, keepdims=True) + (B ** 2).sum(1) - 2 * np.matmul(A, B.T)
            return np.exp(-g * sq)

        return rbf
    raise KeyError(spec.kind)


def kernel_spec_for(X: np.ndarray, kernel="rbf", gamma="scale", coef0=0.0, degree=3,
                    rff=False, D=1000) -> KernelSpec:
    """What ``register_kernel(X.std())`` resolves to for a given training matrix (svc.py:249)."""
    g = resolve_gamma(gamma, X.shape[1], X.std())
    spec = KernelSpec(kind=kernel, gamma=g, coef0=coef0, degree=degree)
    if rff:
        spec.rff_w, spec.rff_b = draw_rff(g, D, X.shape[1])
    return spec


# --------------------------------------------------------------------------------------
# Q columns.  In every estimator Q[t, i] = s_t * K(row(t), row(i)) * s_i with s = y and
# row(t) = t mod l  (svc.py:257-258, svr.py:174-179 / 271-276, oc_svm.py:72-73).
# --------------------------------------------------------------------------------------


class ColumnSource:
    """Lazy Q columns behind a 256-entry LRU (solver.py:207,227-229: maxsize is frozen at 256)."""

    def __init__(self, X, y, kfun, mirror: bool, lru: int = 256):
        self.X, self.y, self.kfun, self.mirror = X, y, kfun, mirror
        self.l = X.shape[0]
        self.lru = lru
        self._memo: "OrderedDict[int, np.ndarray]" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def _evaluate(self, i: int) -> np.ndarray:
        X, y, l = self.X, self.y, self.l
        if self.mirror:
            # svr.py:174-179 -- the sign comes first, then the vertical stack [Qi, -Qi]
            if i < l:
                qi = self.kfun(X, X[i:i + 1]).flatten()
            else:
                qi = -self.kfun(X, X[i - l:i - l + 1]).flatten()
            return np.r_[qi, -qi]
        return y * self.kfun(X, X[i:i + 1]).flatten() * y[i]

    def __call__(self, i: int) -> np.ndarray:
        i = int(i)
        if self.lru > 0 and i in self._memo:
            self._memo.move_to_end(i)
            self.hits += 1
            return self._memo[i]
        self.misses += 1
        col = self._evaluate(i)
        if self.lru > 0:
            self._memo[i] = col
            if len(self._memo) > self.lru:
                self._memo.popitem(last=False)
        return col


class OneClassColumns(ColumnSource):
    """oc_svm.py:72-73 -- plain K column (y == 1 everywhere)."""

    def _evaluate(self, i: int) -> np.ndarray:
        return self.kfun(self.X, self.X[i:i + 1]).flatten()


class LinearColumns(ColumnSource):
    """svc.py:67-68 / svr.py:73-78 -- BiLinearSVC / LinearSVR use a matrix-vector product."""

    def _evaluate(self, i: int) -> np.ndarray:
        X, y, l = self.X, self.y, self.l
        if self.mirror:
            qi = np.matmul(X, X[i]) if i < l else -np.matmul(X, X[i - l])
            return np.r_[qi, -qi]
        return y * np.matmul(X, X[i]) * y[i]


# --------------------------------------------------------------------------------------
# dual state and the SMO rules
# --------------------------------------------------------------------------------------


@dataclass
class DualState:
    y: np.ndarray
    p: np.ndarray
    C: float
    tol: float
    alpha: np.ndarray
    g: np.ndarray                 # "neg_y_grad" = -y * grad f(alpha)
    QD: Optional[np.ndarray] = None
    trace: Optional[List[Tuple[int, int, float, float]]] = None


def fresh_state(p, y, C, tol) -> DualState:
    """solver.py:31-45."""
    return DualState(y=y, p=p, C=C, tol=tol, alpha=np.zeros(p.shape[0]), g=-y * p)


def _masks(st: DualState):
    """solver.py:55-64 -- strict inequalities."""
    below, above = st.alpha < st.C, st.alpha > 0
    pos, neg = st.y > 0, st.y < 0
    up = (below & pos) | (above & neg)
    low = (below & neg) | (above & pos)
    return up, low, pos, neg


def _first_argmax(values: np.ndarray, mask: np.ndarray) -> int:
    """np.argmax over an index-sorted subset returns the lowest index among ties (solver.py:68)."""
    if not mask.any():
        return -1
    idx = np.flatnonzero(mask)
    return int(idx[np.argmax(values[idx])])


def _first_argmin(values: np.ndarray, mask: np.ndarray) -> int:
    if not mask.any():
        return -1
    idx = np.flatnonzero(mask)
    return int(idx[np.argmin(values[idx])])


def select_pair(st: DualState) -> Tuple[int, int]:
    """First-order maximal-violating pair (solver.py:47-75); (-1,-1) when a set is empty or
    g_i - g_j < tol."""
    up, low, _, _ = _masks(st)
    i = _first_argmax(st.g, up)
    j = _first_argmin(st.g, low)
    if i < 0 or j < 0 or st.g[i] - st.g[j] < st.tol:
        return -1, -1
    return i, j


def select_pair_wss3(st: DualState, columns, kdiag: np.ndarray) -> Tuple[int, int]:
    """OPT-IN rule, NOT in the reference (its Solver takes the maximal violating pair, solver.py:47-75): LIBSVM's
    second-order working-set selection WSS 3 (Fan, Chen, Lin, JMLR 6, 2005, Algorithm 2; libsvm svm.cpp
    Solver::select_working_set), restated on the reference's quantities: g = neg_y_grad, I_up / I_low as in
    solver.py:55-64.  i = argmax_{I_up} g; j = argmin_{t in I_low, g_t < g_i} -(g_i - g_t)^2 / a_it with
    a_it = K_ii + K_tt - 2 K_it (<= 0 -> 1e-12); stop like solver.py:72-75.  Ties: the lower index.
    ``columns(i)`` is the Q column (y_t y_i K_ti); ``kdiag`` the kernel diagonal per variable."""
    up, low, _, _ = _masks(st)
    if not up.any() or not low.any():
        return -1, -1
    i = _first_argmax(st.g, up)
    j0 = _first_argmin(st.g, low)
    if st.g[i] - st.g[j0] < st.tol:
        return -1, -1
    K_i = columns(i) * st.y * st.y[i]                     # K(x_t, x_i) for every variable t
    a = (kdiag[i] + kdiag) - 2.0 * K_i
    a = np.where(a <= 0, 1e-12, a)
    b = st.g[i] - st.g
    cand = low & (st.g < st.g[i])
    gain = np.where(cand, -(b * b) / a, np.inf)
    j = int(np.argmin(gain))                              # first minimum = lower index
    return (int(i), j) if cand.any() else (-1, -1)


def _clip_equal_sign(st: DualState, i: int, j: int, delta: float):
    """The y_i*y_j == +1 box clipping, in the reference's order (solver.py:120-142, 349-371)."""
    a, C = st.alpha, st.C
    total = a[i] + a[j]
    a[i] -= delta
    a[j] += delta
    if total > C:
        if a[i] > C:
            a[i] = C
            a[j] = total - C
    else:
        if a[j] < 0:
            a[j] = 0
            a[i] = total
    if total > C:
        if a[j] > C:
            a[j] = C
            a[i] = total - C
    else:
        if a[i] < 0:
            a[i] = 0
            a[j] = total


def take_step(st: DualState, i: int, j: int, Qi: np.ndarray, Qj: np.ndarray):
    """Two-variable analytic step + rank-2 gradient update (solver.py:82-147)."""
    a, C = st.alpha, st.C
    yi, yj = st.y[i], st.y[j]
    ai0, aj0 = a[i], a[j]
    quad = Qi[i] + Qj[j] - 2 * yi * yj * Qi[j]
    if quad <= 0:
        quad = 1e-12
    if yi * yj == -1:
        delta = (st.g[i] * yi + st.g[j] * yj) / quad
        diff = ai0 - aj0
        a[i] += delta
        a[j] += delta
        if diff > 0:
            if a[j] < 0:
                a[j] = 0
                a[i] = diff
        else:
            if a[i] < 0:
                a[i] = 0
                a[j] = -diff
        if diff > 0:
            if a[i] > C:
                a[i] = C
                a[j] = C - diff
        else:
            if a[j] > C:
                a[j] = C
                a[i] = C + diff
    else:
        delta = (st.g[j] * yj - st.g[i] * yi) / quad
        _clip_equal_sign(st, i, j, delta)
    di, dj = a[i] - ai0, a[j] - aj0
    st.g -= st.y * (di * Qi + dj * Qj)
    if st.trace is not None:
        st.trace.append((int(i), int(j), float(di), float(dj)))
    return di, dj


def bias(st: DualState) -> float:
    """solver.py:149-177."""
    free = (st.alpha > 0) & (st.alpha < st.C)
    if free.sum() > 0:
        return -np.average(st.g[free])
    at0, atC = st.alpha == 0, st.alpha == st.C
    ub = (at0 & (st.y < 0)) | (atC & (st.y > 0))
    lb = (at0 & (st.y > 0)) | (atC & (st.y < 0))
    return -(st.g[lb].min() + st.g[ub].max()) / 2


# ---- nu variant ------------------------------------------------------------------------


def nu_fill(y: np.ndarray, t: float) -> np.ndarray:
    """solver.py:269-278 / 456-466 -- per class, walk in index order handing out min(1., rest)
    starting from t/2 (the cap is a hard-coded 1, not C)."""
    rest = {True: t / 2, False: t / 2}
    alpha = np.zeros(y.shape[0])
    for k in range(y.shape[0]):
        key = bool(y[k] == 1)
        alpha[k] = min(1., rest[key])
        rest[key] -= alpha[k]
    return alpha


def nu_state(p, y, t, C, tol, columns: ColumnSource) -> DualState:
    """NuSolverWithCache.__init__ (solver.py:446-473): O(l^2) initial gradient through the
    column callback, QD = diagonal taken from each column."""
    n = p.shape[0]
    alpha = nu_fill(y, t)
    g = np.zeros(n)
    qd = np.empty(n)
    for k in range(n):
        col = columns(k)
        g[k] = col @ alpha + p[k]
        qd[k] = col[k]
    g *= -y
    return DualState(y=y, p=p, C=C, tol=tol, alpha=alpha, g=g, QD=qd)


def _gain(st: DualState, i: int, j: int, Qi: np.ndarray) -> float:
    """solver.py:324-329."""
    quad = st.QD[i] + st.QD[j] - 2 * Qi[j]
    if quad <= 0:
        quad = 1e-12
    return -(st.g[i] - st.g[j]) ** 2 / quad


def select_pair_nu(st: DualState, columns: ColumnSource):
    """solver.py:283-339: one first-order pair per class, then the pair with the smaller
    second-order objective change wins ('+' only when strictly smaller).  There is no tol test;
    (-1, ...) only when neither class offers a pair."""
    up, low, pos, neg = _masks(st)
    ip, jp = _first_argmax(st.g, up & pos), _first_argmin(st.g, low & pos)
    i_n, j_n = _first_argmax(st.g, up & neg), _first_argmin(st.g, low & neg)
    pos_ok = ip >= 0 and jp >= 0
    neg_ok = i_n >= 0 and j_n >= 0
    if not pos_ok and not neg_ok:
        return -1, -1, None, None
    if not pos_ok:
        return i_n, j_n, columns(i_n), columns(j_n)
    if not neg_ok:
        return ip, jp, columns(ip), columns(jp)
    Qip = columns(ip)
    gain_p = _gain(st, ip, jp, Qip)
    Qin = columns(i_n)
    gain_n = _gain(st, i_n, j_n, Qin)
    if gain_p < gain_n:
        return ip, jp, Qip, columns(jp)
    return i_n, j_n, Qin, columns(j_n)


def take_step_nu(st: DualState, i: int, j: int, Qi: np.ndarray, Qj: np.ndarray):
    """solver.py:341-376 (i and j always share a class)."""
    a = st.alpha
    ai0, aj0 = a[i], a[j]
    quad = Qi[i] + Qj[j] - 2 * Qi[j]
    if quad <= 0:
        quad = 1e-12
    delta = st.y[i] * (st.g[j] - st.g[i]) / quad
    _clip_equal_sign(st, i, j, delta)
    di, dj = a[i] - ai0, a[j] - aj0
    st.g -= st.y * (di * Qi + dj * Qj)
    if st.trace is not None:
        st.trace.append((int(i), int(j), float(di), float(dj)))
    return di, dj


def bias_nu(st: DualState) -> Tuple[float, float]:
    """solver.py:378-419 (free SV test uses a hard-coded upper bound of 1)."""
    grad = -st.y * st.g
    out = []
    for s in (1, -1):
        cls = st.y == s
        free = (st.alpha > 0) & (st.alpha < 1) & cls
        if free.sum() == 0:
            hi, lo = grad[(st.alpha == 1) & cls], grad[(st.alpha == 0) & cls]
            r = (hi.max() + lo.min()) / 2 if hi.size and lo.size else 0
        else:
            r = np.average(grad[free])
        out.append(r)
    r1, r2 = out
    return (r1 + r2) / 2, (r2 - r1) / 2


# --------------------------------------------------------------------------------------
# drivers: what each estimator's fit() hands to the solver
# --------------------------------------------------------------------------------------


@dataclass
class Fit:
    kind: str
    alpha: np.ndarray
    g: np.ndarray
    y: np.ndarray
    n_iter: int
    converged: bool
    rho: float
    b: float = 0.0
    coef: np.ndarray = None       # weights on K(X_train, .) in decision_function
    X: np.ndarray = None
    spec: KernelSpec = None
    trace: list = None
    w: np.ndarray = None          # primal weights (linear estimators only)
    lru_hits: int = 0
    lru_misses: int = 0
    extra: dict = field(default_factory=dict)

    def decision_function(self, Xq: np.ndarray) -> np.ndarray:
        Xq = np.array(Xq)
        if self.w is not None:                       # svc.py:84-86, svr.py:98-99
            return np.matmul(self.w, Xq.T) - self.rho
        s = np.matmul(self.coef, make_kernel(self.spec)(self.X, Xq))
        if self.kind in ("csvc", "svr", "oneclass"):  # svc.py:269-272, svr.py:191-194, oc_svm.py:108-111
            return s - self.rho
        if self.kind == "nusvc":                      # svc.py:436-439
            return s / self.rho + self.b / self.rho
        if self.kind == "nusvr":                      # svr.py:297-300
            return s + self.b
        raise KeyError(self.kind)

    def predict(self, Xq):
        dec = self.decision_function(Xq)
        if self.kind in ("csvc", "nusvc"):            # svc.py:88-90
            return (dec >= 0).astype(int)
        if self.kind == "oneclass":                   # oc_svm.py:114-118
            pred = np.sign(dec)
            pred[pred == 0] = -1
            return pred
        return dec


def _run(st: DualState, columns, max_iter: int):
    """svc.py:260-267 -- at most max_iter updates; n_iter = number of updates done."""
    for n_iter in range(max_iter):
        i, j = select_pair(st)
        if i < 0:
            return n_iter, True
        take_step(st, i, j, columns(i), columns(j))
    return max_iter, False


def _run_nu(st: DualState, columns, max_iter: int):
    """svc.py:426-433."""
    for n_iter in range(max_iter):
        i, j, Qi, Qj = select_pair_nu(st, columns)
        if i < 0:
            return n_iter, True
        take_step_nu(st, i, j, Qi, Qj)
    return max_iter, False


def _binary_labels(y) -> np.ndarray:
    y = np.array(y, dtype=float)      # svc.py:244-245
    y[y != 1] = -1
    return y


def fit_csvc(X, y, C=1., kernel="rbf", degree=3, gamma="scale", coef0=0., max_iter=1000,
             rff=False, D=1000, tol=1e-5, trace=False, spec: KernelSpec = None) -> Fit:
    """BiKernelSVC.fit, cached-solver path (svc.py:243-273)."""
    X, y = np.array(X), _binary_labels(y)
    l = X.shape[0]
    spec = spec or kernel_spec_for(X, kernel, gamma, coef0, degree, rff, D)
    cols = ColumnSource(X, y, make_kernel(spec), mirror=False)
    st = fresh_state(-np.ones(l), y, C, tol)
    st.trace = [] if trace else None
    n_iter, conv = _run(st, cols, max_iter)
    return Fit("csvc", st.alpha, st.g, y, n_iter, conv, bias(st), coef=st.alpha * y, X=X, spec=spec,
               trace=st.trace, lru_hits=cols.hits, lru_misses=cols.misses)


def fit_nusvc(X, y, nu=0.5, kernel="rbf", degree=3, gamma="scale", coef0=0., max_iter=1000,
              rff=False, D=1000, tol=1e-5, trace=False, spec: KernelSpec = None) -> Fit:
    """BiNuSVC.fit (svc.py:408-440): p = 0, C = 1, t = nu*l."""
    X, y = np.array(X), _binary_labels(y)
    l = X.shape[0]
    spec = spec or kernel_spec_for(X, kernel, gamma, coef0, degree, rff, D)
    cols = ColumnSource(X, y, make_kernel(spec), mirror=False)
    st = nu_state(np.zeros(l), y, nu * l, 1, tol, cols)
    st.trace = [] if trace else None
    n_iter, conv = _run_nu(st, cols, max_iter)
    rho, b = bias_nu(st)
    return Fit("nusvc", st.alpha, st.g, y, n_iter, conv, rho, b, coef=st.alpha * y, X=X, spec=spec,
               trace=st.trace, lru_hits=cols.hits, lru_misses=cols.misses)


def _svr_vectors(z, l, eps):
    y = np.empty(2 * l)
    y[:l], y[l:] = 1., -1.
    p = np.ones(2 * l) * eps          # svr.py:161-163
    p[:l] -= z
    p[l:] += z
    return y, p


def fit_svr(X, z, C=1., eps=0., kernel="rbf", degree=3, gamma="scale", coef0=0., max_iter=1000,
            rff=False, D=1000, tol=1e-5, trace=False, spec: KernelSpec = None) -> Fit:
    """KernelSVR.fit, cached-solver path (svr.py:154-196): 2l variables."""
    X, z = np.array(X), np.array(z)
    l = X.shape[0]
    y, p = _svr_vectors(z, l, eps)
    spec = spec or kernel_spec_for(X, kernel, gamma, coef0, degree, rff, D)
    cols = ColumnSource(X, y, make_kernel(spec), mirror=True)
    st = fresh_state(p, y, C, tol)
    st.trace = [] if trace else None
    n_iter, conv = _run(st, cols, max_iter)
    return Fit("svr", st.alpha, st.g, y, n_iter, conv, bias(st), coef=st.alpha[:l] - st.alpha[l:],
               X=X, spec=spec, trace=st.trace, lru_hits=cols.hits, lru_misses=cols.misses)


def fit_nusvr(X, z, C=1., nu=0.5, kernel="rbf", degree=3, gamma="scale", coef0=0., max_iter=1000,
              rff=False, D=1000, tol=1e-5, trace=False, spec: KernelSpec = None) -> Fit:
    """NuSVR.fit (svr.py:259-301): p = [-z, z], t = C*l*nu, bound C."""
    X, z = np.array(X), np.array(z)
    l = X.shape[0]
    y = np.empty(2 * l)
    y[:l], y[l:] = 1, -1
    p = np.empty(2 * l)
    p[:l], p[l:] = -z, z
    spec = spec or kernel_spec_for(X, kernel, gamma, coef0, degree, rff, D)
    cols = ColumnSource(X, y, make_kernel(spec), mirror=True)
    st = nu_state(p, y, C * l * nu, C, tol, cols)
    st.trace = [] if trace else None
    n_iter, conv = _run_nu(st, cols, max_iter)
    rho, b = bias_nu(st)
    return Fit("nusvr", st.alpha, st.g, y, n_iter, conv, rho, b, coef=st.alpha[:l] - st.alpha[l:],
               X=X, spec=spec, trace=st.trace, lru_hits=cols.hits, lru_misses=cols.misses)


def oneclass_alpha0(nu: float, l: int) -> np.ndarray:
    """oc_svm.py:76-83 -- including the leaked loop variable: alpha[n-1] gets the fractional part,
    alpha[n] stays 1, the rest is 0.  Raises like the reference when int(nu*l) == 0."""
    n = int(nu * l)
    if n == 0:
        raise UnboundLocalError("int(nu*l) == 0 (reference: oc_svm.py:80 uses the loop variable)")
    alpha = np.ones(l)
    if n < l:
        alpha[n - 1] = nu * l - n
    alpha[n + 1:] = 0
    return alpha


def fit_oneclass(X, nu=0.5, kernel="rbf", degree=3, gamma="scale", coef0=0., max_iter=1000,
                 rff=False, D=1000, tol=1e-5, trace=False, spec: KernelSpec = None) -> Fit:
    """OneClassSVM.fit, cached-solver path (oc_svm.py:57-112)."""
    X = np.array(X)
    l = X.shape[0]
    spec = spec or kernel_spec_for(X, kernel, gamma, coef0, degree, rff, D)
    y = np.ones(l)
    cols = OneClassColumns(X, y, make_kernel(spec), mirror=False)
    st = fresh_state(np.zeros(l), y, 1, tol)
    st.alpha = oneclass_alpha0(nu, l)
    for k in range(l):                       # oc_svm.py:94-95 (bypasses the LRU)
        st.g[k] -= y[k] * np.matmul(cols._evaluate(k), st.alpha)
    st.trace = [] if trace else None
    n_iter, conv = _run(st, cols, max_iter)
    return Fit("oneclass", st.alpha, st.g, y, n_iter, conv, bias(st), coef=st.alpha, X=X, spec=spec,
               trace=st.trace, lru_hits=cols.hits, lru_misses=cols.misses)


def fit_linear_svc(X, y, C=1., max_iter=1000, tol=1e-5, trace=False) -> Fit:
    """BiLinearSVC.fit, cached-solver path (svc.py:45-82): primal w accumulated per update."""
    X, y = np.array(X), _binary_labels(y)
    l, d = X.shape
    cols = LinearColumns(X, y, None, mirror=False)
    st = fresh_state(-np.ones(l), y, C, tol)
    st.trace = [] if trace else None
    w = np.zeros(d)
    n_iter, conv = max_iter, False
    for it in range(max_iter):
        i, j = select_pair(st)
        if i < 0:
            n_iter, conv = it, True
            break
        di, dj = take_step(st, i, j, cols(i), cols(j))
        w += di * y[i] * X[i] + dj * y[j] * X[j]
    return Fit("csvc", st.alpha, st.g, y, n_iter, conv, bias(st), coef=st.alpha * y, X=X,
               spec=KernelSpec("linear"), trace=st.trace, w=w)


def fit_linear_svr(X, z, C=1., eps=0., max_iter=1000, tol=1e-5, trace=False) -> Fit:
    """LinearSVR.fit, cached-solver path (svr.py:43-96)."""
    X, z = np.array(X), np.array(z)
    l, d = X.shape
    y, p = _svr_vectors(z, l, eps)
    cols = LinearColumns(X, y, None, mirror=True)
    st = fresh_state(p, y, C, tol)
    st.trace = [] if trace else None
    w = np.zeros(d)
    n_iter, conv = max_iter, False
    for it in range(max_iter):
        i, j = select_pair(st)
        if i < 0:
            n_iter, conv = it, True
            break
        di, dj = take_step(st, i, j, cols(i), cols(j))
        w += di * y[i] * X[i % l] + dj * y[j] * X[j % l]     # svr.py:89-90
    return Fit("svr", st.alpha, st.g, y, n_iter, conv, bias(st), coef=st.alpha[:l] - st.alpha[l:],
               X=X, spec=KernelSpec("linear"), trace=st.trace, w=w)


def dual_objective(fit: Fit, p: np.ndarray) -> float:
    """0.5*a'Qa + p'a = 0.5 * a.(grad f + p) with grad f = -y*g."""
    grad = -fit.y * fit.g
    return 0.5 * float(fit.alpha @ (grad + p))


# --------------------------------------------------------------------------------------
# synthetic workloads of BASELINE.json (SURVEY section 8d generators)
# --------------------------------------------------------------------------------------


def synth_classification(n=200_000, d=128, seed=0, sep=1.0):
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.5).astype(int)
    mu = rng.standard_normal(d)
    mu /= np.linalg.norm(mu)
    X = (rng.standard_normal((n, d)).astype(np.float32)
         + (2 * y[:, None] - 1) * sep * mu[None, :].astype(np.float32)).astype(np.float32)
    return X, y


def synth_regression(n=100_000, d=64, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    w = rng.standard_normal(d).astype(np.float32) / np.sqrt(d)
    z = (np.sin(X @ w) + 0.1 * rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    return X, z


def synth_oneclass(n=1_000_000, d=32, seed=1):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
