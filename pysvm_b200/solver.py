"""Host-side mirror of ``pysvm/solver.py``: the four solver classes with the reference's
constructor / method names, backed by the C-ABI library (``include/b2svm.h``).

Differences a PySVM user sees (all additive):

* ``alpha`` / ``neg_y_grad`` are float64 CUDA tensors (the reference holds NumPy arrays); assigning
  to them copies into the device buffers (OneClassSVM does that, reference ``oc_svm.py:88-95``).
* ``func`` -- the Q-column callback -- must be a :class:`QColumns` object (it carries X, y and the
  kernel to the device); an arbitrary Python callable cannot run on the GPU and is rejected.  It can
  be given at construction (``func=``) or with the first ``update`` / ``solve`` call.
* ``solve(max_iter)`` runs the estimator's whole ``for n_iter in range(max_iter)`` loop in one
  persistent kernel (``b2svm_solve``); ``working_set_select`` / ``update`` remain for step-wise use
  and go through the same device code.
* The LRU of 256 columns (``solver.py:207,227-229``) is numerically transparent and is not
  reproduced; ``cache_size`` is accepted for signature compatibility.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _native as N


@dataclass
class KernelSpec:
    """Resolved kernel (reference ``register_kernel``, svc.py:210-241)."""
    kind: str = "rbf"
    gamma: float = 1.0
    coef0: float = 0.0
    degree: float = 3.0

    def desc(self) -> N.KernelDesc:
        return N.KernelDesc(N.KERNEL_IDS[self.kind], 0, float(self.gamma), float(self.coef0), float(self.degree))


def host_std(X) -> float:
    """``X.std()`` of the reference's register_kernel (svc.py:212-217), bit-identical to numpy but evaluated
    on several host threads for large C-contiguous float matrices (``b2svm_host_std``).  Returns a NumPy
    scalar of X's dtype, so ``1 / (n_features * std)`` rounds exactly as it does in the reference."""
    X = np.asarray(X)
    if X.dtype in (np.float32, np.float64) and X.flags.c_contiguous and X.size >= (1 << 20):
        out = C.c_double()
        N.check(N.load().b2svm_host_std(X.ctypes.data, X.size, N.F64 if X.dtype == np.float64 else N.F32, C.byref(out)))
        return X.dtype.type(out.value)
    return X.std()


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pysvm_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _dev_f64(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)


def to_device_matrix(X, device) -> torch.Tensor:
    """X as a contiguous float32/float64 CUDA matrix; the dtype of a float input is kept because the
    reference evaluates kernels in X's dtype."""
    if isinstance(X, torch.Tensor):
        t = X
    else:
        X = np.asarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(X))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(device).contiguous()


class QColumns:
    """The reference's ``func(i)`` closure (svc.py:257-258, svr.py:174-179, oc_svm.py:72-73) as an
    object: Q[:, i] = y * K(X, x_{i mod l}) * y_i evaluated on the device."""

    def __init__(self, X, y, kernel: KernelSpec, mirror: bool = False, device=None, cache_bytes=None):
        """``cache_bytes``: HBM budget of the device-resident Q-column cache (the reference's LRU,
        solver.py:207,227-229, at B200 scale).  ``None`` = automatic (up to 70 % of the free memory, never
        more than the full kernel matrix; ``B2SVM_CACHE_GB`` overrides), ``0`` = off (every iteration
        sweeps X: the cold path)."""
        self.cache_bytes = cache_bytes
        self.device = torch.device(device) if device is not None else default_device()
        self.X = to_device_matrix(X, self.device)
        self.y = _dev_f64(y, self.device)
        self.kernel = kernel
        self.mirror = bool(mirror)
        l = self._n_local_rows()
        if self.y.shape[0] != (2 * l if mirror else l):
            raise ValueError("y must have n_rows entries (2*n_rows in mirror mode)")
        self._engine = None

    def _n_local_rows(self) -> int:
        return self.X.shape[0]

    def __call__(self, i: int) -> torch.Tensor:
        eng = self._engine() if self._engine is not None else None
        if eng is None:
            raise RuntimeError("QColumns is not bound to a (live) solver")
        return eng.column(int(i))


class _Engine:
    """Owns one b2svm handle; keeps every tensor the handle points at alive."""

    def __init__(self, func: QColumns, alpha: torch.Tensor, g: torch.Tensor, Cbox: float, tol: float, variant: int,
                 max_ctas: int = 0, max_columns: Optional[int] = None):
        shard = getattr(func, "shard", None)
        self.lib = N.load()
        self.func = func
        self.alpha, self.g = alpha, g
        X = func.X
        d = N.ProblemDesc()
        d.struct_bytes = C.sizeof(N.ProblemDesc)
        d.device = func.device.index if func.device.index is not None else torch.cuda.current_device()
        d.stream = None
        d.n_rows, d.n_features = (shard.n_local if shard is not None else X.shape[0]), X.shape[1]
        d.x_dtype = N.F64 if X.dtype == torch.float64 else N.F32
        d.X = X.data_ptr()
        d.ldx = X.stride(0)
        d.n_vars = func.y.shape[0]
        d.variant = variant
        d.y = func.y.data_ptr()
        d.alpha = alpha.data_ptr()
        d.neg_y_grad = g.data_ptr()
        d.C, d.tol = float(Cbox), float(tol)
        d.kernel = func.kernel.desc()
        esz = 8 if X.dtype == torch.float64 else 4
        n_glob = shard.n_global if shard is not None else X.shape[0]
        full_q = int(n_glob) * int(d.n_rows) * esz
        cb = func.cache_bytes
        if max_columns is not None and cb is None and max_columns < 2 * int(n_glob) * (2 if func.mirror else 1):
            # automatic policy: columns are only re-used once variables get selected again, i.e. after the
            # opening phase in which every step drives fresh variables to a bound.  A run capped below one
            # step per variable cannot get there, so it skips the cache (and its 4 B/row store traffic).
            cb = 0
        if cb is None:
            env = os.environ.get("B2SVM_CACHE_GB")
            if env is not None:
                cb = int(float(env) * (1 << 30))
            else:
                free, _total = torch.cuda.mem_get_info(func.device)      # (a ~3 ms driver call: only when needed)
                cb = int(free * 0.7)
            if max_columns is not None:
                cb = min(int(cb), int(max_columns) * int(d.n_rows) * esz)
        d.cache_bytes = max(0, min(int(cb), full_q))
        d.max_ctas = max_ctas or int(getattr(func, "max_ctas", 0) or 0)
        if shard is None:
            d.rank, d.world = 0, 1
            d.row_offset, d.n_rows_global = 0, X.shape[0]
        else:
            d.rank, d.world = shard.rank, shard.world
            d.row_offset, d.n_rows_global = shard.row_offset, shard.n_global
        self.desc = d
        h = C.c_void_p()
        torch.cuda.synchronize(func.device)          # handle work runs on the legacy default stream
        N.check(self.lib.b2svm_create(C.byref(d), C.byref(h)))
        self.h = h
        func._engine = weakref.ref(self)        # (weak: solver -> engine -> func must not be a cycle, or the handle and its
                                                #  cache -- tens of GB -- would outlive `del solver` until a gc pass)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2svm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def release_cache(self):
        """Hand the Q-column cache back to the device (b2svm_release_cache); the handle stays usable."""
        if getattr(self, "h", None):
            N.check(self.lib.b2svm_release_cache(self.h))

    # -- thin wrappers -----------------------------------------------------------------------
    def init_state(self, p: torch.Tensor):
        N.check(self.lib.b2svm_init_state(self.h, p.data_ptr()))

    def init_gradient(self, p: Optional[torch.Tensor]):
        N.check(self.lib.b2svm_init_gradient(self.h, p.data_ptr() if p is not None else None))

    def select(self):
        i, j = C.c_int64(), C.c_int64()
        N.check(self.lib.b2svm_select(self.h, C.byref(i), C.byref(j)))
        return int(i.value), int(j.value)

    def select_wss3(self):
        i, j = C.c_int64(), C.c_int64()
        N.check(self.lib.b2svm_select_wss3(self.h, C.byref(i), C.byref(j)))
        return int(i.value), int(j.value)

    def update(self, i: int, j: int):
        di, dj = C.c_double(), C.c_double()
        N.check(self.lib.b2svm_update(self.h, int(i), int(j), C.byref(di), C.byref(dj)))
        return di.value, dj.value

    def column(self, i: int) -> torch.Tensor:
        out = torch.empty(self.func.y.shape[0], dtype=torch.float64, device=self.func.device)
        N.check(self.lib.b2svm_kernel_column(self.h, int(i), out.data_ptr()))
        return out

    def solve(self, max_iter: int, before_launch=None) -> N.SolveStats:
        """b2svm_solve; a loop longer than B2SVM_MAX_ITER - 1 iterations is issued in chunks (the state lives in
        alpha / neg_y_grad, so a chunk boundary is invisible to the result).  ``before_launch`` runs ahead of every
        launch (the sharded solver prepares the mailboxes and synchronises the ranks there)."""
        left, total = int(max_iter), None
        while True:
            chunk = min(left, N.MAX_ITER - 1)
            if before_launch is not None:
                before_launch()
            st = N.SolveStats()
            N.check(self.lib.b2svm_solve(self.h, chunk, C.byref(st)))
            done_now = st.n_iter
            if total is None:
                total = st
            else:
                for name in ("n_iter", "cold_sweeps", "cache_hits", "cache_misses", "algorithmic_bytes", "device_ms", "launches"):
                    setattr(st, name, getattr(total, name) + getattr(st, name))
                total = st
            left -= chunk
            if left <= 0 or total.converged or done_now < chunk:
                return total

    def rho(self) -> float:
        r = C.c_double()
        N.check(self.lib.b2svm_rho(self.h, C.byref(r)))
        return r.value

    def rho_b(self):
        r, b = C.c_double(), C.c_double()
        N.check(self.lib.b2svm_rho_b(self.h, C.byref(r), C.byref(b)))
        return r.value, b.value

    def bias_partials(self) -> np.ndarray:
        out = (C.c_double * 18)()
        N.check(self.lib.b2svm_bias_partials(self.h, out))
        return np.array(out[:], dtype=np.float64)

    def prepare(self):
        N.check(self.lib.b2svm_prepare(self.h))

    def export_mailbox(self) -> bytes:
        buf = C.create_string_buffer(64)
        N.check(self.lib.b2svm_mailbox_export(self.h, buf))
        return buf.raw

    def attach_mailboxes(self, handles: bytes):
        N.check(self.lib.b2svm_mailbox_attach(self.h, handles))


class Solver:
    r"""SMO on  min ½ αᵀQα + pᵀα  s.t. yᵀα = 0, 0 ≤ α ≤ C  (reference ``Solver``, solver.py:5-184).

    Columns come from ``func`` (a :class:`QColumns`), or -- the reference's dense-Q mode -- from a full
    ``Q`` matrix, which the device solver uses as a precomputed kernel matrix.
    """
    _variant = N.SOLVER_C

    def __init__(self, Q, p, y, C: float, tol: float = 1e-5, func: Optional[QColumns] = None, device=None,
                 max_iter_hint: Optional[int] = None):
        self._max_iter_hint = max_iter_hint
        if Q is not None:
            # dense-Q mode of the reference (solver.py:25-45 with cache_size == 0): Q = y y^T * K is handed
            # over whole.  The device solver takes K itself as a precomputed kernel matrix.
            if func is not None:
                raise ValueError("pass either Q or func, not both")
            dev = torch.device(device) if device is not None else default_device()
            Qd = to_device_matrix(Q, dev)
            yd = _dev_f64(y, dev).to(Qd.dtype)
            if Qd.shape[0] != Qd.shape[1] or Qd.shape[0] != yd.shape[0]:
                raise ValueError("Q must be n_vars x n_vars (mirror problems: pass func=QColumns(K, y, 'precomputed', mirror=True))")
            K = (Qd * yd[:, None] * yd[None, :]).contiguous()
            func = QColumns(K, y, KernelSpec("precomputed"), device=dev, cache_bytes=0)
        self.device = (func.device if func is not None else
                       (torch.device(device) if device is not None else default_device()))
        self.Q = None
        self.p = _dev_f64(p, self.device)
        self.y = func.y if func is not None else _dev_f64(y, self.device)
        n = self.p.shape[0]
        assert n == self.y.shape[0]
        self.C = C
        self.tol = tol
        self._alpha = torch.zeros(n, dtype=torch.float64, device=self.device)
        self._g = torch.empty(n, dtype=torch.float64, device=self.device)
        self._engine: Optional[_Engine] = None
        self._state_ready = False
        self.last_stats: Optional[N.SolveStats] = None
        if func is not None:
            self.bind(func)

    # -- binding to device data ---------------------------------------------------------------
    def bind(self, func: QColumns):
        if self._engine is not None:
            if func is not None and func is not self._engine.func:
                raise ValueError("this solver is already bound to another QColumns object")
            return self._engine
        if not isinstance(func, QColumns):
            raise TypeError("func must be a pysvm_b200.solver.QColumns (an arbitrary Python callable cannot "
                            "run on the GPU; there is no CPU fallback)")
        self.y = func.y
        hint = getattr(self, "_max_iter_hint", None)
        self._engine = _Engine(func, self._alpha, self._g, self.C, self.tol, self._variant,
                               max_columns=None if hint is None else 2 * int(hint) + 2)
        self._init_state()
        return self._engine

    def _init_state(self):
        # Solver.__init__: alpha = 0, neg_y_grad = -y * p (solver.py:42-45)
        self._engine.init_state(self.p)
        self._state_ready = True

    def _need(self, func=None) -> _Engine:
        if self._engine is None:
            if func is None:
                raise RuntimeError("solver is not bound: pass func=QColumns(...) at construction or call bind()")
            self.bind(func)
        return self._engine

    # -- state ------------------------------------------------------------------------------------
    @property
    def alpha(self) -> torch.Tensor:
        return self._alpha

    @alpha.setter
    def alpha(self, value):
        self._alpha.copy_(_dev_f64(value, self.device))

    @property
    def neg_y_grad(self) -> torch.Tensor:
        self._need()
        return self._g

    @neg_y_grad.setter
    def neg_y_grad(self, value):
        self._g.copy_(_dev_f64(value, self.device))

    # -- the reference's methods --------------------------------------------------------------------
    def working_set_select(self):
        """solver.py:47-75."""
        return self._need().select()

    def working_set_select_wss3(self):
        """OPT-IN, not in the reference: LIBSVM's second-order rule (WSS 3).  Same stopping test as ``working_set_select``."""
        return self._need().select_wss3()

    def update(self, i: int, j: int, func=None):
        """solver.py:77-147."""
        return self._need(func).update(i, j)

    def calculate_rho(self) -> float:
        """solver.py:149-177."""
        return self._need().rho()

    def get_Q(self, i: int, func=None) -> torch.Tensor:
        """solver.py:179-184 / 227-229."""
        return self._need(func).column(i)

    def release_cache(self) -> None:
        """Free the device-resident Q-column cache once no further solve is planned (the estimators do it at the end
        of ``fit``: the reference keeps the solver alive inside its ``decision_function`` closure, svc.py:269-272, and
        here that would pin up to 70 % of the device memory).  rho, alpha and further (cache-less) solves still work."""
        if self._engine is not None:
            self._engine.release_cache()

    # -- fused driver loop ----------------------------------------------------------------------------
    def solve(self, max_iter: int, func=None, wss: str = "mvp") -> N.SolveStats:
        """The estimator loop ``for n_iter in range(max_iter): select; break if i < 0; update``
        (svc.py:260-267) on the device.  ``stats.n_iter`` = number of updates made.

        ``wss="mvp"`` (default) is the reference's maximal-violating-pair rule, fused into one persistent kernel.
        ``wss="wss3"`` opts into LIBSVM's second-order selection: a host-driven loop of ``b2svm_select_wss3`` +
        ``b2svm_update`` (several launches per iteration) that reaches the same optimum in fewer, differently counted
        iterations -- parity with the reference's iteration counts does not apply to it."""
        eng = self._need(func)
        if wss == "mvp":
            st = eng.solve(max_iter)
        elif wss == "wss3":
            st = N.SolveStats()
            n = 0
            nxt = (-1, -1)
            while n < int(max_iter):
                nxt = eng.select_wss3()
                if nxt[0] < 0:
                    st.converged = 1
                    break
                st.last_delta_i, st.last_delta_j = eng.update(*nxt)
                n += 1
            st.n_iter, st.n_ctas = n, 0
            st.last_i, st.last_j = nxt if n >= int(max_iter) else (-1, -1)
            st.cold_sweeps, st.launches = 2 * n, 8 * n
        else:
            raise ValueError("wss must be 'mvp' (the reference's rule) or 'wss3'")
        self.last_stats = st
        return st


class SolverWithCache(Solver):
    """Reference ``SolverWithCache(p, y, C, tol, cache_size)`` (solver.py:187-229)."""
    cache_size = 256

    def __init__(self, p, y, C: float, tol: float = 1e-5, cache_size: int = 256, func: Optional[QColumns] = None,
                 device=None, max_iter_hint: Optional[int] = None):
        super().__init__(None, p, y, C, tol, func=func, device=device, max_iter_hint=max_iter_hint)


def nu_initial_alpha(y: np.ndarray, t: float) -> np.ndarray:
    """Vectorised form of the sequential fill in solver.py:269-278: per class, in index order, each
    variable receives min(1., what is left of t/2).  The running remainder after k unit steps is
    t/2 - k exactly, so alpha_k = clip(t/2 - rank_k, 0, 1)."""
    y = np.asarray(y)
    alpha = np.zeros(y.shape[0])
    for cls in (y == 1, y != 1):
        rank = np.cumsum(cls) - 1
        alpha[cls] = np.clip(t / 2 - rank[cls], 0.0, 1.0)
    return alpha


class NuSolver(Solver):
    r"""ν-variant with the extra constraint eᵀα = t (reference ``NuSolver``, solver.py:232-419)."""
    _variant = N.SOLVER_NU

    def __init__(self, Q, p, y, t: float, C: float, tol: float = 1e-5, func: Optional[QColumns] = None, device=None,
                 max_iter_hint: Optional[int] = None):
        self.t = t
        self._y_host = np.asarray(y.detach().cpu().numpy() if isinstance(y, torch.Tensor) else y, dtype=np.float64)
        super().__init__(Q, p, y, C, tol, func=func, device=device, max_iter_hint=max_iter_hint)

    def _init_state(self):
        # solver.py:269-281 / 456-473: alpha fill, then neg_y_grad = -y * (Q alpha + p)
        self._alpha.copy_(torch.from_numpy(nu_initial_alpha(self._y_host, self.t)))
        self._engine.init_gradient(self.p)
        self._state_ready = True

    def working_set_select(self, func=None):
        """solver.py:283-339.  Returns (i, j, Qi, Qj) like the reference; the columns are produced on
        demand (``update`` does not need them -- it recomputes the three kernel values it uses)."""
        eng = self._need(func)
        i, j = eng.select()
        if i < 0:
            return -1, -1, -1, -1
        return i, j, _LazyColumn(eng, i), _LazyColumn(eng, j)

    def update(self, i: int, j: int, Qi=None, Qj=None):
        """solver.py:341-376."""
        return self._need().update(i, j)

    def calculate_rho_b(self):
        """solver.py:378-419."""
        return self._need().rho_b()


class NuSolverWithCache(NuSolver):
    """Reference ``NuSolverWithCache(p, y, t, C, func, tol, cache_size)`` (solver.py:422-486)."""
    cache_size = 256

    def __init__(self, p, y, t: float, C: float, func: QColumns, tol: float = 1e-5, cache_size: int = 256,
                 max_iter_hint: Optional[int] = None):
        super().__init__(None, p, y, t, C, tol, func=func, max_iter_hint=max_iter_hint)


class _LazyColumn:
    def __init__(self, engine: _Engine, i: int):
        self._engine, self._i, self._t = engine, i, None

    def tensor(self) -> torch.Tensor:
        if self._t is None:
            self._t = self._engine.column(self._i)
        return self._t

    def __getitem__(self, k):
        return self.tensor()[k]

    def __array__(self, dtype=None):
        a = self.tensor().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


def kernel_matvec(kernel: KernelSpec, Xa: torch.Tensor, coef: torch.Tensor, Xb: torch.Tensor) -> torch.Tensor:
    """out[q] = sum_s coef[s] K(Xa[s], Xb[q]) as float64 (b2svm_kernel_matvec)."""
    lib = N.load()
    assert Xa.dtype == Xb.dtype and Xa.shape[1] == Xb.shape[1]
    Xa, Xb = Xa.contiguous(), Xb.contiguous()
    coef = coef.to(dtype=torch.float64).contiguous()
    out = torch.empty(Xb.shape[0], dtype=torch.float64, device=Xa.device)
    kd = kernel.desc()
    torch.cuda.synchronize(Xa.device)
    N.check(lib.b2svm_kernel_matvec(Xa.device.index or 0, None, C.byref(kd),
                                    N.F64 if Xa.dtype == torch.float64 else N.F32, Xa.shape[1],
                                    Xa.data_ptr(), Xa.shape[0], Xa.stride(0), Xb.data_ptr(), Xb.shape[0], Xb.stride(0),
                                    coef.data_ptr(), out.data_ptr()))
    return out


def kernel_matrix(kernel: KernelSpec, Xa: torch.Tensor, Xb: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K[i][j] = K(Xa[i], Xb[j]) as a dense device matrix in X's dtype (``b2svm_kernel_matrix``): what the reference's
    dense-Q mode and its rff=True kernel build with ``kernel_func(X, X)`` on the host (svc.py:251-253, 225-228)."""
    lib = N.load()
    Xb = Xa if Xb is None else Xb
    assert Xa.dtype == Xb.dtype and Xa.shape[1] == Xb.shape[1]
    Xa, Xb = Xa.contiguous(), Xb.contiguous()
    nb = Xb.shape[0]
    ldo = (nb + 3) // 4 * 4 if Xa.dtype == torch.float32 else (nb + 1) // 2 * 2        # 16-byte row pitch
    out = torch.empty((Xa.shape[0], ldo), dtype=Xa.dtype, device=Xa.device)
    kd = kernel.desc()
    torch.cuda.synchronize(Xa.device)
    N.check(lib.b2svm_kernel_matrix(Xa.device.index or 0, None, C.byref(kd), N.F64 if Xa.dtype == torch.float64 else N.F32,
                                    Xa.shape[1], Xa.data_ptr(), Xa.shape[0], Xa.stride(0), Xb.data_ptr(), nb, Xb.stride(0),
                                    out.data_ptr(), ldo))
    return out if ldo == nb else out[:, :nb].contiguous()


def weighted_rows(X: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """sum_r w[r] X[r] as float64 (``b2svm_weighted_rows``): X^T (alpha*y) of the linear estimators, Z^T coef of RFF."""
    lib = N.load()
    X = X.contiguous()
    w = w.to(dtype=torch.float64).contiguous()
    out = torch.empty(X.shape[1], dtype=torch.float64, device=X.device)
    torch.cuda.synchronize(X.device)
    N.check(lib.b2svm_weighted_rows(X.device.index or 0, None, N.F64 if X.dtype == torch.float64 else N.F32, X.data_ptr(),
                                    X.shape[0], X.stride(0), X.shape[1], w.data_ptr(), out.data_ptr()))
    return out


def solve_batched(solvers, max_iter: int):
    """Solve several independent problems concurrently, one CTA group each, in ONE persistent launch
    (``b2svm_solve_batched``): the one-vs-one / one-vs-rest sub-problems of the multiclass wrappers
    (reference svc.py:332-342 + sklearn.multiclass).  Problems are grouped by (dtype, row width)."""
    lib = N.load()
    groups = {}
    for s in solvers:
        eng = s._need()
        key = (eng.desc.x_dtype, eng.desc.n_features, eng.desc.device)
        groups.setdefault(key, []).append(s)
    for group in groups.values():
        n = len(group)
        handles = (C.c_void_p * n)(*[s._engine.h for s in group])
        stats = (N.SolveStats * n)()
        N.check(lib.b2svm_solve_batched(handles, n, int(max_iter), stats))
        for s, st in zip(group, stats):
            cp = N.SolveStats()
            C.memmove(C.byref(cp), C.byref(st), C.sizeof(N.SolveStats))
            s.last_stats = cp
    return [s.last_stats for s in solvers]
