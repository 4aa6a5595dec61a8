"""Row-sharded SMO over the GPUs of one NVLink / NVSwitch box.

One process per GPU (``torch.distributed``).  Each rank holds ONLY its contiguous slice of the training rows and
the matching slices of ``alpha`` / ``neg_y_grad`` / cached columns.  Per SMO iteration the persistent kernels first
agree on a winner per GPU (level 1, local memory), then the CTA that owns a GPU's winner writes it -- record, alpha,
norm and the row x itself, as self-validating 16-byte units -- straight into every GPU's memory over NVLink (CUDA IPC
peer mappings); every GPU picks the same box-wide pair and already holds x_i, x_j for the next sweep.
``torch.distributed`` is plumbing only: the one-off exchange of IPC handles, a barrier before each solve, the
transient all-gather of rows for the nu / one-class initial gradient, and the final gather / reduction of results.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from .solver import KernelSpec, QColumns, SolverWithCache, kernel_matvec, nu_initial_alpha, to_device_matrix


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    n_global: int
    rows_per_rank: int      # roundup32(ceil(n / world)); the same formula as b2svm_create

    @property
    def row_offset(self) -> int:
        return self.rank * self.rows_per_rank

    @property
    def n_local(self) -> int:
        return max(0, min(self.rows_per_rank, self.n_global - self.row_offset))

    def rows(self) -> slice:
        return slice(self.row_offset, self.row_offset + self.n_local)


def rows_per_rank(n_global: int, world: int) -> int:
    return (-(-n_global // world) + 31) // 32 * 32


def shardable(n_global: int, world: int) -> bool:
    """True when EVERY rank of `world` owns at least one row.  A pure function of (n, world): all ranks take the
    same decision without talking (a rank that raised on its own would leave its peers waiting in a collective)."""
    return world >= 1 and (world - 1) * rows_per_rank(n_global, world) < n_global


def plan(n_global: int, world: int, rank: int) -> Shard:
    """Equal slices of whole 32-row tiles: rank r owns rows [r*rpr, min(n, (r+1)*rpr)).  Raises on every rank alike
    when the problem is too small for this many GPUs (estimators check `active_for` first and stay on one GPU)."""
    if not shardable(n_global, world):
        raise ValueError(f"{n_global} rows cannot be dealt to {world} GPUs in whole 32-row tiles without an empty rank")
    return Shard(rank, world, n_global, rows_per_rank(n_global, world))


def local_vector(v: np.ndarray, shard: Shard, mirror: bool) -> np.ndarray:
    """The slice of a per-variable vector this rank owns; mirror problems keep [first | second] halves."""
    r = shard.rows()
    if not mirror:
        return np.ascontiguousarray(v[r])
    n = shard.n_global
    return np.concatenate([v[r], v[n + r.start:n + r.stop]])


def combine_bias_partials(parts: np.ndarray) -> np.ndarray:
    """Merge per-rank b2svm_bias_partials rows [world][18]: sums / counts add, extremes min / max."""
    out = parts.sum(axis=0)
    out[2] = parts[:, 2].min()
    out[4] = parts[:, 4].max()
    for c in range(2):
        o = 6 + 6 * c
        out[o + 2] = parts[:, o + 2].max()
        out[o + 4] = parts[:, o + 4].min()
    return out


def rho_from_partials(p: np.ndarray) -> float:
    """Solver.calculate_rho (solver.py:160-177) from merged partials."""
    if p[1] > 0:
        return -(p[0] / p[1])
    if p[3] == 0 or p[5] == 0:
        raise ValueError("calculate_rho: empty bound set")
    return -(p[2] + p[4]) / 2


def rho_b_from_partials(p: np.ndarray):
    """NuSolver.calculate_rho_b (solver.py:378-419) from merged partials."""
    r = []
    for c in range(2):
        o = 6 + 6 * c
        if p[o + 1] > 0:
            r.append(p[o] / p[o + 1])
        elif p[o + 3] > 0 and p[o + 5] > 0:
            r.append((p[o + 2] + p[o + 4]) / 2)
        else:
            r.append(0.0)
    return (r[0] + r[1]) / 2, (r[1] - r[0]) / 2


def _group_device(group) -> torch.device:
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather_bytes(payload: bytes, group=None) -> bytes:
    """Concatenation over ranks of a fixed-size byte string (IPC handles)."""
    dev = _group_device(group)
    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = torch.empty(world * len(payload), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return bytes(out.cpu().numpy().tobytes())


def all_gather_array(a: np.ndarray, group=None) -> np.ndarray:
    dev = _group_device(group)
    world = dist.get_world_size(group)
    a = np.ascontiguousarray(a, dtype=np.float64)
    t = torch.from_numpy(a.reshape(-1)).to(dev)
    out = torch.empty(world * t.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape((world,) + a.shape)


class ShardedColumns(QColumns):
    """QColumns of one rank: THIS RANK'S rows of X (uploaded from the host slice; the other ranks' rows never reach
    this GPU's memory) with this rank's slice of y; carries the shard geometry to b2svm_create."""

    def __init__(self, X_full, y_local, kernel: KernelSpec, shard: Shard, mirror: bool = False, device=None,
                 cache_bytes=None):
        self.shard = shard
        self.X_host = X_full if not isinstance(X_full, torch.Tensor) else None     # kept by reference (no copy)
        X_local = X_full[shard.rows()]
        super().__init__(X_local, y_local, kernel, mirror=mirror, device=device, cache_bytes=cache_bytes)
        assert self.X.shape[0] == shard.n_local
        self.h2d_bytes = self.X.numel() * self.X.element_size()

    def local_rows(self) -> torch.Tensor:
        return self.X


class ShardedSolver(SolverWithCache):
    """``SolverWithCache`` whose ``solve`` runs row-sharded over the process group."""

    def __init__(self, p_local, y_local, C, tol, func: ShardedColumns, group=None, max_iter_hint=None):
        self.group = group
        super().__init__(p_local, y_local, C, tol, func=func, max_iter_hint=max_iter_hint)
        handles = all_gather_bytes(self._engine.export_mailbox(), group)
        self._engine.attach_mailboxes(handles)

    def solve(self, max_iter: int, func=None):
        eng = self._need(func)

        def sync_ranks():
            eng.prepare()                          # clear mailbox, publish alpha slice (synchronises the stream)
            dist.barrier(group=self.group)         # no peer may start before every mailbox is clean

        st = eng.solve(max_iter, before_launch=sync_ranks)
        self.last_stats = st
        return st

    def calculate_rho(self) -> float:
        parts = all_gather_array(self._engine.bias_partials(), self.group)
        return rho_from_partials(combine_bias_partials(parts))


def all_rows(func: "ShardedColumns", group=None) -> torch.Tensor:
    """Every rank's rows, assembled on this GPU for ONE computation (the caller drops the tensor afterwards): the
    initial gradient of the nu / one-class starts needs K against all support rows.  NCCL: one all-gather of the
    padded slices over NVLink; gloo (CPU tests): through the host."""
    shard, X = func.shard, func.X
    rpr, world = shard.rows_per_rank, shard.world
    mine = torch.zeros((rpr, X.shape[1]), dtype=X.dtype, device=X.device)
    mine[:shard.n_local] = X
    if dist.get_backend(group) == "nccl":
        out = torch.empty((world * rpr, X.shape[1]), dtype=X.dtype, device=X.device)
        dist.all_gather_into_tensor(out, mine, group=group)
    else:
        parts = [torch.empty_like(mine, device="cpu") for _ in range(world)]
        dist.all_gather(parts, mine.cpu(), group=group)
        out = torch.cat(parts).to(X.device)
    return out[:shard.n_global]


def rows_kernel_sum(func: "ShardedColumns", coef_rows: np.ndarray, group=None) -> torch.Tensor:
    """m_t = sum over ALL training rows s of coef_s K(x_s, x_t) for this rank's rows t: the initial gradient of the
    nu / one-class variants (solver.py:279-281, oc_svm.py:94-95).  The support rows of the other ranks are gathered
    for this one mat-vec (K7, b2svm_kernel_matvec) and released again; the solve itself only ever holds local rows."""
    coef = torch.from_numpy(np.ascontiguousarray(coef_rows, dtype=np.float64)).to(func.device)
    Xall = all_rows(func, group)
    out = kernel_matvec(func.kernel, Xall, coef, func.local_rows())
    del Xall
    return out


class ShardedNuSolver(ShardedSolver):
    """``NuSolverWithCache`` (solver.py:422-486) row-sharded: four candidate types cross NVLink per iteration."""
    _variant = N.SOLVER_NU

    def __init__(self, p_global, y_global, t: float, C, tol, func: ShardedColumns, mirror: bool, group=None,
                 max_iter_hint=None):
        self.t = t
        self._mirror = mirror
        self._y_global = np.asarray(y_global, dtype=np.float64)
        super().__init__(local_vector(np.asarray(p_global, dtype=np.float64), func.shard, mirror), func.y, C, tol, func, group,
                         max_iter_hint=max_iter_hint)

    def _init_state(self):
        # solver.py:269-281: the sequential alpha fill runs over the GLOBAL variable order; this rank keeps its slice.
        # neg_y_grad = -y (Q alpha0 + p) with (Q alpha0)_t = y_t sum_s y_s alpha0_s K(x_s mod l, x_t mod l)
        func, shard = self._engine.func, self._engine.func.shard
        a0 = nu_initial_alpha(self._y_global, self.t)
        ya = a0 * self._y_global
        l = shard.n_global
        coef_rows = ya[:l] + ya[l:] if self._mirror else ya
        m = rows_kernel_sum(func, coef_rows, self.group)
        if self._mirror:
            m = torch.cat([m, m])
        self._alpha.copy_(torch.from_numpy(local_vector(a0, shard, self._mirror)))
        self._g.copy_(-(m + self.y * self.p))
        self._state_ready = True

    def calculate_rho_b(self):
        parts = all_gather_array(self._engine.bias_partials(), self.group)
        return rho_b_from_partials(combine_bias_partials(parts))


class EmulatedRanks:
    """The ranks of a row-sharded solve emulated on ONE GPU (``b2svm_solve_ranks_on_one_gpu``): every rank is a
    group of CTAs of one cooperative launch and the level-2 exchange runs through local memory.  Same device code and
    the same host-side slicing as the multi-GPU path; used to test the exchange protocol on a single-GPU box.

        ranks = EmulatedRanks(X, y, p, C, tol, kernel, world=4)           # y, p: global per-variable vectors
        stats = ranks.solve(1000); alpha = ranks.gather(ranks.alphas())
    """

    def __init__(self, X, y, p, C, tol, kernel: KernelSpec, world: int, mirror: bool = False, variant=N.SOLVER_C,
                 cache_bytes=0, device=None):
        from .solver import Solver, default_device
        dev = torch.device(device) if device is not None else default_device()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        y, p = np.asarray(y, dtype=np.float64), np.asarray(p, dtype=np.float64)
        self.world, self.mirror = world, mirror
        self.shards, self.solvers = [], []
        for r in range(world):
            shard = plan(X.shape[0], world, r)
            func = ShardedColumns(X, local_vector(y, shard, mirror), kernel, shard, mirror=mirror, device=dev,
                                  cache_bytes=cache_bytes)
            func.max_ctas = max(1, sms // world)
            sol = Solver.__new__(SolverWithCache)
            sol._variant = variant
            Solver.__init__(sol, None, local_vector(p, shard, mirror), func.y, C, tol, func=func)
            self.shards.append(shard)
            self.solvers.append(sol)

    def set_state(self, alpha, neg_y_grad):
        """Global alpha / neg_y_grad (e.g. the nu / one-class start), dealt to the ranks."""
        for shard, sol in zip(self.shards, self.solvers):
            sol.alpha = local_vector(np.asarray(alpha, dtype=np.float64), shard, self.mirror)
            sol.neg_y_grad = local_vector(np.asarray(neg_y_grad, dtype=np.float64), shard, self.mirror)

    def solve(self, max_iter: int):
        import ctypes as C
        lib = N.load()
        handles = (C.c_void_p * self.world)(*[s._engine.h for s in self.solvers])
        stats = (N.SolveStats * self.world)()
        N.check(lib.b2svm_solve_ranks_on_one_gpu(handles, self.world, int(max_iter), stats))
        return [N.SolveStats.from_buffer_copy(bytes(st)) for st in stats]

    def gather(self, name: str) -> np.ndarray:
        """'alpha' or 'neg_y_grad' of all ranks in global variable order."""
        n = self.shards[0].n_global
        out = np.zeros(2 * n if self.mirror else n)
        for shard, sol in zip(self.shards, self.solvers):
            v = getattr(sol, name).cpu().numpy()
            r = shard.rows()
            out[r] = v[:shard.n_local]
            if self.mirror:
                out[n + r.start:n + r.stop] = v[shard.n_local:]
        return out


def schedule_subproblems(sizes, world: int):
    """Owner rank of each independent sub-problem (one-vs-one / one-vs-rest fits): longest processing time first --
    largest sub-problem to the least loaded GPU, ties to the lower rank / lower index -- a pure function of the sizes,
    so every rank computes the same assignment without talking."""
    load = [0] * world
    owner = [0] * len(sizes)
    for k in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[k] = r
        load[r] += sizes[k]
    return owner


def exchange_subproblem_results(results: dict, sizes, owner, nstat: int, group=None) -> list:
    """After the independent solves: ``results[k] = (stats_words, alpha, gradient)`` for the sub-problems this rank
    solved (stats_words = the solve statistics as ``nstat`` float64 words).  One all-gather of a flat float64 payload
    per rank -- laid out in sub-problem order, padded to the longest rank -- hands every rank every result.  Returns
    the list of (stats_words, alpha, gradient) for ALL sub-problems, in sub-problem order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lengths = [sum(nstat + 2 * sizes[k] for k in range(len(sizes)) if owner[k] == r) for r in range(world)]
    payload = np.zeros(max(lengths) if lengths else 0)
    o = 0
    for k in range(len(sizes)):
        if owner[k] != rank:
            continue
        st, a, g = results[k]
        payload[o:o + nstat] = st
        payload[o + nstat:o + nstat + sizes[k]] = a
        payload[o + nstat + sizes[k]:o + nstat + 2 * sizes[k]] = g
        o += nstat + 2 * sizes[k]
    everything = all_gather_array(payload, group)                     # [world][max length]
    out, offs = [], [0] * world
    for k in range(len(sizes)):
        r, o = owner[k], offs[owner[k]]
        row = everything[r]
        out.append((row[o:o + nstat].copy(), row[o + nstat:o + nstat + sizes[k]].copy(),
                    row[o + nstat + sizes[k]:o + nstat + 2 * sizes[k]].copy()))
        offs[r] = o + nstat + 2 * sizes[k]
    return out


_ENABLED = {"on": False, "group": None}


def enable(group=None):
    """Route ``BiKernelSVC.fit`` / ``KernelSVR.fit`` through the row-sharded solve of the current
    ``torch.distributed`` process group (every rank must call ``fit`` with the same host arrays)."""
    _ENABLED["on"], _ENABLED["group"] = True, group


def disable():
    _ENABLED["on"], _ENABLED["group"] = False, None


def active() -> bool:
    return _ENABLED["on"] and dist.is_available() and dist.is_initialized() and dist.get_world_size(_ENABLED["group"]) > 1


def active_for(n_rows: int) -> bool:
    """Route a fit of `n_rows` training rows through the sharded solve?  The same answer on every rank."""
    return active() and shardable(n_rows, dist.get_world_size(_ENABLED["group"]))


def gather_variables(v_local: np.ndarray, shard: Shard, mirror: bool, group=None) -> np.ndarray:
    """All ranks' slices of a per-variable vector, reassembled in global variable order."""
    rpr, n = shard.rows_per_rank, shard.n_global
    halves = 2 if mirror else 1
    pad = np.zeros((halves, rpr))
    for h in range(halves):
        pad[h, :shard.n_local] = v_local[h * shard.n_local:(h + 1) * shard.n_local]
    allp = all_gather_array(pad, group)                      # [world][halves][rpr]
    return np.concatenate([allp[:, h, :].reshape(-1)[:n] for h in range(halves)])


class ShardedFit:
    """What a sharded estimator keeps after ``fit`` (the counterpart of svc._Fitted)."""

    def __init__(self, solver: "ShardedSolver", shard: Shard, coef_local: torch.Tensor, mirror: bool, group=None):
        self.solver, self.shard, self.coef_local, self.mirror, self.group = solver, shard, coef_local, mirror, group
        self.func = solver._engine.func
        solver.release_cache()        # (as svc._Fitted: the fit is over, the column cache goes back to the device)

    def kernel_sum(self, x) -> np.ndarray:
        return decision_sharded(self.solver, self.coef_local, x, self.group)


def fit_csvc_sharded(X: np.ndarray, y01: np.ndarray, C: float, kernel: KernelSpec, tol: float, max_iter: int,
                     group=None, cache_bytes=None):
    """Row-sharded BiKernelSVC solve.  Every rank passes the full host arrays and keeps its slice.
    Returns (solver, shard, stats); ``solver.alpha`` is this rank's slice."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    y = np.array(y01, dtype=float)
    y[y != 1] = -1
    shard = plan(X.shape[0], world, rank)
    func = ShardedColumns(X, local_vector(y, shard, False), kernel, shard, cache_bytes=cache_bytes)
    solver = ShardedSolver(local_vector(-np.ones(X.shape[0]), shard, False), func.y, C, tol, func, group, max_iter_hint=max_iter)
    stats = solver.solve(max_iter)
    return solver, shard, stats


def fit_svr_sharded(X: np.ndarray, z: np.ndarray, C: float, eps: float, kernel: KernelSpec, tol: float,
                    max_iter: int, group=None):
    """Row-sharded KernelSVR solve (2l variables, variables t and t+l share row t)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    l = X.shape[0]
    y = np.empty(2 * l)
    y[:l], y[l:] = 1., -1.
    p = np.ones(2 * l) * eps
    p[:l] -= z
    p[l:] += z
    shard = plan(l, world, rank)
    func = ShardedColumns(X, local_vector(y, shard, True), kernel, shard, mirror=True)
    solver = ShardedSolver(local_vector(p, shard, True), func.y, C, tol, func, group, max_iter_hint=max_iter)
    stats = solver.solve(max_iter)
    return solver, shard, stats


def fit_oneclass_sharded(X: np.ndarray, nu: float, kernel: KernelSpec, tol: float, max_iter: int, group=None):
    """Row-sharded OneClassSVM solve (oc_svm.py:59-105): y = +1, p = 0, C = 1, alpha0 per oc_svm.py:76-83 over the
    global row order, neg_y_grad = -K alpha0."""
    from .svm.oc_svm import oneclass_initial_alpha
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    l = X.shape[0]
    shard = plan(l, world, rank)
    func = ShardedColumns(X, np.ones(shard.n_local), kernel, shard)
    solver = ShardedSolver(np.zeros(shard.n_local), func.y, 1.0, tol, func, group, max_iter_hint=max_iter)
    a0 = oneclass_initial_alpha(nu, l)
    solver.alpha = local_vector(a0, shard, False)
    solver.neg_y_grad = -rows_kernel_sum(func, a0, group)
    stats = solver.solve(max_iter)
    return solver, shard, stats


def fit_nusvc_sharded(X: np.ndarray, y01: np.ndarray, nu: float, kernel: KernelSpec, tol: float, max_iter: int, group=None):
    """Row-sharded BiNuSVC solve (svc.py:405-433): p = 0, C = 1, e'alpha = nu * l."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    y = np.array(y01, dtype=float)
    y[y != 1] = -1
    l = X.shape[0]
    shard = plan(l, world, rank)
    func = ShardedColumns(X, local_vector(y, shard, False), kernel, shard)
    solver = ShardedNuSolver(np.zeros(l), y, nu * l, 1.0, tol, func, False, group, max_iter_hint=max_iter)
    stats = solver.solve(max_iter)
    return solver, shard, stats


def fit_nusvr_sharded(X: np.ndarray, z: np.ndarray, C: float, nu: float, kernel: KernelSpec, tol: float, max_iter: int,
                      group=None):
    """Row-sharded NuSVR solve (svr.py:256-294): 2l variables, p = [-z, z], e'alpha = C * l * nu."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    l = X.shape[0]
    y = np.empty(2 * l)
    y[:l], y[l:] = 1., -1.
    p = np.empty(2 * l)
    p[:l], p[l:] = -np.asarray(z, dtype=float), np.asarray(z, dtype=float)
    shard = plan(l, world, rank)
    func = ShardedColumns(X, local_vector(y, shard, True), kernel, shard, mirror=True)
    solver = ShardedNuSolver(p, y, C * l * nu, C, tol, func, True, group, max_iter_hint=max_iter)
    stats = solver.solve(max_iter)
    return solver, shard, stats


def decision_sharded(solver: ShardedSolver, coef_local: torch.Tensor, Xq: np.ndarray, group=None) -> np.ndarray:
    """sum over ALL support rows of coef K(x_s, x_q): each rank sums its slice, one all-reduce adds them."""
    func = solver._engine.func
    q = to_device_matrix(np.asarray(Xq), func.device).to(func.X.dtype)
    part = kernel_matvec(func.kernel, func.local_rows(), coef_local, q)
    if dist.get_backend(group) != "nccl":
        part = part.cpu()
    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part.cpu().numpy()
