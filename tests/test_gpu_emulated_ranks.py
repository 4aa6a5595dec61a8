"""The two-level (GPU / box) candidate exchange of the row-sharded solve, exercised on ONE GPU: the ranks run as the
CTA groups of one cooperative launch (b2svm_solve_ranks_on_one_gpu) and push records + alpha + norm + rows into each
other's level-2 boxes exactly like GPUs do over NVLink.  Results must equal the single-GPU solve bit for bit: every
rank reduces the same records with the same (value, lower index) order and takes the same step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import smo_oracle as O


def _mods():
    import torch
    from pysvm_b200 import solver as S, sharded as SH, _native as N
    return torch, S, SH, N


def _single(S, X, y, p, C, spec, max_iter, mirror=False, cache_bytes=0, variant=None, state=None):
    func = S.QColumns(X, y, spec, mirror=mirror, cache_bytes=cache_bytes)
    sol = S.Solver.__new__(S.SolverWithCache)
    if variant is not None:
        sol._variant = variant
    S.Solver.__init__(sol, None, p, y, C, 1e-5, func=func)
    if state is not None:
        sol.alpha, sol.neg_y_grad = state
    st = sol.solve(max_iter)
    return sol, st


@pytest.mark.parametrize("world", [2, 4, 8])
def test_csvc_emulated_ranks_equal_single_gpu(world):
    torch, S, SH, N = _mods()
    X, y01 = O.synth_classification(n=6000, d=128, seed=1)
    y = np.where(y01 == 1, 1.0, -1.0)
    p = -np.ones(len(y))
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    ref, rst = _single(S, X, y, p, 1.0, spec, 10 ** 6)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
    stats = ranks.solve(10 ** 6)
    assert all(st.n_iter == rst.n_iter and st.converged == 1 for st in stats), ([st.n_iter for st in stats], rst.n_iter)
    np.testing.assert_array_equal(ranks.gather("alpha"), ref.alpha.cpu().numpy())
    np.testing.assert_array_equal(ranks.gather("neg_y_grad"), ref.neg_y_grad.cpu().numpy())
    # and against the reference algorithm
    gold = O.fit_csvc(X, y01, max_iter=10 ** 6)
    assert abs(rst.n_iter - gold.n_iter) <= max(1, 0.01 * gold.n_iter)
    np.testing.assert_array_equal(np.flatnonzero(ranks.gather("alpha") > 0), np.flatnonzero(gold.alpha > 0))


def test_capped_run_reports_the_same_next_pair_on_every_rank():
    torch, S, SH, N = _mods()
    X, y01 = O.synth_classification(n=20000, d=128, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    p = -np.ones(len(y))
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    ref, rst = _single(S, X, y, p, 1.0, spec, 700)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, 4)
    stats = ranks.solve(700)
    for st in stats:
        assert (st.n_iter, st.last_i, st.last_j) == (700, rst.last_i, rst.last_j)
        assert (st.last_delta_i, st.last_delta_j) == (rst.last_delta_i, rst.last_delta_j)
    np.testing.assert_array_equal(ranks.gather("alpha"), ref.alpha.cpu().numpy())


def test_svr_mirror_emulated_ranks():
    """2l variables, variables t and t+l share row t: a rank keeps [first half | second half] of its local rows."""
    torch, S, SH, N = _mods()
    X, z = O.synth_regression(n=3000, d=64, seed=0)
    l = X.shape[0]
    y = np.r_[np.ones(l), -np.ones(l)]
    p = np.r_[-z.astype(np.float64), z.astype(np.float64)]
    spec = S.KernelSpec("rbf", 1 / (64 * X.std()))
    ref, rst = _single(S, X, y, p, 1.0, spec, 3000, mirror=True)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, 3, mirror=True)
    stats = ranks.solve(3000)
    assert all(st.n_iter == rst.n_iter for st in stats)
    np.testing.assert_array_equal(ranks.gather("alpha"), ref.alpha.cpu().numpy())


def test_nu_variant_four_candidate_types():
    torch, S, SH, N = _mods()
    X, y01 = O.synth_classification(n=4000, d=20, seed=2)
    y = np.where(y01 == 1, 1.0, -1.0)
    l = len(y)
    p = np.zeros(l)
    spec = S.KernelSpec("rbf", 1 / (20 * X.std()))
    nus = S.NuSolverWithCache(p, y, 0.5 * l, 1.0, S.QColumns(X, y, spec, cache_bytes=0))
    a0, g0 = nus.alpha.cpu().numpy().copy(), nus.neg_y_grad.cpu().numpy().copy()
    rst = nus.solve(2500)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, 2, variant=N.SOLVER_NU)
    ranks.set_state(a0, g0)
    stats = ranks.solve(2500)
    assert all(st.n_iter == rst.n_iter == 2500 for st in stats)
    np.testing.assert_array_equal(ranks.gather("alpha"), nus.alpha.cpu().numpy())


@pytest.mark.parametrize("n,d,dtype,world", [(569, 30, np.float64, 2), (900, 512, np.float32, 2), (333, 7, np.float32, 3),
                                             (2000, 256, np.float64, 4)])
def test_row_widths_and_dtypes(n, d, dtype, world, breast_cancer):
    """Rows of every compiled width travel through the box (2 half-chunk units per 16-byte chunk), float32 and float64."""
    torch, S, SH, N = _mods()
    if n == 569:
        X, y01 = breast_cancer
    else:
        rng = np.random.default_rng(n + d)
        X = rng.standard_normal((n, d)).astype(dtype)
        y01 = (rng.random(n) < 0.5).astype(int)
    y = np.where(y01 == 1, 1.0, -1.0)
    p = -np.ones(n)
    spec = S.KernelSpec("rbf", 1 / (d * X.std()))
    ref, rst = _single(S, X, y, p, 1.0, spec, 10 ** 5)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
    stats = ranks.solve(10 ** 5)
    assert all(st.n_iter == rst.n_iter for st in stats)
    np.testing.assert_array_equal(ranks.gather("alpha"), ref.alpha.cpu().numpy())


def test_column_cache_in_the_sharded_solve():
    """Every rank caches its rows of the columns it has seen; the slot table is replicated and stays in step."""
    torch, S, SH, N = _mods()
    X, y01 = O.synth_classification(n=5000, d=128, seed=2)
    y = np.where(y01 == 1, 1.0, -1.0)
    p = -np.ones(len(y))
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    ref, rst = _single(S, X, y, p, 1.0, spec, 10 ** 6)
    ranks = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, 2, cache_bytes=1 << 30)
    stats = ranks.solve(10 ** 6)
    assert all(st.n_iter == rst.n_iter for st in stats)
    assert stats[0].cache_hits > 0 and stats[0].cache_hits == stats[1].cache_hits
    np.testing.assert_array_equal(ranks.gather("alpha"), ref.alpha.cpu().numpy())


def test_shard_planning_rejects_empty_ranks_on_every_rank():
    _, S, SH, N = _mods()
    assert not SH.shardable(200, 8) and SH.shardable(200, 4)
    for r in range(8):
        with pytest.raises(ValueError):
            SH.plan(200, 8, r)
