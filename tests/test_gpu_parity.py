"""GPU parity tests: the CUDA path (through the C-ABI library) against the oracle and the golden
vectors of the unmodified reference.  Tolerances (BASELINE.json north_star): identical support
vector sets and labels, alpha / rho / decision values within 1e-4 relative, iteration counts
within 1 %.  float64 inputs are evaluated in float64 on the device, so they agree far tighter."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import smo_oracle as O


def _mods():
    import torch
    import pysvm_b200
    from pysvm_b200 import solver as S
    from pysvm_b200.svm import svc, svr, oc_svm
    return torch, pysvm_b200, S, svc, svr, oc_svm


def rel_close(a, b, rtol=1e-4, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def iters_close(n, ref, frac=0.01):
    assert abs(n - ref) <= max(1, frac * ref), (n, ref)


# ------------------------------------------------------------------------------ kernel columns
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["rbf", "linear", "poly", "sigmoid"])
@pytest.mark.parametrize("d", [2, 30, 64, 100])
def test_kernel_column_matches_oracle(dtype, kind, d):
    torch, _, S, *_ = _mods()
    rng = np.random.default_rng(d)
    n = 333
    X = rng.standard_normal((n, d)).astype(dtype)
    y = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    spec = O.KernelSpec(kind=kind, gamma=0.7 / d, coef0=0.25, degree=3)
    cols = O.ColumnSource(X, y, O.make_kernel(spec), mirror=False, lru=0)
    func = S.QColumns(X, y, S.KernelSpec(kind, spec.gamma, spec.coef0, spec.degree))
    solver = S.SolverWithCache(-np.ones(n), y, 1.0, func=func)
    tol = 2e-5 if dtype == np.float32 else 1e-12
    for i in (0, 17, n - 1):
        got = solver.get_Q(i).cpu().numpy()
        want = cols(i).astype(np.float64)
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol * max(1.0, np.abs(want).max()))


def test_kernel_column_mirror():
    torch, _, S, *_ = _mods()
    rng = np.random.default_rng(1)
    l, d = 200, 16
    X = rng.standard_normal((l, d)).astype(np.float32)
    y = np.r_[np.ones(l), -np.ones(l)]
    spec = O.KernelSpec(kind="rbf", gamma=0.05)
    cols = O.ColumnSource(X, y, O.make_kernel(spec), mirror=True, lru=0)
    func = S.QColumns(X, y, S.KernelSpec("rbf", 0.05), mirror=True)
    solver = S.SolverWithCache(np.zeros(2 * l), y, 1.0, func=func)
    for i in (3, l + 3, 2 * l - 1):
        np.testing.assert_allclose(solver.get_Q(i).cpu().numpy(), cols(i), rtol=2e-5, atol=2e-6)


# ------------------------------------------------------------------------------ step-wise API
def test_g4_stepwise_trace(golden):
    """select / update one step at a time reproduce the reference's (i, j, delta) trace, including the lowest-index
    tie-break at iteration 0.  All 22 steps are compared.  Where a pair differs from the reference's, the
    test demands that it is a genuine tie: the two candidates' gradients in the reference trajectory (replayed by the
    oracle up to that step) agree to a few ulp, i.e. the choice is decided by the last bit of exp (libm vs device).
    After such a tie the trajectories may part; the solution is then compared within tolerance."""
    torch, _, S, *_ = _mods()
    g = golden("g4_tiny_trace")
    X, y01 = g["X"], g["y"]
    y = np.where(y01 == 1, 1.0, -1.0)
    func = S.QColumns(X, y, S.KernelSpec("rbf", 0.5))
    solver = S.SolverWithCache(-np.ones(8), y, 1.0, 1e-5, func=func)
    trace = g["trace_head"]
    n_ref = int(g["n_iter"])
    assert trace.shape[0] == n_ref == 22
    spec = O.KernelSpec(kind="rbf", gamma=0.5)
    ost = O.fresh_state(-np.ones(8), y, 1.0, 1e-5)
    cols = O.ColumnSource(X, y, O.make_kernel(spec), mirror=False)
    matched = 0
    for k in range(n_ref):
        i, j = solver.working_set_select()
        ri, rj = int(trace[k, 0]), int(trace[k, 1])
        if (i, j) != (ri, rj):
            assert O.select_pair(ost) == (ri, rj)                       # the oracle still follows the reference here
            eps = 8 * np.finfo(np.float64).eps * max(1.0, np.abs(ost.g).max())
            assert abs(ost.g[i] - ost.g[ri]) <= eps and abs(ost.g[j] - ost.g[rj]) <= eps, (k, i, j, ri, rj, ost.g)
            break
        di, dj = solver.update(i, j)
        np.testing.assert_allclose([di, dj], trace[k, 2:], rtol=0, atol=1e-12)
        O.take_step(ost, ri, rj, cols(ri), cols(rj))
        matched += 1
    assert matched >= 8                                                   # the documented exact tie sits at step 8
    st = solver.solve(10 ** 4)
    assert st.converged == 1 and solver.working_set_select() == (-1, -1)
    iters_close(matched + st.n_iter, n_ref, frac=0.15)
    rel_close(solver.alpha.cpu().numpy(), g["alpha"], atol=1e-5)
    rel_close(solver.calculate_rho(), float(g["rho"]), atol=1e-5)


def test_stepwise_equals_fused(breast_cancer):
    torch, _, S, *_ = _mods()
    X, y01 = breast_cancer
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (30 * X.std()))
    a = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, spec))
    b = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, spec))
    for _ in range(40):
        i, j = a.working_set_select()
        a.update(i, j)
    st = b.solve(40)
    assert st.n_iter == 40 and (st.last_i, st.last_j) == a.working_set_select()
    np.testing.assert_array_equal(a.alpha.cpu().numpy(), b.alpha.cpu().numpy())
    np.testing.assert_array_equal(a.neg_y_grad.cpu().numpy(), b.neg_y_grad.cpu().numpy())


def test_update_matches_oracle_on_random_state():
    torch, _, S, *_ = _mods()
    rng = np.random.default_rng(3)
    n, d = 1000, 24
    X = rng.standard_normal((n, d))
    y = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    spec = O.KernelSpec(kind="rbf", gamma=0.04)
    fit = O.fit_csvc(X, (y > 0).astype(int), gamma=0.04, max_iter=150)      # some mid-solve state
    st = O.DualState(y=y, p=-np.ones(n), C=1.0, tol=1e-5, alpha=fit.alpha.copy(), g=fit.g.copy())
    cols = O.ColumnSource(X, y, O.make_kernel(spec), mirror=False)
    solver = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 0.04)))
    solver.alpha = fit.alpha
    solver.neg_y_grad = fit.g
    for _ in range(25):
        i, j = O.select_pair(st)
        assert solver.working_set_select() == (i, j)
        want = O.take_step(st, i, j, cols(i), cols(j))
        got = solver.update(i, j)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(solver.alpha.cpu().numpy(), st.alpha, rtol=0, atol=1e-10)
    np.testing.assert_allclose(solver.neg_y_grad.cpu().numpy(), st.g, rtol=0, atol=1e-9)


# ------------------------------------------------------------------------------ estimators vs golden
def test_g1_breast_cancer_csvc(golden, breast_cancer):
    *_, svc, _, _ = _mods()
    X, y = breast_cancer
    g = golden("g1_breast_cancer_csvc")
    est = svc.BiKernelSVC(max_iter=10 ** 6).fit(X, y)
    iters_close(est.n_iter_, int(g["n_iter"]))
    np.testing.assert_array_equal(est.support_, np.flatnonzero(g["alpha"] > 0))
    rel_close(est.alpha_, g["alpha"])
    rel_close(est.rho_, float(g["rho"]))
    rel_close(est.decision_function(X), g["dec"])
    np.testing.assert_array_equal(est.predict(X), g["pred"])
    # dual objective 0.5 a.(grad f + p), grad f = -y g      (-59.7613453705)
    s = est._fitted.solver
    yv = np.where(y == 1, 1.0, -1.0)
    obj = 0.5 * float(est.alpha_ @ (-yv * s.neg_y_grad.cpu().numpy() - 1.0))
    assert abs(obj - (-59.7613453705)) < 1e-4 * 59.76


@pytest.mark.parametrize("kernel", ["linear", "poly", "sigmoid"])
def test_g6_kernels(golden, breast_cancer, kernel):
    *_, svc, _, _ = _mods()
    X, y = breast_cancer
    g = golden(f"g6_breast_cancer_{kernel}")
    est = svc.BiKernelSVC(kernel=kernel, max_iter=10 ** 6).fit(X, y)
    # For these kernels the reference's own iteration count moves by +-15 % when the feature columns
    # are merely permuted (exact ties after every unclipped step are broken by rounding noise; measured
    # on the oracle: linear 4635..6024, poly 383..457, sigmoid 254..310 -- DESIGN.md section 6), so the
    # count is held to that band and parity is asserted on the solution instead.
    iters_close(est.n_iter_, int(g["n_iter"]), frac=0.25)
    np.testing.assert_array_equal(est.support_, np.flatnonzero(g["alpha"] > 0))
    rel_close(est.alpha_, g["alpha"], atol=2e-4)
    rel_close(est.rho_, float(g["rho"]), atol=1e-5)
    scale = np.abs(g["dec"]).max()
    np.testing.assert_allclose(est.decision_function(X), g["dec"], rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_array_equal(est.predict(X), g["pred"])


def test_g7_nusvc(golden, breast_cancer, capsys):
    *_, svc, _, _ = _mods()
    X, y = breast_cancer
    g = golden("g7_breast_cancer_nusvc")
    est = svc.BiNuSVC(max_iter=1000).fit(X, y)
    assert est.n_iter_ == 1000                                  # D3: the nu solver never stops on tol
    assert "NuSVC not coverage with 1000 iterations" in capsys.readouterr().out
    rel_close(est.alpha_, g["alpha"])
    rel_close(est.rho_, float(g["rho"]))
    rel_close(est.b_, float(g["b"]))
    rel_close(est.decision_function(X), g["dec"], atol=1e-5)
    np.testing.assert_array_equal(est.predict(X), g["pred"])
    assert abs(est.alpha_.sum() - 0.5 * len(y)) < 1e-8


def test_g8_svr(golden, diabetes):
    *_, svr, _ = _mods()
    X, z = diabetes
    for tag, kw in (("svr", {}), ("svr_eps10_C100", dict(eps=10, C=100))):
        g = golden(f"g8_diabetes_{tag}")
        est = svr.KernelSVR(max_iter=10 ** 6, **kw).fit(X, z)
        iters_close(est.n_iter_, int(g["n_iter"]))
        rel_close(est.alpha_, g["alpha"], atol=1e-5 * float(kw.get("C", 1)))
        rel_close(est.rho_, float(g["rho"]))
        rel_close(est.predict(X), g["dec"], atol=1e-3)
    g = golden("g8_diabetes_nusvr")
    est = svr.NuSVR(max_iter=1000).fit(X, z)
    assert est.n_iter_ == 1000
    rel_close(est.alpha_, g["alpha"])
    rel_close(est.b_, float(g["b"]))
    rel_close(est.predict(X), g["dec"], atol=1e-3)


def test_g5_oneclass(golden):
    *_, oc = _mods()
    g = golden("g5_oneclass")
    est = oc.OneClassSVM(nu=.1, gamma=.1, max_iter=10 ** 5).fit(g["X"])
    iters_close(est.n_iter_, int(g["n_iter"]))
    rel_close(est.alpha_, g["alpha"])
    rel_close(est.rho_, float(g["rho"]))
    rel_close(est.decision_function(g["X"]), g["dec"], atol=1e-6)
    np.testing.assert_array_equal(est.predict(g["X"]), g["pred"])


def test_back_to_back_solves_do_not_leak_state(golden, diabetes):
    """A solve must not depend on what ran on the GPU before it: the persistent kernel's shared-memory hand-over
    slots carry epoch tags, and a unit left by an EARLIER launch (same epoch number, other problem) once passed
    for the current one -- the one-class fixture then took 47 instead of 51 iterations after the SVR tests."""
    import hashlib
    *_, svr, oc = _mods()
    g = golden("g5_oneclass")
    Xd, zd = diabetes

    def fit():
        est = oc.OneClassSVM(nu=.1, gamma=.1, max_iter=10 ** 5).fit(g["X"])
        return est.n_iter_, hashlib.sha256(est.alpha_.tobytes()).hexdigest()

    first = fit()
    for kw in (dict(eps=10, C=100), dict(), dict(C=10.)):
        svr.KernelSVR(max_iter=2000, **kw).fit(Xd, zd)          # another problem leaves its state behind
        svr.NuSVR(max_iter=300).fit(Xd[:300], zd[:300])
        assert fit() == first


def test_fit_hands_the_column_cache_back():
    """After fit() the estimator keeps its solver (the reference keeps it inside the decision_function closure), but not
    the Q-column cache: device memory in use returns to (about) what X, alpha and the gradient take; rho and further
    solves on the kept solver still work, and dropping a solver frees its handle at once (no reference cycle).
    (Caches above 1 GB are plain device allocations; smaller ones live in the stream-ordered pool, which keeps them.)"""
    import gc
    import torch
    _, P, S, svc, _, _ = _mods()
    X, y01 = O.synth_classification(n=20000, d=16, seed=5)            # full kernel matrix: 1.6 GB of float32 columns
    X = X.astype(np.float32)
    gc.collect()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    est = svc.BiKernelSVC(max_iter=60000).fit(X, y01)
    assert est.solve_stats_["cache_stores"] > 0                        # the cache was in use during the fit
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 256 << 20, (free0 - free1) >> 20           # X + vectors + pool / allocator slack, not 1.6 GB
    s = est._fitted.solver
    rel_close(s.calculate_rho(), est.rho_)
    st = s.solve(50)                                                    # still usable, now on the cold path
    assert st.cache_stores == 0
    # no cycle: the handle goes when the last reference goes
    yv = np.where(y01 == 1, 1.0, -1.0)
    func = S.QColumns(X, yv, S.KernelSpec("rbf", 0.05), cache_bytes=1500 << 20)
    sol = S.SolverWithCache(-np.ones(len(yv)), yv, 1.0, func=func)
    sol.solve(10)
    torch.cuda.synchronize()
    held, _ = torch.cuda.mem_get_info()
    del sol, func
    torch.cuda.synchronize()
    after, _ = torch.cuda.mem_get_info()
    assert after - held > 1 << 30, (after - held) >> 20


def test_linear_estimators(golden, breast_cancer, diabetes):
    *_, svc, svr, _ = _mods()
    X, y = breast_cancer
    g = golden("lin_breast_cancer_svc")
    est = svc.BiLinearSVC(max_iter=10 ** 6).fit(X, y)
    iters_close(est.n_iter_, int(g["n_iter"]), frac=0.25)        # chaotic count, see test_g6_kernels
    rel_close(est.coef_[0], g["w"], atol=2e-5)
    rel_close(est.decision_function(X), g["dec"], atol=1e-4)
    X, z = diabetes
    g = golden("lin_diabetes_svr")
    est = svr.LinearSVR(max_iter=10 ** 6).fit(X, z)
    iters_close(est.n_iter_, int(g["n_iter"]), frac=0.25)
    rel_close(est.coef_[0], g["w"], atol=1e-3)


def test_g3_digits_ovo(golden, digits):
    """45 one-vs-one sub-problems: per-sub-problem iteration counts within 1 %, identical
    support-vector counts, identical predicted labels."""
    *_, P, S, svc, _, _ = _mods()
    X, y = digits
    g = golden("g3_digits_ovo")
    est = P.KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(X, y)
    subs = est.multiclass_model.estimators_
    iters = np.array([e.n_iter_ for e in subs])
    nsv = np.array([len(e.support_) for e in subs])
    assert np.all(np.abs(iters - g["sub_iters"]) <= np.maximum(1, 0.01 * g["sub_iters"])), (iters, g["sub_iters"])
    np.testing.assert_array_equal(nsv, g["sub_nsv"])
    np.testing.assert_array_equal(est.predict(X), g["pred"])
    rel_close(est.decision_function(X[:32]), g["dec_head"], atol=1e-5)


# ------------------------------------------------------------------------------ float32 synthetic configs
def test_s3_synth_classification_converged(golden):
    """C3-shaped float32 problem at n=2000 to convergence: same SV set, iterations within 1 %."""
    *_, svc, _, _ = _mods()
    X, y = O.synth_classification(n=2000, d=128, seed=0)
    g = golden("s3_synth_cls_n2000")
    est = svc.BiKernelSVC(max_iter=10 ** 7).fit(X, y)
    iters_close(est.n_iter_, int(g["n_iter"]))
    np.testing.assert_array_equal(est.support_, np.flatnonzero(g["alpha"] > 0))
    rel_close(est.alpha_, g["alpha"], atol=1e-4)
    rel_close(est.rho_, float(g["rho"]), atol=1e-5)
    rel_close(est.decision_function(X[:256]), g["dec"], atol=1e-4)


def test_s3_synth_classification_capped(golden):
    """n=20000, max_iter=1000: the early (clipping) phase is reproduced pair for pair."""
    torch, _, S, svc, _, _ = _mods()
    X, y = O.synth_classification(n=20000, d=128, seed=0)
    g = golden("s3_synth_cls_n20000_cap1000")
    est = svc.BiKernelSVC(max_iter=1000).fit(X, y)
    assert est.n_iter_ == 1000
    rel_close(est.alpha_, g["alpha"], atol=1e-4)
    rel_close(est._fitted.solver.neg_y_grad.cpu().numpy(), g["g"], atol=1e-4)


def test_s4_synth_regression(golden):
    *_, svr, _ = _mods()
    X, z = O.synth_regression(n=1000, d=64, seed=0)
    g = golden("s4_synth_svr_n1000")
    est = svr.KernelSVR(max_iter=10 ** 7).fit(X, z)
    iters_close(est.n_iter_, int(g["n_iter"]))
    # float32 kernel values: both runs stop inside the tol=1e-5 band around the same optimum
    rel_close(est.alpha_, g["alpha"], atol=5e-4)
    rel_close(est.predict(X[:256]), g["dec"], atol=2e-4)
    g = golden("s4_synth_nusvr_n1000")
    est = svr.NuSVR(max_iter=2000).fit(X, z)
    assert est.n_iter_ == 2000
    rel_close(est.alpha_, g["alpha"], atol=1e-4)
    rel_close(est.b_, float(g["b"]), atol=1e-5)


def test_s5_synth_oneclass(golden):
    *_, oc = _mods()
    X = O.synth_oneclass(n=2000, d=32, seed=1)
    g = golden("s5_synth_oneclass_n2000")
    est = oc.OneClassSVM(max_iter=10 ** 7).fit(X)
    iters_close(est.n_iter_, int(g["n_iter"]))
    rel_close(est.alpha_, g["alpha"], atol=1e-4)
    rel_close(est.rho_, float(g["rho"]))
    rel_close(est.decision_function(X[:256]), g["dec"], atol=1e-4)


# ------------------------------------------------------------------------------ RFF
def test_rff_vectors(golden):
    _, P, *_ = _mods()
    from pysvm_b200.rff import NormalRFF
    g = golden("rff_vectors")
    np.random.seed(3)
    r = NormalRFF(gamma=0.3, D=32).fit(g["X"])
    np.testing.assert_array_equal(r.w, g["w"])
    np.testing.assert_array_equal(r.b, g["b"])
    np.testing.assert_allclose(r.transform(g["X"]), g["Z"], rtol=0, atol=1e-13)


# ------------------------------------------------------------------------------ BASELINE configs at their REAL sizes
def _sparse_alpha(g):
    a = np.zeros(int(g["n_vars"]))
    a[g["alpha_idx"]] = g["alpha_val"]
    return a


def _replay(solver, trace, upto):
    """One fused iteration per launch: stats.last_i / last_j before the launch is the pair the launch takes, so every
    working pair and step of the device solve is observable and compared with the reference's trace."""
    nxt = solver.working_set_select()
    for k in range(upto):
        assert nxt == (int(trace[k, 0]), int(trace[k, 1])), (k, nxt, trace[k])
        st = solver.solve(1)
        assert st.n_iter == 1
        np.testing.assert_allclose([st.last_delta_i, st.last_delta_j], trace[k, 2:], rtol=1e-4, atol=1e-7)
        nxt = (st.last_i, st.last_j)
    return nxt


def test_s3_full_size_capped_trace(golden):
    """configs[2] at its real size (n=200000, d=128 float32, max_iter=1000, fixture from the live reference): the
    first 300 working pairs are replayed one by one, the remaining 700 iterations run fused; alpha must be equal, the
    gradient within the float32 kernel tolerance, rho and decision values within 1e-4."""
    torch, _, S, svc, _, _ = _mods()
    from pysvm_b200 import synth
    g = golden("s3_full_n200000_cap1000")
    X, y01 = synth.classification(n=200_000, d=128, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    tr = g["trace"]
    func = S.QColumns(X, y, S.KernelSpec("rbf", 1 / (128 * X.std())), cache_bytes=0)
    solver = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=func)
    nxt = _replay(solver, tr, 300)
    st = solver.solve(700)
    assert st.n_iter == 700
    a = solver.alpha.cpu().numpy()
    np.testing.assert_allclose(a, _sparse_alpha(g), rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(np.flatnonzero(a), g["alpha_idx"])                 # identical support set
    rel_close(solver.neg_y_grad.cpu().numpy()[g["g_idx"]], g["g_val"], atol=1e-4)
    rel_close(solver.calculate_rho(), float(g["rho"]), atol=1e-4)
    # the same fit through the estimator: 1000 fused iterations, decision values of the fixture's 256 query rows
    est = svc.BiKernelSVC(max_iter=1000).fit(X, y01)
    assert est.n_iter_ == 1000
    np.testing.assert_array_equal(est.support_, g["alpha_idx"])
    rel_close(est.alpha_[g["alpha_idx"]], g["alpha_val"], atol=1e-6)
    rel_close(est.decision_function(X[:256]), g["dec"], atol=1e-4)


def test_s4_full_size_svr_trace(golden):
    """configs[3] at its real size (l=100000, d=64 -> 200000 variables), KernelSVR, the reference's 200 iterations
    replayed pair by pair (mirror variables: one row sweep serves both halves)."""
    torch, _, S, _, svr, _ = _mods()
    from pysvm_b200 import synth
    g = golden("s4_full_svr_l100000_cap200")
    X, z = synth.regression(n=100_000, d=64, seed=0)
    l = X.shape[0]
    y = np.r_[np.ones(l), -np.ones(l)]
    p = np.r_[-z.astype(np.float64), z.astype(np.float64)]          # eps = 0 (svr.py:166-172)
    func = S.QColumns(X, y, S.KernelSpec("rbf", 1 / (64 * X.std())), mirror=True, cache_bytes=0)
    solver = S.SolverWithCache(p, y, 1.0, func=func)
    _replay(solver, g["trace"], 200)
    a = solver.alpha.cpu().numpy()
    np.testing.assert_allclose(a, _sparse_alpha(g), rtol=1e-4, atol=1e-6)
    rel_close(solver.neg_y_grad.cpu().numpy()[g["g_idx"]], g["g_val"], atol=1e-4)
    est = svr.KernelSVR(max_iter=200).fit(X, z)
    assert est.n_iter_ == 200
    rel_close(est.rho_, float(g["rho"]), atol=1e-4)
    rel_close(est.predict(X[:256]), g["dec"], atol=1e-4)


def test_s5_oneclass_n10000(golden):
    """configs[4]-shaped one-class problem at n=10000 to convergence."""
    *_, oc = _mods()
    from pysvm_b200 import synth
    X = synth.oneclass(n=10_000, d=32, seed=1)
    g = golden("s5_synth_oneclass_n10000")
    est = oc.OneClassSVM(max_iter=10 ** 7).fit(X)
    iters_close(est.n_iter_, int(g["n_iter"]))
    rel_close(est.alpha_, g["alpha"], atol=1e-4)
    rel_close(est.rho_, float(g["rho"]))
    # decision values are sums of ~5000 float32 kernel values of size ~1e2: tolerance relative to that scale
    rel_close(est.decision_function(X[:512]), g["dec"], atol=1e-5 * float(np.abs(g["dec"]).max()))
    far = np.abs(g["dec"]) > 1e-3                       # labels away from the decision boundary
    np.testing.assert_array_equal(est.predict(X[:512])[far], g["pred"][far])


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_invariants():
    """BASELINE.json configs[2] shape (n=200k, d=128, float32): after 2000 fused iterations the
    equality constraint and the box hold exactly, and the incrementally updated gradient equals a
    from-scratch recomputation -y*(Q alpha + p) (size-independent checks; the oracle cannot run here)."""
    torch, _, S, *_ = _mods()
    X, y01 = O.synth_classification(n=200_000, d=128, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    gamma = 1 / (128 * X.std())
    func = S.QColumns(X, y, S.KernelSpec("rbf", gamma))
    solver = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=func)
    st = solver.solve(2000)
    assert st.n_iter == 2000 and st.n_ctas > 100
    a = solver.alpha.cpu().numpy()
    assert a.min() >= 0 and a.max() <= 1.0
    assert abs(float(a @ y)) < 1e-9
    g_inc = solver.neg_y_grad.clone()
    solver._engine.init_gradient(solver.p)
    g_new = solver.neg_y_grad
    assert float((g_inc - g_new).abs().max()) < 1e-3      # 2000 float32 rank-2 updates vs one float32 mat-vec
    # the pair the fused kernel would take next is the arg-max / arg-min of the oracle rule
    stt = O.DualState(y=y, p=-np.ones(len(y)), C=1.0, tol=1e-5, alpha=a, g=g_inc.cpu().numpy())
    assert O.select_pair(stt) == (st.last_i, st.last_j)


# ------------------------------------------------------------------------------ Q-column cache
def test_column_cache_is_numerically_transparent():
    """The HBM column cache must not change a single bit (the reference's LRU is transparent too,
    SURVEY D2): same iterations, same alpha / gradient, and a large share of warm iterations."""
    torch, _, S, *_ = _mods()
    X, y01 = O.synth_classification(n=6000, d=128, seed=2)
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    runs = {}
    for tag, cb in (("off", 0), ("on", None), ("tiny", 40 * 6000 * 4)):
        sol = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, spec, cache_bytes=cb))
        st = sol.solve(10 ** 7)
        runs[tag] = (st.n_iter, sol.alpha.cpu().numpy(), sol.neg_y_grad.cpu().numpy(), st.cache_hits, st.cold_sweeps)
    assert runs["off"][3] == 0 and runs["off"][4] == runs["off"][0]
    for tag in ("on", "tiny"):
        assert runs[tag][0] == runs["off"][0]
        np.testing.assert_array_equal(runs[tag][1], runs["off"][1])
        np.testing.assert_array_equal(runs[tag][2], runs["off"][2])
    assert runs["on"][3] > runs["on"][0]              # > 50 % of the iterations were warm (2 hits each)
    assert 0 < runs["tiny"][3] < runs["on"][3]        # a 40-column cache still helps, but less


def test_column_cache_survives_a_refit_on_the_same_handle():
    torch, _, S, *_ = _mods()
    X, y01 = O.synth_classification(n=4000, d=64, seed=4)
    y = np.where(y01 == 1, 1.0, -1.0)
    sol = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 0.01)))
    a = sol.solve(10 ** 7)
    alpha1 = sol.alpha.clone()
    sol._engine.init_state(sol.p)
    b = sol.solve(10 ** 7)
    assert a.n_iter == b.n_iter and torch.equal(alpha1, sol.alpha)
    assert b.cold_sweeps < a.cold_sweeps          # columns computed by the first fit are reused


def test_batched_solve_keeps_the_cache_bookkeeping_of_every_handle():
    """b2svm_solve_batched fills the column caches of its handles; a later solve on one of those handles must go on
    from the slots already handed out (the device slot table persists), not restart at slot 0 and read another
    row's column.  Re-solving after a batched solve must reproduce a fresh solve bit for bit."""
    torch, _, S, *_ = _mods()
    sols, fresh = [], []
    for seed in (11, 12, 13):
        X, y01 = O.synth_classification(n=700 + 100 * (seed - 11), d=64, seed=seed)
        y = np.where(y01 == 1, 1.0, -1.0)
        mk = lambda: S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 0.01)),
                                       max_iter_hint=10 ** 6)
        sols.append(mk())
        fresh.append(mk())
    stats = S.solve_batched(sols, 10 ** 6)
    for sol, ref, st in zip(sols, fresh, stats):
        r = ref.solve(10 ** 6)
        assert r.n_iter == st.n_iter
        assert torch.equal(ref.alpha, sol.alpha)
        sol._engine.init_state(sol.p)                 # a second fit on the same handle, now through b2svm_solve
        again = sol.solve(10 ** 6)
        assert again.n_iter == r.n_iter and torch.equal(ref.alpha, sol.alpha)
        assert torch.equal(ref.neg_y_grad, sol.neg_y_grad)
        assert again.cold_sweeps <= r.cold_sweeps     # columns cached by the batched solve are found again


# ------------------------------------------------------------------------------ batched one-vs-one
def test_ovo_subproblems_are_solved_in_one_batched_launch(golden, digits):
    """KernelSVC(ovo) on digits: sklearn builds the 45 sub-problems, one persistent launch solves them
    (one CTA each); results equal the sequential per-sub-problem solves bit for bit."""
    torch, P, S, svc, _, _ = _mods()
    X, y = digits
    g = golden("g3_digits_ovo")
    est = P.KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(X, y)
    subs = est.multiclass_model.estimators_
    assert len(subs) == 45 and all(e.solve_stats_["n_ctas"] == 1 for e in subs)
    classes = np.unique(y)
    k = 0
    for a in range(10):
        for b in range(a + 1, 10):
            m = (y == classes[a]) | (y == classes[b])
            seq = svc.BiKernelSVC(max_iter=10 ** 6).fit(X[m], (y[m] == classes[b]).astype(int))
            assert seq.n_iter_ == subs[k].n_iter_
            np.testing.assert_array_equal(seq.alpha_, subs[k].alpha_)
            k += 1
            if k >= 6:
                break
        if k >= 6:
            break
    assert sum(e.n_iter_ for e in subs) == pytest.approx(int(g["n_iter"]), rel=0.01)


def test_readme_accuracy_table_digits_and_breast_cancer():
    """tests/dataset_classify.py of the reference (README.md:128-132), KernelSVC / NuSVC / LinearSVC
    columns for breast_cancer and digits: default 'ovr', max_iter=1000, seed 2022."""
    _, P, *_ = _mods()
    from sklearn.datasets import load_iris, load_breast_cancer, load_digits, load_wine
    from sklearn.preprocessing import StandardScaler
    from sklearn.model_selection import train_test_split
    import contextlib, io
    np.random.seed(2022)
    want = {"iris": (0.94736842, 0.97368421, 0.97368421), "wine": (0.97777778, 0.97777778, 0.97777778),
            "breast_cancer": (0.96503497, 0.96503497, 0.92307692), "digits": (0.95555556, 0.98222222, 0.92222222)}
    for name, load in (("iris", load_iris), ("wine", load_wine), ("breast_cancer", load_breast_cancer),
                       ("digits", load_digits)):
        X, y = load(return_X_y=True)
        trX, teX, trY, teY = train_test_split(X, y)
        sc = StandardScaler().fit(trX)
        trX, teX = sc.transform(trX), sc.transform(teX)
        for j, model in enumerate((P.LinearSVC, P.KernelSVC, P.NuSVC)):
            with contextlib.redirect_stdout(io.StringIO()):
                acc = model(n_jobs=6, max_iter=1000).fit(trX, trY).score(teX, teY)
            # RBF estimators: the README figure exactly (identical test labels).  LinearSVC stops at max_iter=1000
            # far from convergence on a kernel whose SMO trajectory is chaotic: the truncated
            # iterate differs, so its accuracy is held to 1.5 points.
            slack = 0.015 if model is P.LinearSVC else 1e-8
            assert abs(acc - want[name][j]) <= slack + 1e-9, (name, model.__name__, acc, want[name][j])


# ------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n,d,dtype", [(3, 1, np.float64), (33, 5, np.float32), (97, 130, np.float32),
                                       (64, 512, np.float32), (50, 256, np.float64), (1000, 17, np.float64)])
def test_ragged_shapes_match_oracle(n, d, dtype):
    """Row counts that are not a multiple of the 32-row tile, feature counts that need padding, the widest
    compiled rows (512 float32 / 256 float64): same trajectory as the oracle."""
    *_, svc, _, _ = _mods()
    rng = np.random.default_rng(n * 1000 + d)
    X = rng.standard_normal((n, d)).astype(dtype)
    y = (rng.random(n) < 0.5).astype(int)
    y[0], y[1] = 1, 0
    ref = O.fit_csvc(X, y, max_iter=10 ** 5)
    est = svc.BiKernelSVC(max_iter=10 ** 5).fit(X, y)
    iters_close(est.n_iter_, ref.n_iter, frac=0.02)
    rel_close(est.alpha_, ref.alpha, atol=2e-4)
    np.testing.assert_array_equal(est.predict(X), ref.predict(X))


def test_single_class_stops_immediately():
    """All labels equal: I_low is empty, the reference's select returns (-1, -1) at once (solver.py:66-75)."""
    *_, svc, _, _ = _mods()
    X = np.random.default_rng(0).standard_normal((40, 3))
    est = svc.BiKernelSVC(max_iter=100).fit(X, np.ones(40, dtype=int))
    assert est.n_iter_ == 0 and est.solve_stats_["converged"] == 1 and est.alpha_.sum() == 0.0
    with pytest.raises(ValueError):          # the reference's calculate_rho: min() of an empty set
        O.fit_csvc(X, np.ones(40, dtype=int), max_iter=100)
    with pytest.raises(ValueError):          # ... which surfaces at decision time there, and here
        est.decision_function(X)


def test_too_many_features_is_an_error_not_a_fallback():
    torch, _, S, *_ = _mods()
    from pysvm_b200._native import B2SVMError
    X = np.zeros((8, 600), dtype=np.float32)
    y = np.array([1, -1] * 4, dtype=float)
    with pytest.raises(B2SVMError, match="n_features"):
        S.SolverWithCache(-np.ones(8), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 0.1)))


def test_max_iter_zero_and_print(capsys):
    *_, svc, _, _ = _mods()
    X = np.random.default_rng(1).standard_normal((20, 2))
    y = np.array([0, 1] * 10)
    est = svc.BiKernelSVC(max_iter=0).fit(X, y)
    assert est.n_iter_ == 0
    assert "KernelSVC not coverage with 0 iterations" in capsys.readouterr().out      # for/else of svc.py:260-267


# ------------------------------------------------------------------------------ precomputed kernel: RFF (any D), dense Q
def test_g6_rff_default_D(golden, breast_cancer):
    """BiKernelSVC(rff=True, D=1000): same RNG draws as the reference (seed 7), kernel z(a).z(b) handed to
    the solver as a precomputed matrix."""
    *_, svc, _, _ = _mods()
    X, y = breast_cancer
    g = golden("g6_breast_cancer_rff")
    np.random.seed(7)
    est = svc.BiKernelSVC(rff=True, D=1000, max_iter=10 ** 6).fit(X, y)
    iters_close(est.n_iter_, int(g["n_iter"]), frac=0.05)
    np.testing.assert_array_equal(est.support_, np.flatnonzero(g["alpha"] > 0))
    rel_close(est.alpha_, g["alpha"], atol=2e-4)
    rel_close(est.rho_, float(g["rho"]), atol=1e-5)
    rel_close(est.decision_function(X), g["dec"], atol=1e-4)
    np.testing.assert_array_equal(est.predict(X), g["pred"])


def test_dense_q_solver_equals_column_solver(breast_cancer):
    """Solver(Q, p, y, C): the reference's cache_size == 0 path (svc.py:251-253) gives the same trajectory."""
    torch, _, S, *_ = _mods()
    X, y01 = breast_cancer
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = O.KernelSpec("rbf", 1 / (30 * X.std()))
    K = O.make_kernel(spec)(X, X)
    Q = y.reshape(-1, 1) * y * K
    dense = S.Solver(Q, -np.ones(len(y)), y, 1.0, 1e-5)
    sd = dense.solve(10 ** 6)
    col = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", spec.gamma)))
    sc = col.solve(10 ** 6)
    iters_close(sd.n_iter, sc.n_iter)
    assert sd.cold_sweeps == 0                       # a kernel matrix is a full cache: never a sweep of X
    rel_close(dense.alpha.cpu().numpy(), col.alpha.cpu().numpy(), atol=1e-6)
    np.testing.assert_allclose(dense.get_Q(5).cpu().numpy(), Q[:, 5], rtol=1e-12, atol=1e-14)


def test_readme_regression_table_diabetes():
    """tests/dataset_regression.py of the reference (README.md:136-140), diabetes column.  load_boston is
    gone from scikit-learn, so its split's RNG consumption is emulated: the reference
    then gives R^2 = 0.4537 / 0.1756 / 0.1459 for LinearSVR / KernelSVR / NuSVR at max_iter=1000."""
    _, P, *_ = _mods()
    from sklearn.datasets import load_diabetes
    from sklearn.preprocessing import StandardScaler
    from sklearn.model_selection import train_test_split
    import contextlib, io
    np.random.seed(2022)
    train_test_split(np.zeros((506, 13)), np.zeros(506))           # the boston split that ran first
    X, y = load_diabetes(return_X_y=True)
    trX, teX, trY, teY = train_test_split(X, y)
    sc = StandardScaler().fit(trX)
    trX, teX = sc.transform(trX), sc.transform(teX)
    want = {"LinearSVR": 0.4537, "KernelSVR": 0.1756, "NuSVR": 0.1459}
    for model in (P.LinearSVR, P.KernelSVR, P.NuSVR):
        with contextlib.redirect_stdout(io.StringIO()):
            r2 = model(max_iter=1000).fit(trX, trY).score(teX, teY)
        assert abs(r2 - want[model.__name__]) < 2e-3, (model.__name__, r2)


# ------------------------------------------------------------------------------ K7 on the tensor cores
def _matvec_truth(torch, kind, gamma, coef0, degree, Xa, coef, Xb):
    a, q = Xa.double(), Xb.double()
    dot = q @ a.T
    if kind == "rbf":
        k = torch.exp(-gamma * ((q * q).sum(1)[:, None] + (a * a).sum(1)[None, :] - 2 * dot).clamp_min(0))
    elif kind == "poly":
        k = (gamma * dot + coef0) ** degree
    elif kind == "sigmoid":
        k = torch.tanh(gamma * dot + coef0)
    else:
        k = dot
    return k @ coef.double()


@pytest.mark.parametrize("na,nb,d,kind", [(1000, 777, 32, "rbf"), (2500, 1300, 64, "rbf"), (3333, 700, 7, "rbf"),
                                          (2048, 1500, 33, "poly"), (2048, 900, 48, "sigmoid"), (5000, 129, 20, "linear"),
                                          (130, 3000, 100, "rbf"), (1111, 2222, 128, "rbf"), (1, 300, 16, "rbf")])
def test_tensor_core_matvec_matches_float64(monkeypatch, na, nb, d, kind):
    """b2svm_kernel_matvec through the tcgen05 3xTF32 path (forced) and through the FFMA path against a float64
    evaluation of svc.py:230-241 + the coefficient product of svc.py:269-272.  Tolerance 2e-6 of the output
    scale: float32 kernel values, float64 accumulation -- both device paths sit at a few 1e-7."""
    torch, _, S, *_ = _mods()
    g = torch.Generator().manual_seed(na + nb + d)
    dev = torch.device("cuda:0")
    Xa, Xb = torch.randn(na, d, generator=g).to(dev), torch.randn(nb, d, generator=g).to(dev)
    coef = torch.randn(na, generator=g, dtype=torch.float64).to(dev)
    if na > 4:
        coef[torch.rand(na, generator=g).to(dev) < 0.4] = 0.0
    spec = S.KernelSpec(kind, gamma=1.0 / d, coef0=0.5 if kind in ("poly", "sigmoid") else 0.0, degree=3)
    ref = _matvec_truth(torch, kind, spec.gamma, spec.coef0, 3, Xa, coef, Xb)
    scale = max(ref.abs().max().item(), 1e-30)
    for mode in ("2", "0"):
        monkeypatch.setenv("B2SVM_MATVEC_TC", mode)
        out = S.kernel_matvec(spec, Xa, coef, Xb)
        assert (out - ref).abs().max().item() / scale < 2e-6, (mode, kind)


def test_decision_function_is_the_same_on_both_matvec_paths(monkeypatch, breast_cancer):
    """A fitted KernelSVC scores the same rows through the tensor-core and the FFMA mat-vec (1e-5 relative)."""
    _, _, _, svc, *_ = _mods()
    X, y = breast_cancer
    X = np.ascontiguousarray(X, dtype=np.float32)
    est = svc.BiKernelSVC(max_iter=2000).fit(X, y)
    monkeypatch.setenv("B2SVM_MATVEC_TC", "2")
    a = np.asarray(est.decision_function(X))
    monkeypatch.setenv("B2SVM_MATVEC_TC", "0")
    b = np.asarray(est.decision_function(X))
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
    assert (np.sign(a) == np.sign(b)).all()


# ------------------------------------------------------------------------------ narrow rows: several tiles per ring stage
@pytest.mark.parametrize("n,d,dtype,max_ctas", [(20011, 32, np.float32, 2), (20011, 32, np.float32, 148),
                                                (9000, 10, np.float32, 1), (7001, 16, np.float64, 3),
                                                (5000, 50, np.float32, 2)])
def test_macro_tiles_are_bit_identical(monkeypatch, n, d, dtype, max_ctas):
    """Grouping 2-8 tiles of narrow rows into one ring stage (one bulk copy, one ticket) changes how X travels,
    not what is computed: 400 iterations give bit-identical alpha / gradient with and without grouping, with the
    slice resident in shared memory (many CTAs) and streamed through the ring (few CTAs)."""
    torch, _, S, *_ = _mods()
    X, y01 = O.synth_classification(n=n, d=d, seed=3)
    X = X.astype(dtype)
    y = np.where(y01 == 1, 1.0, -1.0)
    monkeypatch.setenv("B2SVM_MAX_CTAS", str(max_ctas))
    out = []
    for sub in ("1", "2", "8"):
        monkeypatch.setenv("B2SVM_SUB", sub)
        sol = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 1.0 / d), cache_bytes=0))
        st = sol.solve(400)
        assert st.n_iter == 400
        out.append((sol.alpha.cpu().numpy().copy(), sol.neg_y_grad.cpu().numpy().copy()))
    for a, g in out[1:]:
        assert np.array_equal(a, out[0][0]) and np.array_equal(g, out[0][1])


@pytest.mark.parametrize("nq,d,D", [(3000, 30, 1000), (1500, 32, 4096), (2049, 7, 300), (1100, 100, 512)])
def test_rff_decision_on_tensor_cores(monkeypatch, nq, d, D):
    """z(x_q) . wz (rff.py:19-20 + the coefficient product): the tcgen05 GEMM + cos epilogue (float32 queries,
    3xTF32, argument error ~1e-6) against the float64 kernel, 2e-5 of the output scale; both against NumPy."""
    torch, _, S, *_ = _mods()
    from pysvm_b200.rff import NormalRFF, rff_decision_device
    rng = np.random.default_rng(nq + D)
    Xq = rng.standard_normal((nq, d)).astype(np.float32)
    np.random.seed(3)
    r = NormalRFF(gamma=1.0 / d, D=D).fit(Xq[:1])
    wz = rng.standard_normal(D)
    want = (np.sqrt(2 / D) * np.cos(Xq.astype(np.float64) @ r.w.T + r.b)) @ wz
    dev = torch.device("cuda:0")
    w, b = r._device_params(dev)
    outs = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("B2SVM_RFF_TC", mode)
        outs[mode] = rff_decision_device(torch.from_numpy(Xq).to(dev), w, b, torch.from_numpy(wz).to(dev)).cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(outs["0"] - want).max() / scale < 1e-12
    assert np.abs(outs["2"] - want).max() / scale < 2e-5


@pytest.mark.parametrize("variant", ["svr", "nusvc"])
def test_macro_tiles_generic_kernel(monkeypatch, variant):
    """The same grouping under the run-time-generic kernel: mirror problems (2l variables share l rows) and the nu
    variant (four candidate types).  Grouped and ungrouped runs must agree bit for bit."""
    torch, _, S, *_ = _mods()
    monkeypatch.setenv("B2SVM_MAX_CTAS", "3")
    out = []
    for sub in ("1", "4"):
        monkeypatch.setenv("B2SVM_SUB", sub)
        if variant == "svr":
            X, z = O.synth_regression(n=6001, d=24, seed=4)
            l = X.shape[0]
            y = np.r_[np.ones(l), -np.ones(l)]
            p = np.r_[-z.astype(float), z.astype(float)]
            func = S.QColumns(X, y, S.KernelSpec("rbf", 1.0 / 24), mirror=True, cache_bytes=0)
            sol = S.SolverWithCache(p, y, 1.0, func=func)
        else:
            X, y01 = O.synth_classification(n=5003, d=20, seed=5)
            y = np.where(y01 == 1, 1.0, -1.0)
            func = S.QColumns(X, y, S.KernelSpec("rbf", 1.0 / 20), cache_bytes=0)
            sol = S.NuSolverWithCache(np.zeros(len(y)), y, 0.5 * len(y), 1.0, func)
        st = sol.solve(300)
        assert st.n_iter == 300
        out.append((sol.alpha.cpu().numpy().copy(), sol.neg_y_grad.cpu().numpy().copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


# ------------------------------------------------------------------------------ dense kernel matrix / dense-Q mode / K8 on tensor cores
@pytest.mark.parametrize("kind", ["rbf", "linear", "poly", "sigmoid"])
@pytest.mark.parametrize("na,nb,d,dtype", [(70, 130, 30, np.float64), (1500, 1100, 64, np.float32), (2000, 2000, 33, np.float32),
                                           (300, 40, 100, np.float32)])
def test_kernel_matrix_matches_oracle(kind, na, nb, d, dtype):
    """b2svm_kernel_matrix against the oracle's kernel_func (svc.py:230-241): float64 to 1e-12; float32 on the FMA
    kernel to float32 rounding, on the tensor cores (>= 1024 rows, d <= 96: 3xTF32) to ~1e-6 of the value scale."""
    torch, _, S, *_ = _mods()
    rng = np.random.default_rng(na + d)
    A = rng.standard_normal((na, d)).astype(dtype)
    B = rng.standard_normal((nb, d)).astype(dtype)
    spec = O.KernelSpec(kind=kind, gamma=0.5 / d, coef0=0.3, degree=3)
    want = O.make_kernel(spec)(A.astype(np.float64), B.astype(np.float64))
    dev = torch.device("cuda")
    got = S.kernel_matrix(S.KernelSpec(kind, spec.gamma, spec.coef0, spec.degree), torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev))
    assert got.shape == (na, nb) and got.dtype == (torch.float64 if dtype == np.float64 else torch.float32)
    tol = 1e-12 if dtype == np.float64 else 3e-6
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=tol, atol=tol * max(1.0, np.abs(want).max()))


def test_dense_q_mode_of_the_estimators(breast_cancer, golden):
    """cache_size=0 is the reference's dense-Q mode (svc.py:251-253): the kernel matrix is built once, on the device,
    and the solve reads its columns; the fit must equal the column-mode fit and the reference's golden vectors."""
    *_, svc, svr, oc = _mods()
    X, y = breast_cancer
    g = golden("g1_breast_cancer_csvc")
    dense = svc.BiKernelSVC(max_iter=10 ** 6, cache_size=0).fit(X, y)
    cols = svc.BiKernelSVC(max_iter=10 ** 6).fit(X, y)
    iters_close(dense.n_iter_, int(g["n_iter"]))
    np.testing.assert_array_equal(dense.support_, np.flatnonzero(g["alpha"] > 0))
    rel_close(dense.alpha_, g["alpha"])
    rel_close(dense.decision_function(X), g["dec"])
    rel_close(dense.alpha_, cols.alpha_, atol=1e-9)
    # float32 rows large enough for the tensor-core Gram kernel (3xTF32: kernel values to ~1e-6)
    Xs, ys = O.synth_classification(n=3000, d=64, seed=5)
    ref = O.fit_csvc(Xs, ys, max_iter=10 ** 6)
    est = svc.BiKernelSVC(max_iter=10 ** 6, cache_size=0).fit(Xs, ys)
    iters_close(est.n_iter_, ref.n_iter, frac=0.02)
    np.testing.assert_array_equal(est.support_, np.flatnonzero(ref.alpha > 0))
    rel_close(est.alpha_, ref.alpha, atol=2e-4)
    np.testing.assert_array_equal(est.predict(Xs), ref.predict(Xs))


@pytest.mark.parametrize("n,d,D,out", [(5000, 32, 512, "f64"), (3000, 20, 1000, "f64"), (4096, 32, 4096, "f32"), (1100, 64, 130, "f64")])
def test_rff_transform_on_tensor_cores(n, d, D, out):
    """NormalRFF.transform for float32 X (configs[4]: d = 32, D = 4096) as a tcgen05 GEMM + cos epilogue + TMA stores.
    Against rff.py:19-20 evaluated in float64: the cosine argument carries the 3xTF32 error (~1e-6), so the features
    agree to 8e-6 of their scale sqrt(2/D) -- stated, not bit parity; float64 X stays on the exact kernel (test_rff_vectors)."""
    torch, *_ = _mods()
    from pysvm_b200.rff import NormalRFF
    rng = np.random.default_rng(n + D)
    X = rng.standard_normal((n, d)).astype(np.float32)
    np.random.seed(1)
    r = NormalRFF(gamma=1.0 / d, D=D).fit(X[:1])
    want = np.sqrt(2 / D) * np.cos(X.astype(np.float64) @ r.w.T + r.b)
    Z = r.transform_device(torch.from_numpy(X).cuda(), dtype=torch.float32 if out == "f32" else torch.float64)
    assert Z.shape == (n, D) and Z.dtype == (torch.float32 if out == "f32" else torch.float64)
    np.testing.assert_allclose(Z.cpu().numpy().astype(np.float64), want, rtol=0, atol=8e-6 * np.sqrt(2 / D) + (1e-7 if out == "f32" else 0))
    # and the float64 CUDA-core kernel on the same float32 input is exact
    import os
    os.environ["B2SVM_RFF_TC"] = "0"
    try:
        Ze = r.transform_device(torch.from_numpy(X).cuda())
    finally:
        del os.environ["B2SVM_RFF_TC"]
    np.testing.assert_allclose(Ze.cpu().numpy(), want, rtol=0, atol=1e-13)


def test_weighted_rows_is_x_transposed_times_w():
    torch, _, S, *_ = _mods()
    rng = np.random.default_rng(9)
    for n, d, dtype in ((1000, 7, np.float64), (50000, 130, np.float32), (3, 1000, np.float64)):
        X = rng.standard_normal((n, d)).astype(dtype)
        w = rng.standard_normal(n)
        got = S.weighted_rows(torch.from_numpy(X).cuda(), torch.from_numpy(w).cuda()).cpu().numpy()
        np.testing.assert_allclose(got, X.astype(np.float64).T @ w, rtol=1e-12, atol=1e-10)


# ------------------------------------------------------------------------------ opt-in second-order selection (WSS 3)
def test_wss3_opt_in_pairs_and_optimum(breast_cancer):
    """wss="wss3" is an opt-in (north_star names LIBSVM's second-order rule; the reference itself uses the maximal
    violating pair).  The device selection must pick the oracle's WSS-3 pair step for step, reach the same optimum as
    the default rule, and need fewer iterations -- which it reports as its own count."""
    torch, _, S, svc, _, _ = _mods()
    X, y01 = breast_cancer
    y = np.where(y01 == 1, 1.0, -1.0)
    n = len(y)
    gamma = 1 / (30 * X.std())
    spec = O.KernelSpec(kind="rbf", gamma=gamma)
    cols = O.ColumnSource(X, y, O.make_kernel(spec), mirror=False)
    kdiag = np.array([cols(t)[t] for t in range(n)])
    st = O.fresh_state(-np.ones(n), y, 1.0, 1e-5)
    solver = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", gamma)))
    for k in range(60):
        i, j = O.select_pair_wss3(st, cols, kdiag)
        assert solver.working_set_select_wss3() == (i, j), k
        want = O.take_step(st, i, j, cols(i), cols(j))
        np.testing.assert_allclose(solver.update(i, j), want, rtol=1e-9, atol=1e-12)
    mvp = svc.BiKernelSVC(max_iter=10 ** 6).fit(X, y01)
    w3 = svc.BiKernelSVC(max_iter=10 ** 6, wss="wss3").fit(X, y01)
    assert w3.n_iter_ < mvp.n_iter_                           # 487 for the reference's rule
    np.testing.assert_array_equal(w3.support_, mvp.support_)
    rel_close(w3.alpha_, mvp.alpha_, atol=2e-4)
    rel_close(w3.rho_, mvp.rho_, atol=1e-4)
    np.testing.assert_array_equal(w3.predict(X), mvp.predict(X))
    print("WSS3 iterations", w3.n_iter_, "vs maximal violating pair", mvp.n_iter_)


def test_column_cache_evicts_and_stays_transparent():
    """A cache far smaller than the set of columns the solve touches: once full, the oldest column is replaced (the
    reference's LRU evicts too, solver.py:207,227-229).  Results stay bit-identical to the cache-less solve, later hits
    exist that a fill-only cache could not have (columns first seen after the cache was full), and a second fit on the
    same handle starts from the adapted cache."""
    torch, _, S, *_ = _mods()
    X, y01 = O.synth_classification(n=6000, d=128, seed=2)
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    off = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, spec, cache_bytes=0))
    s0 = off.solve(10 ** 6)
    slots = 600
    small = S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=S.QColumns(X, y, spec, cache_bytes=slots * 6000 * 4))
    s1 = small.solve(10 ** 6)
    assert s1.n_iter == s0.n_iter
    assert torch.equal(off.alpha, small.alpha) and torch.equal(off.neg_y_grad, small.neg_y_grad)
    assert s1.cache_stores > slots and s1.cache_evictions == s1.cache_stores - slots     # the full cache kept replacing
    assert s1.cache_hits > 0
    small._engine.init_state(small.p)
    s2 = small.solve(10 ** 6)
    assert s2.n_iter == s0.n_iter and torch.equal(off.alpha, small.alpha)
    # two emulated ranks keep their replicated slot tables in step while evicting
    from pysvm_b200 import sharded as SH
    ranks = SH.EmulatedRanks(X, y, -np.ones(len(y)), 1.0, 1e-5, spec, 2, cache_bytes=slots * 3008 * 4)
    st = ranks.solve(10 ** 6)
    assert all(t.n_iter == s0.n_iter for t in st) and st[0].cache_hits == st[1].cache_hits > 0
    assert st[0].cache_evictions == st[1].cache_evictions > 0
    np.testing.assert_array_equal(ranks.gather("alpha"), off.alpha.cpu().numpy())
