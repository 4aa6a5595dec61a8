"""The oracle (oracle/smo_oracle.py) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only.  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import smo_oracle as O

TIGHT = dict(rtol=0, atol=1e-12)


def check(fit, gold, Xq, keep=64):
    assert fit.n_iter == int(gold["n_iter"])
    tr = np.array(fit.trace, dtype=np.float64).reshape(-1, 4)
    k = gold["trace_head"].shape[0]
    np.testing.assert_array_equal(tr[:k, :2], gold["trace_head"][:, :2])       # same working pairs
    np.testing.assert_allclose(tr[:k, 2:], gold["trace_head"][:, 2:], **TIGHT)  # same steps
    np.testing.assert_allclose(fit.alpha, gold["alpha"], **TIGHT)
    np.testing.assert_allclose(fit.g, gold["g"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(fit.rho, float(gold["rho"]), rtol=0, atol=1e-10)
    np.testing.assert_allclose(fit.decision_function(Xq), gold["dec"], rtol=1e-10, atol=1e-9)


def test_g1_breast_cancer(golden, breast_cancer):
    X, y = breast_cancer
    g = golden("g1_breast_cancer_csvc")
    fit = O.fit_csvc(X, y, max_iter=10 ** 6, trace=True)
    check(fit, g, X)
    assert fit.n_iter == 487 and int((fit.alpha > 0).sum()) == 119
    np.testing.assert_array_equal(fit.predict(X), g["pred"])
    p = -np.ones(X.shape[0])
    assert abs(O.dual_objective(fit, p) - (-59.7613453705)) < 1e-8


def test_g4_tiny_trace(golden):
    g = golden("g4_tiny_trace")
    fit = O.fit_csvc(g["X"], g["y"], gamma=0.5, trace=True)
    check(fit, g, g["X"])
    assert fit.trace[0][:2] == (0, 1)            # lowest-index tie-break at iteration 0
    assert fit.n_iter == 22


def test_g5_oneclass(golden):
    g = golden("g5_oneclass")
    fit = O.fit_oneclass(g["X"], nu=.1, gamma=.1, max_iter=10 ** 5, trace=True)
    check(fit, g, g["X"])
    np.testing.assert_array_equal(fit.predict(g["X"]), g["pred"])
    assert abs(fit.alpha.sum() - 20.0) < 1e-9


@pytest.mark.parametrize("kernel", ["linear", "poly", "sigmoid"])
def test_g6_kernels(golden, breast_cancer, kernel):
    X, y = breast_cancer
    fit = O.fit_csvc(X, y, kernel=kernel, max_iter=10 ** 6, trace=True)
    check(fit, golden(f"g6_breast_cancer_{kernel}"), X)


def test_g6_rff(golden, breast_cancer):
    X, y = breast_cancer
    np.random.seed(7)
    fit = O.fit_csvc(X, y, rff=True, D=1000, max_iter=10 ** 6, trace=True)
    check(fit, golden("g6_breast_cancer_rff"), X)


def test_g7_nusvc(golden, breast_cancer):
    X, y = breast_cancer
    g = golden("g7_breast_cancer_nusvc")
    fit = O.fit_nusvc(X, y, max_iter=1000, trace=True)
    check(fit, g, X)
    assert abs(fit.b - float(g["b"])) < 1e-10
    assert fit.n_iter == 1000 and not fit.converged          # D3: never stops on tol
    np.testing.assert_array_equal(fit.predict(X), g["pred"])


def test_g8_svr(golden, diabetes):
    X, z = diabetes
    check(O.fit_svr(X, z, max_iter=10 ** 6, trace=True), golden("g8_diabetes_svr"), X)
    check(O.fit_svr(X, z, eps=10, C=100, max_iter=10 ** 6, trace=True),
          golden("g8_diabetes_svr_eps10_C100"), X)
    g = golden("g8_diabetes_nusvr")
    fit = O.fit_nusvr(X, z, max_iter=1000, trace=True)
    check(fit, g, X)
    assert abs(fit.b - float(g["b"])) < 1e-9


def test_g3_digits_ovo(golden, digits):
    X, y = digits
    g = golden("g3_digits_ovo")
    classes = np.unique(y)
    iters, nsv = [], []
    for a in range(10):
        for b in range(a + 1, 10):
            m = (y == classes[a]) | (y == classes[b])
            fit = O.fit_csvc(X[m], (y[m] == classes[b]).astype(int), max_iter=10 ** 6)
            iters.append(fit.n_iter)
            nsv.append(int((fit.alpha > 0).sum()))
    np.testing.assert_array_equal(iters, g["sub_iters"])
    np.testing.assert_array_equal(nsv, g["sub_nsv"])
    assert sum(iters) == 21571


def test_s3_synth_classification(golden):
    X, y = O.synth_classification(n=2000, d=128, seed=0)
    assert X.dtype == np.float32
    g = golden("s3_synth_cls_n2000")
    check(O.fit_csvc(X, y, max_iter=10 ** 7, trace=True), g, X[:256])


def test_s4_synth_regression(golden):
    X, z = O.synth_regression(n=1000, d=64, seed=0)
    check(O.fit_svr(X, z, max_iter=10 ** 7, trace=True), golden("s4_synth_svr_n1000"), X[:256])
    g = golden("s4_synth_nusvr_n1000")
    fit = O.fit_nusvr(X, z, max_iter=2000, trace=True)
    check(fit, g, X[:256])
    assert abs(fit.b - float(g["b"])) < 1e-10


def test_s5_synth_oneclass(golden):
    X = O.synth_oneclass(n=2000, d=32, seed=1)
    check(O.fit_oneclass(X, max_iter=10 ** 7, trace=True), golden("s5_synth_oneclass_n2000"), X[:256])


def test_linear_estimators(golden, breast_cancer, diabetes):
    X, y = breast_cancer
    g = golden("lin_breast_cancer_svc")
    fit = O.fit_linear_svc(X, y, max_iter=10 ** 6, trace=True)
    assert fit.n_iter == int(g["n_iter"])
    np.testing.assert_allclose(fit.w, g["w"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(fit.decision_function(X), g["dec"], rtol=0, atol=1e-9)
    X, z = diabetes
    g = golden("lin_diabetes_svr")
    fit = O.fit_linear_svr(X, z, max_iter=10 ** 6, trace=True)
    assert fit.n_iter == int(g["n_iter"])
    np.testing.assert_allclose(fit.w, g["w"], rtol=0, atol=1e-9)


def test_rff_vectors(golden):
    g = golden("rff_vectors")
    np.random.seed(3)
    w, b = O.draw_rff(0.3, 32, 6)
    np.testing.assert_array_equal(w, g["w"])
    np.testing.assert_array_equal(b, g["b"])
    np.testing.assert_allclose(O.rff_features(g["X"], w, b), g["Z"], rtol=0, atol=1e-15)


def test_oneclass_alpha0_quirk():
    a = O.oneclass_alpha0(0.1, 200)          # D9: n = 20 -> a[19] = 0.0, a[20] = 1
    assert a[:19].tolist() == [1.0] * 19 and a[19] == 0.0 and a[20] == 1.0 and a[21:].sum() == 0
    with pytest.raises(UnboundLocalError):
        O.oneclass_alpha0(0.001, 100)


# ------------------------------------------------------------------------------ BASELINE configs at their real sizes
def test_s3_full_size_head(golden):
    """configs[2] at n=200000: the oracle's first 40 working pairs / steps equal the reference's (the full 1000
    iterations of the fixture take minutes on the host; the GPU parity test replays all of them)."""
    from pysvm_b200 import synth
    g = golden("s3_full_n200000_cap1000")
    X, y = synth.classification(n=200_000, d=128, seed=0)
    fit = O.fit_csvc(X, y, max_iter=40, trace=True)
    tr = np.array(fit.trace, dtype=np.float64).reshape(-1, 4)
    np.testing.assert_array_equal(tr[:, :2], g["trace"][:40, :2])
    np.testing.assert_allclose(tr[:, 2:], g["trace"][:40, 2:], **TIGHT)


def test_s4_full_size_svr(golden):
    """configs[3] at l=100000 (200000 variables), KernelSVR capped at 200 iterations: all pairs, steps, alpha, gradient."""
    from pysvm_b200 import synth
    g = golden("s4_full_svr_l100000_cap200")
    X, z = synth.regression(n=100_000, d=64, seed=0)
    fit = O.fit_svr(X, z, max_iter=200, trace=True)
    tr = np.array(fit.trace, dtype=np.float64).reshape(-1, 4)
    np.testing.assert_array_equal(tr[:, :2], g["trace"][:, :2])
    np.testing.assert_allclose(tr[:, 2:], g["trace"][:, 2:], **TIGHT)
    alpha = np.zeros(int(g["n_vars"]))
    alpha[g["alpha_idx"]] = g["alpha_val"]
    np.testing.assert_allclose(fit.alpha, alpha, **TIGHT)
    np.testing.assert_allclose(fit.g[g["g_idx"]], g["g_val"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(fit.rho, float(g["rho"]), rtol=0, atol=1e-10)


def test_s5_oneclass_n10000(golden):
    """configs[4]-shaped one-class problem at n=10000, converged (2545 iterations in the reference)."""
    from pysvm_b200 import synth
    g = golden("s5_synth_oneclass_n10000")
    X = synth.oneclass(n=10_000, d=32, seed=1)
    fit = O.fit_oneclass(X, max_iter=10 ** 7, trace=True)
    check(fit, g, X[:512])
    np.testing.assert_array_equal(fit.predict(X[:512]), g["pred"])
