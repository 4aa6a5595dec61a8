"""Regenerate tests/golden/*.npz by running the UNMODIFIED reference (Kaslanarian/PySVM).

Run in the build container only (the reference lives at /root/reference there and does not
travel to the GPU box):

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py g1 s3      # a subset

Solver internals are reached the only way the reference allows: wrapping
``Solver.update`` / ``NuSolver.update`` to log every step, and reading the fitted solver out of
the ``decision_function`` closure.  Nothing under /root/reference is modified.
"""
import os
import sys
import contextlib
import io

import numpy as np

REF = os.environ.get("PYSVM_REF", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import pysvm                                    # noqa: E402  (the reference)
from pysvm import solver as ref_solver          # noqa: E402
from pysvm.svm.svc import BiKernelSVC, BiNuSVC, BiLinearSVC   # noqa: E402
from pysvm.svm.svr import KernelSVR, NuSVR, LinearSVR         # noqa: E402
from pysvm.svm.oc_svm import OneClassSVM                      # noqa: E402
from pysvm import KernelSVC                                   # noqa: E402
from oracle import smo_oracle as O                            # noqa: E402  (generators only)

assert os.path.realpath(pysvm.__file__).startswith(os.path.realpath(REF)), pysvm.__file__

TRACE = []


def _wrap(cls):
    orig = cls.update

    def logged(self, i, j, *a, **k):
        di, dj = orig(self, i, j, *a, **k)
        TRACE.append((int(i), int(j), float(di), float(dj)))
        return di, dj

    cls.update = logged


_wrap(ref_solver.Solver)
_wrap(ref_solver.NuSolver)
# SolverWithCache / NuSolverWithCache call super().update -> logged once


def solver_of(est):
    for cell in est.decision_function.__closure__:
        v = cell.cell_contents
        if isinstance(v, ref_solver.Solver):
            return v
    raise RuntimeError("no solver in closure")


def closure_float(est, name):
    fn = est.decision_function
    names = fn.__code__.co_freevars
    return fn.__closure__[names.index(name)].cell_contents


def run(est, *fit_args):
    TRACE.clear()
    with contextlib.redirect_stdout(io.StringIO()) as out:
        est.fit(*fit_args)
    s = solver_of(est)
    return s, np.array(TRACE, dtype=np.float64).reshape(-1, 4), out.getvalue()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: wrote {os.path.getsize(path)} bytes; n_iter={arrays.get('n_iter')}")


def breast_cancer():
    from sklearn.datasets import load_breast_cancer
    from sklearn.preprocessing import StandardScaler
    X, y = load_breast_cancer(return_X_y=True)
    return StandardScaler().fit_transform(X), y


def diabetes():
    from sklearn.datasets import load_diabetes
    from sklearn.preprocessing import StandardScaler
    X, z = load_diabetes(return_X_y=True)
    return StandardScaler().fit_transform(X), z


def pack(s, tr, est, Xq, keep_trace=64, **more):
    out = dict(n_iter=np.int64(tr.shape[0]), alpha=s.alpha.copy(), g=s.neg_y_grad.copy(),
               trace_head=tr[:keep_trace], dec=np.asarray(est.decision_function(Xq)), **more)
    return out


# ---------------------------------------------------------------------------- fixtures


def g1():
    X, y = breast_cancer()
    est = BiKernelSVC(max_iter=10 ** 6)
    s, tr, _ = run(est, X, y)
    save("g1_breast_cancer_csvc", **pack(s, tr, est, X, rho=s.calculate_rho(), pred=est.predict(X)))


def g4():
    X = np.random.default_rng(0).standard_normal((8, 2))
    y = np.array([1, 0, 1, 0, 1, 1, 0, 0])
    est = BiKernelSVC(gamma=0.5)
    s, tr, _ = run(est, X, y)
    save("g4_tiny_trace", **pack(s, tr, est, X, keep_trace=10 ** 6, rho=s.calculate_rho(), X=X, y=y))


def g5():
    np.random.seed(0)
    X = 0.3 * np.random.randn(100, 2)
    Xtr = np.r_[X + 2, X - 2]
    est = OneClassSVM(nu=.1, gamma=.1, max_iter=10 ** 5)
    s, tr, _ = run(est, Xtr)
    save("g5_oneclass", **pack(s, tr, est, Xtr, rho=s.calculate_rho(), X=Xtr, pred=est.predict(Xtr)))


def g6():
    X, y = breast_cancer()
    for k in ("linear", "poly", "sigmoid"):
        est = BiKernelSVC(kernel=k, max_iter=10 ** 6)
        s, tr, _ = run(est, X, y)
        save(f"g6_breast_cancer_{k}", **pack(s, tr, est, X, rho=s.calculate_rho(), pred=est.predict(X)))
    np.random.seed(7)
    est = BiKernelSVC(rff=True, D=1000, max_iter=10 ** 6)
    s, tr, _ = run(est, X, y)
    save("g6_breast_cancer_rff", **pack(s, tr, est, X, rho=s.calculate_rho(), pred=est.predict(X)))


def g7():
    X, y = breast_cancer()
    est = BiNuSVC(max_iter=1000)
    s, tr, msg = run(est, X, y)
    rho, b = s.calculate_rho_b()
    save("g7_breast_cancer_nusvc", **pack(s, tr, est, X, rho=rho, b=b, pred=est.predict(X),
                                         printed=np.array(msg.strip())))


def g8():
    X, z = diabetes()
    for tag, est in (("svr", KernelSVR(max_iter=10 ** 6)),
                     ("svr_eps10_C100", KernelSVR(eps=10, C=100, max_iter=10 ** 6))):
        s, tr, _ = run(est, X, z)
        save(f"g8_diabetes_{tag}", **pack(s, tr, est, X, rho=s.calculate_rho()))
    est = NuSVR(max_iter=1000)
    s, tr, _ = run(est, X, z)
    rho, b = s.calculate_rho_b()
    save("g8_diabetes_nusvr", **pack(s, tr, est, X, rho=rho, b=b))


def g3():
    """digits one-vs-one: 45 sub-problems, gamma per sub-matrix."""
    from sklearn.datasets import load_digits
    from sklearn.preprocessing import StandardScaler
    X, y = load_digits(return_X_y=True)
    X = StandardScaler().fit_transform(X)
    TRACE.clear()
    est = KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(X, y)
    subs = est.multiclass_model.estimators_
    n_iters, nsv, rhos = [], [], []
    for e in subs:
        s = solver_of(e)
        nsv.append(int((s.alpha > 0).sum()))
        rhos.append(s.calculate_rho())
    # split the global trace per sub-problem: each sub-fit starts from alpha = 0, so replay sizes
    # come from counting updates between fits -- simpler: refit each pair alone.
    classes = np.unique(y)
    k = 0
    for a in range(len(classes)):
        for b in range(a + 1, len(classes)):
            m = (y == classes[a]) | (y == classes[b])
            yy = (y[m] == classes[b]).astype(int)       # sklearn _fit_ovo_binary: y==j -> 1
            e = BiKernelSVC(max_iter=10 ** 6)
            s, tr, _ = run(e, X[m], yy)
            assert int((s.alpha > 0).sum()) == nsv[k]
            n_iters.append(tr.shape[0])
            k += 1
    save("g3_digits_ovo", n_iter=np.int64(sum(n_iters)), sub_iters=np.array(n_iters),
         sub_nsv=np.array(nsv), sub_rho=np.array(rhos), pred=est.predict(X),
         dec_head=est.decision_function(X[:32]))


def s3():
    """C3-shaped synthetic classification, float32: converged at n=2000, capped at n=20000."""
    X, y = O.synth_classification(n=2000, d=128, seed=0)
    est = BiKernelSVC(max_iter=10 ** 7)
    s, tr, _ = run(est, X, y)
    save("s3_synth_cls_n2000", **pack(s, tr, est, X[:256], keep_trace=2000, rho=s.calculate_rho()))
    X, y = O.synth_classification(n=20000, d=128, seed=0)
    est = BiKernelSVC(max_iter=1000)
    s, tr, _ = run(est, X, y)
    save("s3_synth_cls_n20000_cap1000", **pack(s, tr, est, X[:256], keep_trace=1000, rho=s.calculate_rho()))


def s4():
    X, z = O.synth_regression(n=1000, d=64, seed=0)
    est = KernelSVR(max_iter=10 ** 7)
    s, tr, _ = run(est, X, z)
    save("s4_synth_svr_n1000", **pack(s, tr, est, X[:256], keep_trace=2000, rho=s.calculate_rho()))
    est = NuSVR(max_iter=2000)
    s, tr, _ = run(est, X, z)
    rho, b = s.calculate_rho_b()
    save("s4_synth_nusvr_n1000", **pack(s, tr, est, X[:256], keep_trace=2000, rho=rho, b=b))


def s5():
    X = O.synth_oneclass(n=2000, d=32, seed=1)
    est = OneClassSVM(max_iter=10 ** 7)
    s, tr, _ = run(est, X)
    save("s5_synth_oneclass_n2000", **pack(s, tr, est, X[:256], keep_trace=2000, rho=s.calculate_rho()))


def pack_big(s, tr, est, Xq, stride=16, **more):
    """Full-size fixtures: alpha is sparse after a capped run (indices + values), the gradient is kept at the
    variables the run touched plus every `stride`-th variable, so the file stays small."""
    a = s.alpha
    nz = np.flatnonzero(a)
    touched = np.unique(tr[:, :2].astype(np.int64).reshape(-1))
    gidx = np.unique(np.r_[touched, np.arange(0, a.shape[0], stride)])
    return dict(n_iter=np.int64(tr.shape[0]), n_vars=np.int64(a.shape[0]), alpha_idx=nz, alpha_val=a[nz],
                g_idx=gidx, g_val=s.neg_y_grad[gidx], trace=tr, dec=np.asarray(est.decision_function(Xq)), **more)


def s3full():
    """BASELINE configs[2] at its real size (n=200000, d=128, float32), capped at 1000 iterations."""
    X, y = O.synth_classification(n=200_000, d=128, seed=0)
    est = BiKernelSVC(max_iter=1000)
    s, tr, _ = run(est, X, y)
    save("s3_full_n200000_cap1000", **pack_big(s, tr, est, X[:256], rho=s.calculate_rho()))


def s4full():
    """BASELINE configs[3] at its real size (l=100000, d=64 -> 200000 variables), KernelSVR capped at 200 iterations."""
    X, z = O.synth_regression(n=100_000, d=64, seed=0)
    est = KernelSVR(max_iter=200)
    s, tr, _ = run(est, X, z)
    save("s4_full_svr_l100000_cap200", **pack_big(s, tr, est, X[:256], rho=s.calculate_rho()))


def s5big():
    """configs[4]-shaped one-class problem at n=10000 (the largest the O(l^2) reference start finishes quickly), converged."""
    X = O.synth_oneclass(n=10_000, d=32, seed=1)
    est = OneClassSVM(max_iter=10 ** 7)
    s, tr, _ = run(est, X)
    save("s5_synth_oneclass_n10000", n_iter=np.int64(tr.shape[0]), alpha=s.alpha.copy(), g=s.neg_y_grad.copy(),
         trace_head=tr[:4000], dec=np.asarray(est.decision_function(X[:512])), pred=est.predict(X[:512]),
         rho=s.calculate_rho())


def lin():
    X, y = breast_cancer()
    est = BiLinearSVC(max_iter=10 ** 6)
    TRACE.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        est.fit(X, y)
    tr = np.array(TRACE).reshape(-1, 4)
    save("lin_breast_cancer_svc", n_iter=np.int64(tr.shape[0]), w=est.coef_[0], rho=est.coef_[1],
         dec=est.decision_function(X), trace_head=tr[:64])
    X, z = diabetes()
    est = LinearSVR(max_iter=10 ** 6)
    TRACE.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        est.fit(X, z)
    tr = np.array(TRACE).reshape(-1, 4)
    save("lin_diabetes_svr", n_iter=np.int64(tr.shape[0]), w=est.coef_[0], rho=est.coef_[1],
         dec=est.decision_function(X), trace_head=tr[:64])


def rffvec():
    """NormalRFF draw order + transform values (rff.py:12-20)."""
    from pysvm.rff import NormalRFF
    np.random.seed(3)
    X = np.random.default_rng(5).standard_normal((16, 6))
    r = NormalRFF(gamma=0.3, D=32).fit(X)
    save("rff_vectors", X=X, w=r.w, b=r.b, Z=r.transform(X))


ALL = dict(g1=g1, g3=g3, g4=g4, g5=g5, g6=g6, g7=g7, g8=g8, s3=s3, s4=s4, s5=s5, lin=lin, rffvec=rffvec,
           s3full=s3full, s4full=s4full, s5big=s5big)

if __name__ == "__main__":
    for name in (sys.argv[1:] or ALL):
        ALL[name]()
