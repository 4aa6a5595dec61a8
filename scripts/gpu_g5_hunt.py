"""Why does the g5 one-class fixture sometimes take 47 instead of 51 iterations when it runs after the SVR tests?
Replays [NuSVR fit on diabetes] -> [one-class fit, step by step] and prints the state after each stage."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn.datasets import load_diabetes
from sklearn.preprocessing import StandardScaler
from pysvm_b200 import solver as S
from pysvm_b200.svm import oc_svm, svr
from pysvm_b200.svm.oc_svm import oneclass_initial_alpha


def dg(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]


Xd, zd = load_diabetes(return_X_y=True)
Xd = StandardScaler().fit_transform(Xd)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "g5_oneclass.npz"))
X = g["X"]
l = len(X)
print("one-class X", X.shape, X.dtype)
K = np.exp(-0.1 * ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))
a0 = oneclass_initial_alpha(.1, l)
a0 = a0.cpu().numpy() if hasattr(a0, "cpu") else np.asarray(a0)
g_ref = -(K @ a0)

for rep in range(int(os.environ.get("REPS", 12))):
    pre = os.environ.get("PRE", "nusvr")
    if pre == "nusvr":
        svr.NuSVR(max_iter=1000).fit(Xd, zd)
    elif pre == "svr":
        svr.KernelSVR(max_iter=10 ** 6, eps=10, C=100).fit(Xd, zd)
    est = oc_svm.OneClassSVM(nu=.1, gamma=.1, max_iter=10 ** 5)
    est.n_features = X.shape[1]
    func, spec, rff = est._columns(X, np.ones(l))
    sol = S.SolverWithCache(np.zeros(l), np.ones(l), 1, est.tol, est.cache_size, func=func, max_iter_hint=est.max_iter)
    sol.alpha = oneclass_initial_alpha(.1, l)
    sol._engine.init_gradient(None)
    g0 = sol.neg_y_grad.cpu().numpy().copy()
    a_in = sol.alpha.cpu().numpy().copy()
    # iterate one step at a time through the fused path, logging the pairs
    trace = []
    if os.environ.get("STEPWISE"):
        for it in range(60):
            i, j = sol.working_set_select()
            if i < 0:
                break
            trace.append((i, j))
            sol.update(i, j)
        n_iter = len(trace)
    else:
        st = sol.solve(10 ** 5)
        n_iter = st.n_iter
    print(f"rep {rep}: n_iter {n_iter}  alpha_in {dg(a_in)}  g0 {dg(g0)} (max dev from numpy {np.abs(g0 - g_ref).max():.2e})  "
          f"alpha_out {dg(sol.alpha.cpu().numpy())}  X dtype on device {func.X.dtype} stride {func.X.stride()} ptr%512 {func.X.data_ptr() % 512}", flush=True)
    if trace:
        print("   pairs:", trace[:12])
