"""First-light diagnostics on a B200: small parity cases first, then the C3 shape with timing.
Prints progressively so a hang shows how far it got."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t00 = time.time()
def log(*a):
    print(f"[{time.time()-t00:7.2f}s]", *a, flush=True)

import torch
log("torch", torch.__version__, torch.cuda.get_device_name(0))
from oracle import smo_oracle as O
from pysvm_b200 import solver as S
from pysvm_b200.svm import svc

g = np.load("tests/golden/g4_tiny_trace.npz")
X, y = g["X"], np.where(g["y"] == 1, 1.0, -1.0)
func = S.QColumns(X, y, S.KernelSpec("rbf", 0.5))
sol = S.SolverWithCache(-np.ones(8), y, 1.0, 1e-5, func=func)
log("bound; G0 =", sol.neg_y_grad.cpu().numpy())
for k in range(3):
    i, j = sol.working_set_select()
    d = sol.update(i, j)
    log("step", k, (i, j), d, "want", g["trace_head"][k])
st = sol.solve(100)
log("g4 fused: n_iter", st.n_iter, "converged", st.converged, "alpha", sol.alpha.cpu().numpy(), "want", g["alpha"])

from sklearn.datasets import load_breast_cancer
from sklearn.preprocessing import StandardScaler
Xb, yb = load_breast_cancer(return_X_y=True)
Xb = StandardScaler().fit_transform(Xb)
t = time.time()
est = svc.BiKernelSVC(max_iter=10**6).fit(Xb, yb)
log("breast_cancer: n_iter", est.n_iter_, "(ref 487) nSV", len(est.support_), "(ref 119) rho", est.rho_, "(ref 0.235368318162)",
    "fit s", time.time() - t, est.solve_stats_)

for n, iters in ((20000, 1000), (200000, 2000)):
    Xs, ys = O.synth_classification(n=n, d=128, seed=0)
    yv = np.where(ys == 1, 1.0, -1.0)
    gamma = 1 / (128 * Xs.std())
    func = S.QColumns(Xs, yv, S.KernelSpec("rbf", gamma))
    sol = S.SolverWithCache(-np.ones(n), yv, 1.0, func=func)
    st = sol.solve(10)
    log(n, "warm", st.as_dict())
    for rep in range(3):
        st = sol.solve(iters)
        us = st.device_ms * 1e3 / max(1, st.n_iter)
        gbs = st.algorithmic_bytes / (st.device_ms * 1e-3) / 1e9
        log(f"n={n}: {st.n_iter} it in {st.device_ms:.2f} ms -> {us:.2f} us/it, {1e6/us:.0f} it/s, {gbs:.0f} GB/s algorithmic, ctas {st.n_ctas}")
    if n == 20000:
        gold = np.load("tests/golden/s3_synth_cls_n20000_cap1000.npz")
        sol2 = S.SolverWithCache(-np.ones(n), yv, 1.0, func=S.QColumns(Xs, yv, S.KernelSpec("rbf", gamma)))
        sol2.solve(1000)
        log("n=20000 cap1000 max|alpha-ref|", float(np.abs(sol2.alpha.cpu().numpy() - gold["alpha"]).max()),
            "max|g-ref|", float(np.abs(sol2.neg_y_grad.cpu().numpy() - gold["g"]).max()))
log("done")
