"""Where does the wall time of BiKernelSVC(max_iter=20000).fit(X_host, y_host) go?"""
import os, sys, time, contextlib, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S
from pysvm_b200.svm.svc import BiKernelSVC
X, y01 = G.classification(n=200000, d=128, seed=0)
Xp = torch.from_numpy(X).pin_memory().numpy()
def T(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"  {label:34s} {1e3*(time.perf_counter()-t):8.2f} ms", flush=True); return r
for rep in range(2):
    print("rep", rep)
    with contextlib.redirect_stdout(io.StringIO()):
        T("whole fit(max_iter=20000)", lambda: BiKernelSVC(max_iter=20000).fit(Xp, y01))
    std = T("X.std() numpy", lambda: Xp.std())
    y = T("labels", lambda: np.where(y01 == 1, 1.0, -1.0))
    Xd = T("H2D X (pinned)", lambda: S.to_device_matrix(Xp, torch.device("cuda", 0)))
    func = T("QColumns", lambda: S.QColumns(Xd, y, S.KernelSpec("rbf", 1 / (128 * std))))
    sol = T("SolverWithCache (create handle)", lambda: S.SolverWithCache(-np.ones(len(y)), y, 1.0, func=func, max_iter_hint=20000))
    st = T("solve 20000", lambda: sol.solve(20000))
    print("     kernel ms", st.device_ms, "cold", st.cold_sweeps, "warm", st.cache_hits // 2)
    T("rho", lambda: sol.calculate_rho())
    T("alpha D2H", lambda: sol.alpha.cpu().numpy())
    T("destroy", lambda: sol._engine.close())
