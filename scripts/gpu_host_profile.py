"""Host-side time of (a) the digits one-vs-one fit and (b) the converged C3 fit + decision_function (cProfile)."""
import cProfile, io, os, pstats, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysvm_b200 as P
from pysvm_b200 import synth
from sklearn.datasets import load_digits
from sklearn.preprocessing import StandardScaler


def quiet(f):
    with contextlib.redirect_stdout(io.StringIO()):
        return f()


def prof(f, n=25):
    pr = cProfile.Profile(); pr.enable(); r = quiet(f); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(n); print(s.getvalue()[:6000]); return r


Xd, yd = load_digits(return_X_y=True); Xd = StandardScaler().fit_transform(Xd)
quiet(lambda: P.KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(Xd, yd))
for _ in range(2):
    t0 = time.perf_counter(); quiet(lambda: P.KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(Xd, yd)); torch.cuda.synchronize(); print("digits ovo fit", time.perf_counter() - t0, flush=True)
prof(lambda: P.KernelSVC(multiclass="ovo", max_iter=10 ** 6).fit(Xd, yd))
if not os.environ.get("SKIP_C3"):
    X, y01 = synth.classification()
    Xp = torch.from_numpy(X).pin_memory().numpy()
    t0 = time.perf_counter(); full = quiet(lambda: P.svm.svc.BiKernelSVC(max_iter=10 ** 7).fit(Xp, y01)); torch.cuda.synchronize(); print("C3 converged fit", time.perf_counter() - t0, "device", full.solve_stats_["device_ms"] * 1e-3, flush=True)
    full.decision_function(Xp[:4096])
    for _ in range(2):
        t0 = time.perf_counter(); d = full.decision_function(Xp); torch.cuda.synchronize(); print("decision all rows", time.perf_counter() - t0, flush=True)
    del full
    full = prof(lambda: P.svm.svc.BiKernelSVC(max_iter=10 ** 7).fit(Xp, y01), 18)
    prof(lambda: full.decision_function(Xp), 12)
