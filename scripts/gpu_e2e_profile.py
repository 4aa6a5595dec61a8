"""Where the host time of the public estimator call goes: cProfile + wall-clock split of
BiKernelSVC(max_iter=ITERS).fit(X_host, y_host) on the bench workload."""
import cProfile, io, os, pstats, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pysvm_b200 import synth
from pysvm_b200.svm.svc import BiKernelSVC

ITERS = int(os.environ.get("ITERS", 20000))
X, y01 = synth.classification(n=200_000, d=128, seed=0)
Xp = torch.from_numpy(X).pin_memory().numpy()


def fit():
    with contextlib.redirect_stdout(io.StringIO()):
        return BiKernelSVC(max_iter=ITERS).fit(Xp, y01)


est = fit()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    est = fit()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"fit {rep}: wall {dt*1e3:.1f} ms, device {est.solve_stats_['device_ms']:.1f} ms, host overhead {dt*1e3 - est.solve_stats_['device_ms']:.1f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
est = fit()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
