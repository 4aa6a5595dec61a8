"""Converged C3 fit (n=200000, d=128) timing: device seconds, warm / cold sweeps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysvm_b200 import synth as G
from pysvm_b200.svm import svc
n = int(os.environ.get("N", 200000))
X, y = G.classification(n=n, d=128, seed=0)
X = X.astype(np.float32)
for rep in range(int(os.environ.get("REPS", 2))):
    t0 = time.perf_counter()
    est = svc.BiKernelSVC(max_iter=10 ** 6).fit(X, y)
    torch.cuda.synchronize()
    st = est.solve_stats_
    print(f"fit {time.perf_counter() - t0:.3f} s wall, device {st['device_ms'] / 1e3:.3f} s, n_iter {st['n_iter']}, cold {st['cold_sweeps']}, warm {st['n_iter'] - st['cold_sweeps']}", flush=True)
