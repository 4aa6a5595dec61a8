"""Host-side cost of creating a handle with a large column cache (cudaMalloc of up to 0.7 x free memory)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysvm_b200 import synth as G, solver as S
n, d = 200000, 128
X, y01 = G.classification(n=n, d=d, seed=0)
X = X.astype(np.float32)
y = np.where(y01 == 1, 1.0, -1.0)
spec = S.KernelSpec("rbf", 1 / (d * X.std()))
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for gb in (None, 120, 60, 20, 4, 120, None):
    t0 = time.perf_counter()
    func = S.QColumns(X, y, spec, cache_bytes=None if gb is None else gb << 30)
    t1 = time.perf_counter()
    sol = S.SolverWithCache(-np.ones(n), y, 1.0, 1e-5, func=func, max_iter_hint=10 ** 6)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = sol.solve(10)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    del sol, func
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"cache {gb} GB: QColumns (upload) {1e3*(t1-t0):.1f} ms, solver create {1e3*(t2-t1):.1f} ms, 10 iterations {1e3*(t3-t2):.1f} ms, destroy {1e3*(t4-t3):.1f} ms", flush=True)
