"""Short C3-shaped run for an ncu capture of the persistent kernel (launch 0 = warm-up, 1 = measured)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S
n = int(os.environ.get("N", 200000)); d = int(os.environ.get("D", 128)); iters = int(os.environ.get("ITERS", 200))
Xs, ys = G.classification(n=n, d=d, seed=0)
yv = np.where(ys == 1, 1.0, -1.0)
sol = S.SolverWithCache(-np.ones(n), yv, 1.0, func=S.QColumns(Xs, yv, S.KernelSpec("rbf", 1 / (d * Xs.std())), cache_bytes=0 if os.environ.get("CACHE", "off") == "off" else None))
sol.solve(20)
st = sol.solve(iters)
print(f"n={n} d={d}: {st.n_iter} it, {st.device_ms*1e3/st.n_iter:.2f} us/it, ctas {st.n_ctas}")
