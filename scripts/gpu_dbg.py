import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import smo_oracle as O
from pysvm_b200 import solver as S
for n, d in [(8, 2), (500, 30), (4096, 32), (200000, 128)]:
    X, y01 = O.synth_classification(n=n, d=d, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    sol = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 1.0 / d), cache_bytes=0))
    print(n, d, "select ->", sol.working_set_select())
    st = sol.solve(50)
    print("   solve(50):", {k: getattr(st, k) for k in ("n_iter", "converged", "n_ctas", "cold_sweeps")})
