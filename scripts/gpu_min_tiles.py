"""us/iteration (cold path) against the CTA plan: B2SVM_MIN_TILES_PER_CTA for slices the size of one GPU's share."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysvm_b200 import synth as G, solver as S

for n, d in ((25000, 128), (12500, 128), (50000, 128), (25000, 32), (4096, 128)):
    X, y01 = G.classification(n=n, d=d, seed=0)
    X = X.astype(np.float32)
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (d * X.std()))
    line = []
    for mt in (11, 8, 5, 3, 2):
        os.environ["B2SVM_MIN_TILES_PER_CTA"] = str(mt)
        best = 1e9
        for rep in range(3):
            sol = S.SolverWithCache(-np.ones(n), y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
            st = sol.solve(3000)
            best = min(best, st.device_ms * 1e3 / st.n_iter)
        line.append(f"min_tiles {mt}: {best:.2f} us ({st.n_ctas} CTAs)")
    print(f"n={n} d={d}: " + " | ".join(line), flush=True)
