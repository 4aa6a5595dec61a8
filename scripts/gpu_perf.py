"""Quick throughput numbers of the fused solve (no instrumentation)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S
def run(n, d, iters, cache):
    X, y01 = G.classification(n=n, d=d, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    sol = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 1 / (d * X.std())), cache_bytes=cache))
    sol.solve(50)
    sol._engine.init_state(sol.p)
    st = sol.solve(iters)
    print(f"n={n} d={d} cache={'on' if cache is None or cache else 'off'}: {st.n_iter} it, {st.device_ms*1e3/st.n_iter:.2f} us/it, "
          f"{st.algorithmic_bytes/st.device_ms/1e6:.0f} GB/s alg, ctas {st.n_ctas}, cold {st.cold_sweeps} warm {st.cache_hits//2} conv {st.converged}", flush=True)
for a in sys.argv[1:]:
    n, d, it, c = a.split(":")
    run(int(n), int(d), int(it), 0 if c == "off" else None)
