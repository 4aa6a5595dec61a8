"""Single-step variant: emulated ranks and the single-GPU solver advance K iterations per launch; the gradient is
compared after every launch; on a deviation print which rows differ and re-synchronise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smo_oracle as O
from pysvm_b200 import solver as S, sharded as SH

X, y01 = O.synth_classification(n=6000, d=128, seed=1)
y = np.where(y01 == 1, 1.0, -1.0)
p = -np.ones(len(y))
spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
world = int(os.environ.get("WORLD", 2))
K = int(os.environ.get("K", 1))
sol = S.SolverWithCache(p, y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
sol.solve(700)
r = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
r.set_state(sol.alpha.cpu().numpy(), sol.neg_y_grad.cpu().numpy())
rpr = r.shards[0].rows_per_rank
found = 0
for step in range(int(os.environ.get("STEPS", 4000))):
    pair = sol.working_set_select() if K == 1 else None
    s1 = sol.solve(K)
    st = r.solve(K)
    g1 = sol.neg_y_grad.cpu().numpy()
    g = r.gather("neg_y_grad")
    bad = np.flatnonzero(g != g1)
    abad = np.flatnonzero(r.gather("alpha") != sol.alpha.cpu().numpy())
    if len(bad) or len(abad):
        found += 1
        ncta = st[0].n_ctas
        print(f"step {step} (pair {pair}): {len(bad)} rows deviate in G, max {np.abs(g - g1).max():.3e}; alpha differs at {abad.tolist()[:6]}")
        print("   rows:", bad[:24].tolist(), "..." if len(bad) > 24 else "", " ranks:", sorted(set((bad // rpr).tolist())),
              " local tiles:", sorted(set(((bad % rpr) // 32).tolist()))[:24], " ctas/rank", ncta, flush=True)
        r.set_state(sol.alpha.cpu().numpy(), g1)
        if found >= 4:
            break
print("done; deviations:", found)
