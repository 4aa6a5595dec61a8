"""Which rows deviate?  Emulated ranks vs single GPU from the same state, in chunks; on the first deviation print the
rows whose gradient differs (one row / a tile / a whole CTA / a whole rank tells where the value came from)."""
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
CH = int(os.environ.get("CHUNK", 400))
sol = S.SolverWithCache(p, y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
sol.solve(600)
found = 0
for chunk in range(12):
    a0, g0 = sol.alpha.cpu().numpy().copy(), sol.neg_y_grad.cpu().numpy().copy()
    sol.solve(CH)
    a1, g1 = sol.alpha.cpu().numpy().copy(), sol.neg_y_grad.cpu().numpy().copy()
    for rep in range(4):
        r = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
        r.set_state(a0, g0)
        st = r.solve(CH)
        g = r.gather("neg_y_grad")
        bad = np.flatnonzero(g != g1)
        if len(bad):
            found += 1
            ncta = st[0].n_ctas
            rpr = r.shards[0].rows_per_rank
            print(f"chunk {chunk} rep {rep}: {len(bad)} rows deviate, max |dG| {np.abs(g - g1).max():.3e}; alpha rows differ: {np.flatnonzero(r.gather('alpha') != a1).tolist()[:10]}")
            print("   rows:", bad[:40].tolist(), "..." if len(bad) > 40 else "")
            print("   ranks:", sorted(set((bad // rpr).tolist())), " tiles (local):", sorted(set((((bad % rpr)) // 32).tolist()))[:40], " ctas/rank", ncta, flush=True)
            vals = np.abs(g - g1)[bad]
            print("   |dG| distinct values:", np.unique(np.round(vals / vals.max(), 3))[:10], flush=True)
    if found >= 3:
        break
print("done, deviations found:", found)
