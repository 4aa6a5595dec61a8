import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smo_oracle as O
from pysvm_b200 import solver as S, sharded as SH
X, y01 = O.synth_classification(n=6000, d=128, seed=1)
y = np.where(y01 == 1, 1.0, -1.0)
p = -np.ones(len(y))
spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
def emu(world, iters):
    r = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
    st = r.solve(iters)
    return r.gather("alpha")
def single(iters):
    sol = S.SolverWithCache(p, y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
    sol.solve(iters)
    return sol.alpha.cpu().numpy()
ref = single(4000)
outs = [emu(2, 4000) for _ in range(5)]
print(os.environ.get("TAG"), "emu2 vs single:", [float(np.abs(o - ref).max()) for o in outs], "single again:", float(np.abs(single(4000) - ref).max()), flush=True)
outs = [emu(3, 4000) for _ in range(3)]
print(os.environ.get("TAG"), "emu3 vs single:", [float(np.abs(o - ref).max()) for o in outs], flush=True)
