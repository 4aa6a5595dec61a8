"""Determinism / equivalence probe: single-GPU solve twice, emulated ranks (2, 4) twice; first diverging iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smo_oracle as O
from pysvm_b200 import solver as S, sharded as SH

X, y01 = O.synth_classification(n=6000, d=128, seed=1)
y = np.where(y01 == 1, 1.0, -1.0)
p = -np.ones(len(y))
spec = S.KernelSpec("rbf", 1 / (128 * X.std()))


def single(iters):
    sol = S.SolverWithCache(p, y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
    st = sol.solve(iters)
    return sol.alpha.cpu().numpy(), sol.neg_y_grad.cpu().numpy(), st


def emu(world, iters):
    r = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
    st = r.solve(iters)
    return r.gather("alpha"), r.gather("neg_y_grad"), st[0]


for iters in (50, 200, 1000, 3000, 10 ** 6):
    a1, g1, s1 = single(iters)
    a2, g2, s2 = single(iters)
    e1, h1, t1 = emu(2, iters)
    e2, h2, t2 = emu(2, iters)
    f1, k1, u1 = emu(4, iters)
    print(iters, "n_iter", s1.n_iter, s2.n_iter, t1.n_iter, t2.n_iter, u1.n_iter,
          "| single-single", np.abs(a1 - a2).max(), "emu2-emu2", np.abs(e1 - e2).max(), "single-emu2", np.abs(a1 - e1).max(),
          "single-emu4", np.abs(a1 - f1).max(), "G single-emu2", np.abs(g1 - h1).max(), "next pair", (s1.last_i, s1.last_j), (t1.last_i, t1.last_j), flush=True)
