"""Gradient consistency at the C3 shape: after K fused iterations, compare the incrementally updated
neg_y_grad with a from-scratch recomputation.  Usage: gpu_consistency.py iters [iters ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import smo_oracle as O
from pysvm_b200 import solver as S
n, d = int(os.environ.get("N", 200000)), 128
X, y01 = O.synth_classification(n=n, d=d, seed=0)
y = np.where(y01 == 1, 1.0, -1.0)
gamma = 1 / (d * X.std())
for iters in [int(a) for a in sys.argv[1:]]:
    sol = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", gamma)))
    st = sol.solve(iters)
    a = sol.alpha.cpu().numpy()
    g_inc = sol.neg_y_grad.clone()
    sol._engine.init_gradient(sol.p)
    diff = (g_inc - sol.neg_y_grad).abs()
    bad = (diff > 1e-3).nonzero().flatten().cpu().numpy()
    print(f"flags={os.environ.get('B2SVM_DEBUG_FLAGS','0')} iters={iters}: {st.device_ms*1e3/max(1,st.n_iter):.2f} us/it  sum(y*a)={float(a@y):.2e}  max|dG|={float(diff.max()):.3e}  bad rows={len(bad)} {bad[:12]}", flush=True)
    if len(bad):
        ntiles = (n + 31) // 32
        ncta = st.n_ctas
        tb, tr = divmod(ntiles, ncta)
        tile = bad // 32
        starts = np.array([c * tb + min(c, tr) for c in range(ncta + 1)])
        cta = np.searchsorted(starts, tile, side="right") - 1
        tin = tile - starts[cta]
        ut, cnt = np.unique(tile, return_counts=True)
        print("   bad tiles", len(ut), "rows per bad tile: min %d max %d" % (cnt.min(), cnt.max()))
        h = np.bincount(tin, minlength=44)
        print("   bad rows by tile-index-within-CTA:", h.tolist())
        hc = np.bincount(cta, minlength=ncta)
        print("   bad rows by CTA (first 40):", hc[:40].tolist())
        d = diff.cpu().numpy()
        t0 = ut[0]
        print("   first bad tile", t0, "diffs:", np.round(d[t0*32:(t0+1)*32], 4).tolist())
