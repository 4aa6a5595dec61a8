"""Per-phase timeline of the persistent kernel's leader (CTA 0), B2SVM_TIMELINE=1.
slots: 0 sweep done | 1 CTA warps joined | 2 record published | 3 all records read | 4 winners known
       5 step taken | 6 sweep starts"""
import os, sys, ctypes as C
os.environ["B2SVM_TIMELINE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S, _native as N

def run(n, d, iters, dtype=np.float32):
    Xs, ys = G.classification(n=n, d=d, seed=0)
    Xs = Xs.astype(dtype)
    yv = np.where(ys == 1, 1.0, -1.0)
    func = S.QColumns(Xs, yv, S.KernelSpec("rbf", 1 / (d * Xs.std())), cache_bytes=int(os.environ.get("CACHE_BYTES", "0")))
    sol = S.SolverWithCache(-np.ones(n), yv, 1.0, func=func)
    sol.solve(int(os.environ.get("FIRST_ITERS", iters)))   # e.g. enough iterations to fill the column cache
    st = sol.solve(iters)
    print(f"   timed solve: {st.n_iter} it, cold sweeps {st.cold_sweeps}, warm {st.cache_hits // 2}")
    lib = N.load()
    NT_ALL = 64 * 16 + 64 + 320
    buf = (C.c_longlong * NT_ALL)()
    lib.b2svm_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    N.check(lib.b2svm_debug_timeline(sol._engine.h, buf, NT_ALL))
    allb = np.array(buf[:], dtype=np.int64)
    t = allb[:1024].reshape(64, 16)
    wt = allb[1024:1088].reshape(16, 4)
    ct = allb[1088:].reshape(160, 2)[:st.n_ctas]
    for e in (0, 1):
        c = ct[:, e]
        print('   per-CTA sweep cycles (epoch %d): min %d median %d p90 %d max %d' % (10 + 30 * e, c.min(), np.median(c), np.percentile(c, 90), c.max()))
    order = np.argsort(ct[:, 0])
    print('   slowest CTAs at epoch 10:', [(int(c), int(ct[c, 0]), int(ct[c, 1])) for c in order[-10:]])
    print('   fastest CTAs at epoch 10:', [(int(c), int(ct[c, 0]), int(ct[c, 1])) for c in order[:6]])
    print('   correlation of per-CTA sweep time between the two epochs: %.3f' % np.corrcoef(ct[:, 0], ct[:, 1])[0, 1])
    print('   sweep time by CTA index block of 16 (epoch 10 / 40):', [(int(ct[b:b+16, 0].mean()), int(ct[b:b+16, 1].mean())) for b in range(0, st.n_ctas, 16)])
    k = min(iters, 60)
    seg = t[2:k]      # skip first epochs
    names = ["join(0->1)", "publish(1->2)", "poll+read(2->3)", "reduce(3->4)", "step(4->5)", "bar(5->6)", "sweep(6->next0)"]
    d01 = [seg[:, 1] - seg[:, 0], seg[:, 2] - seg[:, 1], seg[:, 3] - seg[:, 2], seg[:, 4] - seg[:, 3],
           seg[:, 5] - seg[:, 4], seg[:, 6] - seg[:, 5], np.r_[seg[1:, 0] - seg[:-1, 6], 0]]
    us = st.device_ms * 1e3 / st.n_iter
    print(f"n={n} d={d} {np.dtype(dtype).name}: {us:.2f} us/it, ctas {st.n_ctas}; cycles per phase (median):")
    tot = 0
    for nm, v in zip(names, d01):
        print(f"   {nm:18s} {np.median(v):9.0f}")
        tot += np.median(v)
    print(f"   total {tot:.0f} cycles -> {tot/us:.0f} cycles/us")
    fine = [("join->CTA winners (1->7)", seg[:, 7] - seg[:, 1]), ("push record (7->2)", seg[:, 2] - seg[:, 7]), ("decide + issue loads (4->8)", seg[:, 8] - seg[:, 4]),
            ("rows + pair kernels (8->9)", seg[:, 9] - seg[:, 8]), ("alpha settled (9->10)", seg[:, 10] - seg[:, 9]), ("smo_step (10->11)", seg[:, 11] - seg[:, 10]), ("cache plan + publish (11->5)", seg[:, 5] - seg[:, 11])]
    for nm, v in fine:
        print(f"      {nm:30s} {np.median(v):9.0f}")
    base = wt[:11, 0].min()
    print("   per-warp sweep at epoch 10 (start, end rel. cycles; tiles; wait cycles):")
    for w in range(11):
        print(f"     warp {w:2d}: {wt[w,0]-base:7d} {wt[w,1]-base:7d}  tiles {wt[w,2]:3d}  wait {wt[w,3]:7d}")

import sys
for a in sys.argv[1:]:
    n, d, dt = a.split(':')
    run(int(n), int(d), 60, np.float64 if dt == 'f64' else np.float32)
