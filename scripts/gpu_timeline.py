"""Per-phase timeline of the persistent kernel's leader (CTA 0), B2SVM_TIMELINE=1.
slots (type warp 0 of CTA 0): 0 sweep done | 1 CTA warps joined | 7 CTA winner | 2 level-1 record published |
       3 GPU winner known (level-1 poll) | 4 box winner known (level-2, world > 1) | 8 winner's row / alpha on hand |
       9 type warps met | 10 K_ij | 11 step | 5 leader done | 6 sweep starts"""
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
    pub, seen = ct[:, 0].astype(np.int64), ct[:, 1].astype(np.int64)
    t0 = pub.min()
    print(f"   epoch 41, global ns timer: publish times spread {pub.max() - t0} ns (sorted, first/last 4: {np.sort(pub - t0)[:4].tolist()} .. {np.sort(pub - t0)[-4:].tolist()})")
    print(f"   all records seen: {np.sort(seen - t0)[:4].tolist()} .. {np.sort(seen - t0)[-4:].tolist()} ns after the first publish;"
          f" last publish -> first 'all seen' {seen.min() - pub.max()} ns, -> last 'all seen' {seen.max() - pub.max()} ns; CTA 0 published at {pub[0] - t0}, saw all at {seen[0] - t0}")
    k = min(iters, 60)
    seg = t[2:k]      # skip first epochs
    chain = [("join (0->1)", 0, 1), ("CTA winner (1->7)", 1, 7), ("publish L1 (7->2)", 7, 2), ("alpha park + L1 poll (2->3)", 2, 3),
             ("level 2 (3->4)", 3, 4), ("row / alpha fetch (4->8)", 4, 8), ("kself + type barrier (8->9)", 8, 9),
             ("select + K_ij (9->10)", 9, 10), ("smo_step (10->11)", 10, 11), ("cache plan + hand-over (11->5)", 11, 5),
             ("bar (5->6)", 5, 6)]
    us = st.device_ms * 1e3 / st.n_iter
    print(f"n={n} d={d} {np.dtype(dtype).name}: {us:.2f} us/it, ctas {st.n_ctas}; cycles per phase (median):")
    tot = 0
    for nm, a, b in chain:
        v = np.median(seg[:, b] - seg[:, a])
        print(f"   {nm:34s} {v:9.0f}")
        tot += v
    print(f"      poll detail: publish->poll entry {np.median(seg[:, 12] - seg[:, 2]):.0f}, first pass {np.median(seg[:, 13] - seg[:, 12]):.0f}, "
          f"loop total {np.median(seg[:, 15] - seg[:, 12]):.0f}, passes median {np.median(seg[:, 14]):.0f} max {seg[:, 14].max()}, reduce after {np.median(seg[:, 3] - seg[:, 15]):.0f}")
    sw = np.median(seg[1:, 0] - seg[:-1, 6])
    print(f"   {'sweep (6->next 0)':34s} {sw:9.0f}")
    print(f"   chain {tot:.0f} + sweep {sw:.0f} = {tot + sw:.0f} cycles -> {(tot + sw)/us:.0f} cycles/us")
    if os.environ.get("PER_EPOCH"):
        # a warm state mixes cached-column and recomputing sweeps: list them one by one
        print("   epoch: chain(0->6) sweep(6->next 0) | join L1poll fetch select+K step plan")
        for e in range(2, k - 1):
            r = t[e]
            print(f"     {e:3d}: {r[6]-r[0]:6d} {t[e+1][0]-r[6]:6d} | {r[1]-r[0]:5d} {r[3]-r[2]:5d} {r[8]-r[4]:5d} {r[10]-r[9]:5d} {r[11]-r[10]:5d} {r[5]-r[11]:5d}")
    print(f"   warp 0, first tile of the sweep at epoch 10 (cycles from sweep start): tile loop entered {wt[12, 3]}, dot products done {wt[12, 0]}, "
          f"transposed reduce done {wt[12, 1]}, row epilogue done {wt[12, 2]}")
    if st.cold_sweeps == 0:
        b6 = t[9, 6]
        print(f"   warm sweep of warp 0 at epoch 10 (cycles from sweep start {b6}): path entered {wt[12,3]-b6}, loads issued {wt[12,0]-b6}, first tile done {wt[12,1]-b6}, all done {wt[12,2]-b6}, next epoch stamp 0 at {t[10,0]-b6}")
    base = wt[:11, 0].min()
    print("   per-warp sweep at epoch 10 (start, end rel. cycles; tiles; wait cycles):")
    for w in range(11):
        print(f"     warp {w:2d}: {wt[w,0]-base:7d} {wt[w,1]-base:7d}  tiles {wt[w,2]:3d}  wait {wt[w,3]:7d}")

import sys
for a in sys.argv[1:]:
    n, d, dt = a.split(':')
    run(int(n), int(d), 60, np.float64 if dt == 'f64' else np.float32)
