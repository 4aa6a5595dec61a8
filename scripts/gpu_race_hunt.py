"""Race hunt (TIMELINE build + B2SVM_DEBUG_STEPS=1): every CTA of every emulated rank records the pair, gradients and
alphas its leader saw per epoch; all views of an epoch must be identical.  Prints the first epoch where they are not."""
import os, sys, ctypes as C
os.environ["B2SVM_TIMELINE"] = "1"
os.environ["B2SVM_DEBUG_STEPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smo_oracle as O
from pysvm_b200 import solver as S, sharded as SH, _native as N

X, y01 = O.synth_classification(n=6000, d=128, seed=1)
y = np.where(y01 == 1, 1.0, -1.0)
p = -np.ones(len(y))
spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
W = 16
WARM = int(os.environ.get("WARM", 1500))
EPOCHS = 2000
sol = S.SolverWithCache(p, y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=0))
sol.solve(WARM)
a0, g0 = sol.alpha.cpu().numpy().copy(), sol.neg_y_grad.cpu().numpy().copy()
lib = N.load()
lib.b2svm_debug_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
st1 = sol.solve(EPOCHS)
ref = sol.alpha.cpu().numpy().copy()
buf = np.zeros(EPOCHS * st1.n_ctas * W + EPOCHS * 2)
N.check(lib.b2svm_debug_steps(sol._engine.h, buf.ctypes.data, EPOCHS))
SV = buf[:EPOCHS * st1.n_ctas * W].reshape(EPOCHS, st1.n_ctas, W)[:, 0, :]
SS = buf[EPOCHS * st1.n_ctas * W:].reshape(EPOCHS, 2)
world = int(os.environ.get("WORLD", 2))
for rep in range(int(os.environ.get("REPS", W))):
    r = SH.EmulatedRanks(X, y, p, 1.0, 1e-5, spec, world)
    r.set_state(a0, g0)
    st = r.solve(EPOCHS)
    ncta = st[0].n_ctas
    views = []
    steps = None
    for s in r.solvers:
        buf = np.zeros(EPOCHS * ncta * W + EPOCHS * 2)
        N.check(lib.b2svm_debug_steps(s._engine.h, buf.ctypes.data, EPOCHS))
        views.append(buf[:EPOCHS * ncta * W].reshape(EPOCHS, ncta, W))
        steps = buf[EPOCHS * ncta * W:].reshape(EPOCHS, 2)
    V = np.concatenate(views, axis=1)                # [epoch][all CTAs][6]
    same = (V == V[:, :1, :]).all(axis=(1, 2))
    err = float(np.abs(r.gather("alpha") - ref).max())
    if same.all():
        print(f"rep {rep}: all {V.shape[1]} leader views agree over {EPOCHS} epochs; final alpha diff vs single {err:.3e}", flush=True)
        d = (V[:, 0, :] != SV).any(axis=1) | (steps != SS).any(axis=1)
        if d.any():
            e = int(np.argmax(d))
            print(f"   first epoch that differs from the single-GPU trajectory: {e}")
            for q in range(max(0, e - 1), min(EPOCHS, e + 2)):
                print(f"     epoch {q}: emu view {V[q, 0].tolist()} step {steps[q].tolist()}")
                print(f"     epoch {q}: one view {SV[q].tolist()} step {SS[q].tolist()}")
    else:
        e = int(np.argmin(same))
        bad = np.flatnonzero((V[e] != V[e, 0]).any(axis=1))
        print(f"rep {rep}: FIRST DISAGREEMENT at epoch {e}: CTAs {bad.tolist()} differ from CTA 0; alpha diff vs single {err:.3e}")
        print("   CTA 0 view  (si, sj, gi, gj, ai, aj, Kii, Kjj, Kij, ci, cj, rowsum | consumer ci, cj, rowsum, pair):", V[e, 0].tolist())
        for b in bad[:4]:
            print(f"   CTA {b} view:", V[e, b].tolist(), " rank", b // ncta)
        print("   previous epoch view:", V[e - 1, 0].tolist(), flush=True)
        print("   e-2 epoch view:", V[e - 2, 0].tolist(), flush=True)
        print("   SINGLE-GPU view of this epoch:", SV[e].tolist(), flush=True)
        vals = sorted(set(V[e, :, 6].tolist())), sorted(set(V[e, :, 7].tolist()))
        print("   distinct Kii / Kjj over CTAs:", vals, " per-CTA Kii:", V[e, :, 6].tolist(), " per-CTA Kjj:", V[e, :, 7].tolist(), flush=True)
