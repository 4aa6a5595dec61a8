"""Do all CTA leaders see the same (i, j, g_i, g_j, alpha_i, alpha_j) every epoch?  (B2SVM_DEBUG_STEPS=1)"""
import os, sys, ctypes as C
os.environ["B2SVM_DEBUG_STEPS"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import smo_oracle as O
from pysvm_b200 import solver as S, _native as N
n, d = 200000, 128
X, y01 = O.synth_classification(n=n, d=d, seed=0)
y = np.where(y01 == 1, 1.0, -1.0)
sol = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, S.KernelSpec("rbf", 1 / (d * X.std()))))
E = 2000
st = sol.solve(E)
lib = N.load()
ncta = st.n_ctas
raw = np.zeros(E * ncta * 6 + E * 2)
buf = raw[:E * ncta * 6].reshape(E, ncta, 6)
new = raw[E * ncta * 6:].reshape(E, 2)
lib.b2svm_debug_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
N.check(lib.b2svm_debug_steps(sol._engine.h, raw.ctypes.data, E))
ref = buf[:, :1, :]
bad = np.argwhere(buf != ref)
print("epochs", E, "ctas", ncta, "mismatching entries", len(bad))
if len(bad):
    e0 = bad[:, 0].min()
    print("first divergent epoch", e0)
    for e in range(e0, min(E, e0 + 3)):
        rows = buf[e]
        uniq, counts = np.unique(rows, axis=0, return_counts=True)
        print(" epoch", e, "distinct views:", len(uniq))
        for u, c in zip(uniq[:6], counts[:6]):
            print("    x%d: i=%d j=%d gi=%.17g gj=%.17g ai=%.17g aj=%.17g" % (c, u[0], u[1], u[2], u[3], u[4], u[5]))
        print("    previous epoch view: i=%d j=%d" % (buf[e-1,0,0], buf[e-1,0,1]))

# replay: does every leader read the alpha the previous step stored?
last = {}
nbad = 0
for e in range(E):
    si, sj, gi, gj, ai, aj = buf[e, 0]
    si, sj = int(si), int(sj)
    for idx, a in ((si, ai), (sj, aj)):
        want = last.get(idx, 0.0)
        if a != want:
            nbad += 1
            if nbad <= 10:
                print(f"epoch {e}: alpha[{idx}] read {a!r} but last stored {want!r}; previous pair {tuple(int(v) for v in buf[e-1,0,:2])}")
    last[si], last[sj] = new[e, 0], new[e, 1]
print("stale alpha reads:", nbad)
a_dev = sol.alpha.cpu().numpy()
mism = [(k, v, a_dev[k]) for k, v in last.items() if a_dev[k] != v]
print("final alpha mismatches vs replay:", len(mism), mism[:5])
