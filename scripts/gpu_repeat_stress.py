"""Run-to-run determinism stress: the same small problems solved REPS times each; every run must give the same
iteration count and the same alpha bytes.  Sizes span one CTA up to the full grid, one-class (rows_kernel_sum
initial gradient) and C-SVC, with and without the column cache."""
import os, sys, hashlib, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smo_oracle as O
from pysvm_b200 import solver as S
from pysvm_b200.svm import oc_svm, svc

REPS = int(os.environ.get("REPS", 200))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def tally(name, fn):
    seen = collections.Counter()
    for _ in range(REPS):
        seen[fn()] += 1
    flag = "OK " if len(seen) == 1 else "BAD"
    print(f"{flag} {name}: {dict(seen)}", flush=True)
    return len(seen) == 1


ok = True
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "g5_oneclass.npz"))
Xg = g["X"]


def oc_fit():
    est = oc_svm.OneClassSVM(nu=.1, gamma=.1, max_iter=10 ** 5).fit(Xg)
    return est.n_iter_, digest(est.alpha_)


ok &= tally(f"g5 one-class n={len(Xg)}", oc_fit)

for n, d, cache in ((64, 8, 0), (300, 16, 0), (1500, 32, 0), (1500, 32, 64 << 20), (6000, 128, 0), (20000, 64, 256 << 20)):
    X, y01 = O.synth_classification(n=n, d=d, seed=3)
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (d * X.std()))
    iters = 3000 if n <= 6000 else 1500

    def fit():
        sol = S.SolverWithCache(-np.ones(n), y, 1.0, 1e-5, func=S.QColumns(X, y, spec, cache_bytes=cache))
        st = sol.solve(iters)
        return st.n_iter, st.n_ctas, digest(sol.alpha.cpu().numpy())

    ok &= tally(f"svc n={n} d={d} cache={cache >> 20}MB", fit)

print("ALL DETERMINISTIC" if ok else "NONDETERMINISM FOUND")
