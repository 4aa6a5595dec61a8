"""Worker of tests/test_gpu_sharded.py (one process per GPU, NCCL): the row-sharded solve must
reproduce the single-GPU solve bit for bit -- same kernel code, same reduction order rules."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from oracle import smo_oracle as O
from pysvm_b200 import sharded as SH
from pysvm_b200 import solver as S
from pysvm_b200.svm.svc import BiKernelSVC
from pysvm_b200.svm.svr import KernelSVR

# ---- classification, estimator level (sharded.enable routes fit through the process group) ------------
X, y01 = O.synth_classification(n=6000, d=128, seed=1)
SH.enable()
est = BiKernelSVC(max_iter=10 ** 6).fit(X, y01)
dec = est.decision_function(X[:300])
SH.disable()
if rank == 0:
    ref = BiKernelSVC(max_iter=10 ** 6).fit(X, y01)
    assert est.n_iter_ == ref.n_iter_, (est.n_iter_, ref.n_iter_)
    assert np.array_equal(est.alpha_, ref.alpha_)
    assert abs(est.rho_ - ref.rho_) <= 1e-13 * max(1.0, abs(ref.rho_))     # a mean over the free SVs: summed per rank / per block
    np.testing.assert_allclose(dec, ref.decision_function(X[:300]), rtol=0, atol=1e-12)
    gold = O.fit_csvc(X, y01, max_iter=10 ** 6)
    assert abs(est.n_iter_ - gold.n_iter) <= max(1, 0.01 * gold.n_iter)
    assert np.array_equal(est.support_, np.flatnonzero(gold.alpha > 0))
dist.barrier()

# ---- regression (mirror layout: variables t and t+l share row t), solver level ---------------------------
Xr, z = O.synth_regression(n=3000, d=64, seed=0)
spec = S.KernelSpec("rbf", 1 / (64 * Xr.std()))
solver, shard, st = SH.fit_svr_sharded(Xr, z, 1.0, 0.0, spec, 1e-5, 3000)
a_all = SH.gather_variables(solver.alpha.cpu().numpy(), shard, True)
if rank == 0:
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ref = KernelSVR(max_iter=3000).fit(Xr, z)
    assert st.n_iter == ref.n_iter_ == 3000
    assert np.array_equal(a_all, ref.alpha_), float(np.abs(a_all - ref.alpha_).max())
    print("sharded worker ok: world", world, "csvc iters", est.n_iter_, flush=True)
dist.barrier()

# ---- nu variants and one-class: four candidate types per iteration, initial gradient from the replicated X --------
# (the single-GPU and the sharded initial mat-vec split their float64 sums differently, so the start gradients agree
#  to ~1e-15 rather than bit for bit: solutions are compared within 1e-7, iteration counts within 1 %)
import contextlib, io
from pysvm_b200.svm.svc import BiNuSVC
from pysvm_b200.svm.svr import NuSVR
from pysvm_b200.svm.oc_svm import OneClassSVM


def quiet(f):
    with contextlib.redirect_stdout(io.StringIO()):
        return f()


Xo = O.synth_oneclass(n=5000, d=32, seed=1)
SH.enable()
oc = quiet(lambda: OneClassSVM(max_iter=10 ** 6).fit(Xo))
oc_dec = oc.decision_function(Xo[:200])
Xn, yn = O.synth_classification(n=4000, d=20, seed=2)
nus = quiet(lambda: BiNuSVC(max_iter=4000).fit(Xn, yn))
nus_dec = nus.decision_function(Xn[:200])
nur = quiet(lambda: NuSVR(max_iter=3000).fit(Xr, z))
nur_dec = nur.decision_function(Xr[:200])
ksr = quiet(lambda: KernelSVR(max_iter=3000).fit(Xr, z))
ksr_dec = ksr.decision_function(Xr[:200])
SH.disable()
if rank == 0:
    ref = quiet(lambda: OneClassSVM(max_iter=10 ** 6).fit(Xo))
    assert abs(oc.n_iter_ - ref.n_iter_) <= max(1, 0.01 * ref.n_iter_), (oc.n_iter_, ref.n_iter_)
    np.testing.assert_allclose(oc.alpha_, ref.alpha_, rtol=0, atol=1e-5)
    np.testing.assert_allclose(oc.rho_, ref.rho_, rtol=1e-6)
    np.testing.assert_allclose(oc_dec, ref.decision_function(Xo[:200]), rtol=0, atol=1e-5)
    ref = quiet(lambda: BiNuSVC(max_iter=4000).fit(Xn, yn))
    assert nus.n_iter_ == ref.n_iter_ == 4000
    np.testing.assert_allclose(nus.alpha_, ref.alpha_, rtol=0, atol=1e-7)
    np.testing.assert_allclose([nus.rho_, nus.b_], [ref.rho_, ref.b_], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(nus_dec, ref.decision_function(Xn[:200]), rtol=1e-6, atol=1e-7)
    ref = quiet(lambda: NuSVR(max_iter=3000).fit(Xr, z))
    assert nur.n_iter_ == ref.n_iter_ == 3000
    np.testing.assert_allclose(nur.alpha_, ref.alpha_, rtol=0, atol=1e-7)
    np.testing.assert_allclose(nur_dec, ref.decision_function(Xr[:200]), rtol=1e-6, atol=1e-7)
    ref = quiet(lambda: KernelSVR(max_iter=3000).fit(Xr, z))
    assert np.array_equal(ksr.alpha_, ref.alpha_)
    np.testing.assert_allclose(ksr_dec, ref.decision_function(Xr[:200]), rtol=0, atol=1e-12)
    print("sharded worker ok: nu-SVC, nu-SVR, one-class, KernelSVR estimators", flush=True)
dist.barrier()

# ---- one-vs-one sub-problems scheduled across the GPUs (no communication during the solves) -----------------------
from sklearn.datasets import load_digits
from pysvm_b200.svm.svc import KernelSVC
Xd, yd = load_digits(return_X_y=True)
SH.enable()
ovo = quiet(lambda: KernelSVC(multiclass="ovo", max_iter=10 ** 5).fit(Xd, yd))
pred = ovo.predict(Xd)
SH.disable()
ref = quiet(lambda: KernelSVC(multiclass="ovo", max_iter=10 ** 5).fit(Xd, yd))
it_sh = sum(e.n_iter_ for e in ovo.multiclass_model.estimators_)
it_ref = sum(e.n_iter_ for e in ref.multiclass_model.estimators_)
assert it_sh == it_ref, (it_sh, it_ref)
for a, b in zip(ovo.multiclass_model.estimators_, ref.multiclass_model.estimators_):
    assert np.array_equal(a.alpha_, b.alpha_) and abs(a.rho_ - b.rho_) <= 1e-13 * max(1.0, abs(b.rho_))
assert np.array_equal(pred, ref.predict(Xd))
if rank == 0:
    print("sharded worker ok: 45 one-vs-one sub-problems over", world, "GPUs,", it_sh, "iterations", flush=True)
dist.barrier()
dist.destroy_process_group()
