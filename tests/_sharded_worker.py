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
    assert est.rho_ == ref.rho_
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
dist.destroy_process_group()
