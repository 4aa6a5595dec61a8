"""Leader timeline of rank 0 / CTA 0 in the row-sharded solve (B2SVM_TIMELINE=1)."""
import os, sys, ctypes as C
os.environ["B2SVM_TIMELINE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S, sharded as SH, _native as N
n = int(os.environ.get("N", 200000))
X, y01 = G.classification(n=n, d=128, seed=0)
spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
solver, shard, st = SH.fit_csvc_sharded(X, y01, 1.0, spec, 1e-5, 200)
solver._engine.init_state(solver.p)
st = solver.solve(300)
lib = N.load()
NT_ALL = 64 * 16 + 64 + 320
buf = (C.c_longlong * NT_ALL)()
lib.b2svm_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
N.check(lib.b2svm_debug_timeline(solver._engine.h, buf, NT_ALL))
t = np.array(buf[:1024], dtype=np.int64).reshape(64, 16)
seg = t[2:60]
names = ["join(0->1)", "publish(1->2)", "box poll(2->3)", "final reduce(3->4)", "step(4->5)", "bar(5->6)", "sweep(6->next0)"]
d01 = [seg[:, 1] - seg[:, 0], seg[:, 2] - seg[:, 1], seg[:, 3] - seg[:, 2], seg[:, 4] - seg[:, 3], seg[:, 5] - seg[:, 4], seg[:, 6] - seg[:, 5], np.r_[seg[1:, 0] - seg[:-1, 6], 0]]
for r in range(world):
    if r == rank:
        print(f"rank {rank}: n={n} world={world}: {st.device_ms*1e3/st.n_iter:.2f} us/it, ctas {st.n_ctas}", flush=True)
        for nm, v in zip(names, d01):
            print(f"   {nm:30s} {np.median(v):9.0f}", flush=True)
        fine = [("join->CTA winners (1->7)", seg[:, 7] - seg[:, 1]), ("push record (7->2)", seg[:, 2] - seg[:, 7]), ("decide + issue loads (4->8)", seg[:, 8] - seg[:, 4]),
                ("rows + pair kernels (8->9)", seg[:, 9] - seg[:, 8]), ("alpha settled (9->10)", seg[:, 10] - seg[:, 9]), ("smo_step (10->11)", seg[:, 11] - seg[:, 10]),
                ("cache plan + publish (11->5)", seg[:, 5] - seg[:, 11])]
        for nm, v in fine:
            print(f"      {nm:30s} {np.median(v):9.0f}", flush=True)
    dist.barrier()
dist.destroy_process_group()
