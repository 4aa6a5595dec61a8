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
solver, shard, st = SH.fit_csvc_sharded(X, y01, 1.0, spec, 1e-5, 200, cache_bytes=0)
solver._engine.init_state(solver.p)
st = solver.solve(300)
lib = N.load()
NT_ALL = 64 * 16 + 64 + 320
buf = (C.c_longlong * NT_ALL)()
lib.b2svm_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
N.check(lib.b2svm_debug_timeline(solver._engine.h, buf, NT_ALL))
t = np.array(buf[:1024], dtype=np.int64).reshape(64, 16)
seg = t[2:60]
chain = [("join (0->1)", 0, 1), ("CTA winner (1->7)", 1, 7), ("publish L1 (7->2)", 7, 2), ("alpha park + L1 poll (2->3)", 2, 3),
         ("level 2: push + poll (3->4)", 3, 4), ("row / alpha fetch (4->8)", 4, 8), ("type barrier (8->9)", 8, 9),
         ("select + K_ij (9->10)", 9, 10), ("smo_step (10->11)", 10, 11), ("cache plan + hand-over (11->5)", 11, 5), ("bar (5->6)", 5, 6)]
for r in range(world):
    if r == rank:
        print(f"rank {rank}: n={n} world={world}: {st.device_ms*1e3/st.n_iter:.2f} us/it, ctas {st.n_ctas}", flush=True)
        tot = 0
        for nm, a, b in chain:
            v = np.median(seg[:, b] - seg[:, a])
            tot += v
            print(f"   {nm:34s} {v:9.0f}", flush=True)
        print(f"      poll detail: publish->poll entry {np.median(seg[:, 12] - seg[:, 2]):.0f}, first pass {np.median(seg[:, 13] - seg[:, 12]):.0f}, "
              f"loop total {np.median(seg[:, 15] - seg[:, 12]):.0f}, passes median {np.median(seg[:, 14]):.0f} max {seg[:, 14].max()}, reduce after {np.median(seg[:, 3] - seg[:, 15]):.0f}", flush=True)
        sw = np.median(seg[1:, 0] - seg[:-1, 6])
        print(f"   {'sweep (6->next 0)':34s} {sw:9.0f}   chain {tot:.0f} + sweep {sw:.0f} = {tot + sw:.0f} cycles", flush=True)
    dist.barrier()
dist.destroy_process_group()
