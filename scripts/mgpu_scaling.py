"""Cold-path SMO iterations/s of the row-sharded solve at a chosen problem size (default: 8x the rows of C3),
next to the same problem on one GPU of the same box.  Strong scaling needs per-GPU sweeps that dwarf the ~6 us
exchange chain: at n = 1.6M every GPU of 8 sweeps what one GPU sweeps at C3.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/mgpu_scaling.py            # N_ROWS=1600000 ITERS=3000 by default
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S, sharded as SH

n, iters, d = int(os.environ.get("N_ROWS", 1_600_000)), int(os.environ.get("ITERS", 3000)), 128
X, y01 = G.classification(n=n, d=d, seed=0)
y = np.where(y01 == 1, 1.0, -1.0)
spec = S.KernelSpec("rbf", float(1 / (d * S.host_std(X))))
out = {"n_rows": n, "d": d, "iterations": iters, "world": world}

# ---- all GPUs, rows sharded ----
solver, shard, st = SH.fit_csvc_sharded(X, y01, 1.0, spec, 1e-5, 200, cache_bytes=0)
solver._engine.init_state(solver.p)
torch.cuda.synchronize(); dist.barrier()
st = solver.solve(iters)
t = torch.tensor([st.device_ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
out["sharded_us_per_iteration"] = float(t.item()) * 1e3 / st.n_iter
out["sharded_it_per_s"] = st.n_iter / (float(t.item()) * 1e-3)
a_sh = SH.gather_variables(solver.alpha.cpu().numpy(), shard, False)
del solver
torch.cuda.empty_cache()
dist.barrier()

# ---- one GPU, same problem, same box (rank 0) ----
if rank == 0:
    ref = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, spec, cache_bytes=0))
    ref.solve(200)
    ref._engine.init_state(ref.p)
    s1 = ref.solve(iters)
    out["single_us_per_iteration"] = s1.device_ms * 1e3 / s1.n_iter
    out["single_it_per_s"] = s1.n_iter / (s1.device_ms * 1e-3)
    out["speedup"] = out["sharded_it_per_s"] / out["single_it_per_s"]
    out["bit_identical_alpha"] = bool(np.array_equal(a_sh, ref.alpha.cpu().numpy()))
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
