"""Row-sharded solve across the GPUs of one box vs the single-GPU solve of the same problem.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from pysvm_b200 import synth as G
from pysvm_b200 import solver as S, sharded as SH

def log(*a):
    if rank == 0:
        print(*a, flush=True)

for n, iters in ((20000, 1500), (200000, 4000)):
    X, y01 = G.classification(n=n, d=128, seed=0)
    y = np.where(y01 == 1, 1.0, -1.0)
    spec = S.KernelSpec("rbf", 1 / (128 * X.std()))
    solver, shard, st = SH.fit_csvc_sharded(X, y01, 1.0, spec, 1e-5, iters)
    st = None
    # timed repeat
    solver._engine.init_state(solver.p)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    st = solver.solve(iters)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    a_loc = solver.alpha.cpu().numpy()
    g_loc = solver.neg_y_grad.cpu().numpy()
    rho = solver.calculate_rho()
    pad = np.zeros(shard.rows_per_rank); pad[:shard.n_local] = a_loc
    padg = np.zeros(shard.rows_per_rank); padg[:shard.n_local] = g_loc
    a_all = SH.all_gather_array(pad).reshape(-1)[:n] if world * shard.rows_per_rank >= n else None
    g_all = SH.all_gather_array(padg).reshape(-1)[:n]
    dec = SH.decision_sharded(solver, solver.alpha * solver._engine.func.y, X[:256])
    if rank == 0:
        ref = S.SolverWithCache(-np.ones(n), y, 1.0, func=S.QColumns(X, y, spec))
        rs = ref.solve(iters)
        ra, rg = ref.alpha.cpu().numpy(), ref.neg_y_grad.cpu().numpy()
        rdec = S.kernel_matvec(spec, ref._engine.func.X, ref.alpha * ref._engine.func.y,
                               S.to_device_matrix(X[:256], ref.device)).cpu().numpy()
        log(f"n={n} world={world}: sharded {st.n_iter} it, kernel {st.device_ms*1e3/st.n_iter:.2f} us/it (wall {dt*1e6/st.n_iter:.2f}), ctas/gpu {st.n_ctas};"
            f" single-GPU {rs.device_ms*1e3/rs.n_iter:.2f} us/it")
        log(f"   max|alpha-single|={np.abs(a_all-ra).max():.3e} max|G-single|={np.abs(g_all-rg).max():.3e} "
            f"rho {rho:.12f} vs {ref.calculate_rho():.12f}  max|dec-single|={np.abs(dec-rdec).max():.3e} "
            f"next pair {(st.last_i, st.last_j)} vs {(rs.last_i, rs.last_j)}")
    dist.barrier()
dist.destroy_process_group()
