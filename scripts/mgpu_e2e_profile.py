"""Where the per-fit set-up time of a row-sharded fit goes (host side): torchrun --nproc-per-node N scripts/mgpu_e2e_profile.py"""
import os, sys, time, collections, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pysvm_b200 import sharded, solver as S, synth as G
from pysvm_b200.svm import svc
from pysvm_b200.svm.svc import BiKernelSVC

acc = collections.defaultdict(float)


def timed(owner, name, label=None):
    f = getattr(owner, name)
    label = label or name

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t0
    setattr(owner, name, w)


def fine_gather(a, group=None):
    dev = sharded._group_device(group)
    w = dist.get_world_size(group)
    t0 = time.perf_counter()
    a = np.ascontiguousarray(a, dtype=np.float64)
    t = torch.from_numpy(a.reshape(-1)).to(dev)
    out = torch.empty(w * t.numel(), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dist.all_gather_into_tensor(out, t, group=group)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    r = out.cpu().numpy().reshape((w,) + a.shape)
    t3 = time.perf_counter()
    acc["  gather: h2d"] += t1 - t0
    acc["  gather: nccl all_gather (+sync)"] += t2 - t1
    acc["  gather: d2h"] += t3 - t2
    return r


if os.environ.get("FINE", "1") == "1" or os.environ.get("FINE") == "gather":
    sharded.all_gather_array = fine_gather


def fine_to_device(X, device):
    t0 = time.perf_counter()
    X = np.asarray(X)
    t = torch.from_numpy(np.ascontiguousarray(X))
    t1 = time.perf_counter()
    pinned = t.is_pinned()
    t2 = time.perf_counter()
    d = torch.empty(t.shape, dtype=t.dtype, device=device)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    d.copy_(t)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    acc["  X: from_numpy"] += t1 - t0
    acc[f"  X: is_pinned -> {pinned}"] += t2 - t1
    acc["  X: device alloc"] += t3 - t2
    acc["  X: copy"] += t4 - t3
    return d


if os.environ.get("FINE", "1") == "1" or os.environ.get("FINE") == "x":
    S.to_device_matrix = fine_to_device
timed(S, "_dev_f64")
timed(svc, "host_std")
timed(sharded, "all_gather_bytes")
timed(sharded, "all_gather_array")
timed(sharded, "gather_variables")
timed(sharded.ShardedColumns, "__init__", "ShardedColumns (H2D of the slice)")
timed(S._Engine, "__init__", "engine create")
timed(S._Engine, "attach_mailboxes")
timed(S._Engine, "export_mailbox")
timed(S._Engine, "prepare")
timed(S._Engine, "solve", "engine solve (incl. prepare + barrier)")
timed(S._Engine, "close", "engine destroy")
timed(S._Engine, "init_state")
timed(dist, "barrier", "dist.barrier")

X, y01 = G.classification(n=200000, d=128, seed=0)
Xp = torch.from_numpy(X).pin_memory().numpy()
sharded.enable()
with contextlib.redirect_stdout(io.StringIO()):
    est = BiKernelSVC(max_iter=20000).fit(Xp, y01)
acc.clear()
dist.barrier()
t0 = time.perf_counter()
K = 3
dev_ms = 0.0
for _ in range(K):
    with contextlib.redirect_stdout(io.StringIO()):
        est = BiKernelSVC(max_iter=20000).fit(Xp, y01)
    dev_ms += est.solve_stats_["device_ms"]
dist.barrier()
dt = time.perf_counter() - t0
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: {1e3 * dt / K:.1f} ms per fit, device {dev_ms / K:.1f} ms; host phases (ms per fit, nested ones overlap):")
        for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
            print(f"    {k:45s} {1e3 * v / K:8.2f}")
        sys.stdout.flush()
dist.destroy_process_group()
