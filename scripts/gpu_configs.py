"""Full-size runs of BASELINE.json configs[3] (C4) and configs[4] (C5) -- do they run, how long do they take."""
import os, sys, time, contextlib, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200 import synth as G
import pysvm_b200 as P
from pysvm_b200.rff import NormalRFF
t00 = time.time()
def log(*a): print(f"[{time.time()-t00:6.1f}s]", *a, flush=True)
def timed(f):
    torch.cuda.synchronize(); t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        r = f()
    torch.cuda.synchronize(); return r, time.perf_counter() - t
which = sys.argv[1:] or ["c4", "c5"]
if "c4" in which:
    X, z = G.regression(n=100_000, d=64, seed=0)
    est, dt = timed(lambda: P.KernelSVR(max_iter=10 ** 7).fit(X, z))
    s = est.solve_stats_
    log(f"C4 KernelSVR l=100k d=64: fit {dt:.2f} s, {s['n_iter']} it ({s['device_ms']*1e3/max(1,s['n_iter']):.2f} us/it), cold {s['cold_sweeps']} warm {s['cache_hits']//2}, converged {s['converged']}, R2(first 20k) {est.score(X[:20000], z[:20000]):.4f}")
    est, dt = timed(lambda: P.NuSVR(max_iter=200_000).fit(X, z))
    s = est.solve_stats_
    log(f"C4 NuSVR l=100k d=64 max_iter=2e5: fit {dt:.2f} s, {s['n_iter']} it ({s['device_ms']*1e3/max(1,s['n_iter']):.2f} us/it), cold {s['cold_sweeps']} warm {s['cache_hits']//2}, sum(alpha) {est.alpha_.sum():.3f} (C*l*nu = 50000)")
if "c5" in which:
    X = G.oneclass(n=1_000_000, d=32, seed=1)
    est, dt = timed(lambda: P.OneClassSVM(max_iter=10 ** 7).fit(X))
    s = est.solve_stats_
    log(f"C5 OneClassSVM n=1M d=32: fit {dt:.2f} s, {s['n_iter']} it ({s['device_ms']*1e3/max(1,s['n_iter']):.2f} us/it), cold {s['cold_sweeps']} warm {s['cache_hits']//2}, nSV {int((est.alpha_>0).sum())}, converged {s['converged']}")
    Xq = np.random.default_rng(2).standard_normal((1_000_000, 32)).astype(np.float32)
    dec, dt = timed(lambda: est.decision_function(Xq))
    nsv = int((est.alpha_ > 0).sum())
    log(f"C5 decision_function 1M queries x {nsv} SVs: {dt:.2f} s -> {1e6*nsv/dt:.3e} kernel evals/s (reference CPU: 5.6e7); outliers {(dec<0).mean():.3f}")
    # the full 10M-query workload, device-resident query blocks of 1M rows (the host copy is timed above)
    g = torch.Generator(device="cuda").manual_seed(2)
    tot = 0.0
    for _ in range(10):
        Xd = torch.randn(1_000_000, 32, generator=g, device="cuda")
        _, dt = timed(lambda: est.decision_function(Xd))
        tot += dt
    log(f"C5 decision_function 10M device-resident queries x {nsv} SVs: {tot:.2f} s -> {1e7*nsv/tot:.3e} kernel evals/s")
    np.random.seed(0)
    r = NormalRFF(gamma=est.gamma_, D=4096).fit(X[:1])
    Z, dt = timed(lambda: r.transform_device(torch.from_numpy(X[:200_000]).cuda()))
    log(f"C5 NormalRFF.transform 200k x 32 -> 4096 (float64): {dt:.2f} s -> {200_000/dt:.3e} rows/s (reference CPU: 5.8e3)")
    # RFF decision form over the 10M queries: z(x_q) . wz, never materialising z (tensor cores, then the float64 kernel)
    from pysvm_b200.rff import rff_decision_device
    w, b = r._device_params(torch.device("cuda", 0))
    wz = torch.randn(4096, dtype=torch.float64, device="cuda")
    for mode, nblk in (("1", 10), ("0", 1)):
        os.environ["B2SVM_RFF_TC"] = mode
        g = torch.Generator(device="cuda").manual_seed(2)
        tot = 0.0
        for _ in range(nblk):
            Xd = torch.randn(1_000_000, 32, generator=g, device="cuda")
            rff_decision_device(Xd[:2048], w, b, wz)
            _, dt = timed(lambda: rff_decision_device(Xd, w, b, wz))
            tot += dt
        log(f"C5 RFF decision (D=4096) {nblk}M queries, {'tcgen05 3xTF32' if mode == '1' else 'float64 CUDA cores'}: {tot:.3f} s -> {nblk*1e6/tot:.3e} rows/s, {nblk*1e6*4096/tot:.3e} cos/s")
