"""K8 / Gram on the tensor cores at BASELINE configs[4] size: NormalRFF.transform 1M x 32 -> D = 4096 (float64 and
float32 features), the dense kernel matrix, and the float64 CUDA-core kernels beside them.  CUDA-event times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysvm_b200 import synth, solver as S
from pysvm_b200.rff import NormalRFF

PEAK = 6535.1
N = int(os.environ.get("ROWS", 1_000_000))
QUICK = bool(os.environ.get("QUICK"))
dev = torch.device("cuda")
X = torch.from_numpy(synth.oneclass(n=N, d=32, seed=1)).to(dev)
np.random.seed(0)
rff = NormalRFF(gamma=1 / 32, D=4096).fit(np.ones((1, 32)))


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        del r
    return best


for dt, name in ((torch.float64, "float64"), (torch.float32, "float32")):
    ms = timed(lambda: rff.transform_device(X, dtype=dt), reps=1 if QUICK else 3)
    by = N * 4096 * (8 if dt == torch.float64 else 4)
    print(f"rff_transform {N} x 32 -> 4096 {name}: {ms:.2f} ms = {N / ms * 1e3:.3e} rows/s, {by / ms / 1e6:.0f} GB/s written = {by / ms / 1e6 / PEAK:.2f} of the measured HBM peak, {N * 4096 / ms * 1e3:.3e} cos/s", flush=True)
if not QUICK:
    os.environ["B2SVM_RFF_TC"] = "0"
    Xs = X[:100_000]
    ms = timed(lambda: rff.transform_device(Xs), reps=1)
    print(f"rff_transform (float64 CUDA-core kernel, exact) 100000 x 32 -> 4096: {ms:.2f} ms = {1e5 / ms * 1e3:.3e} rows/s, {1e5 * 4096 * 8 / ms / 1e6:.0f} GB/s", flush=True)
    del os.environ["B2SVM_RFF_TC"]
    for n, d in ((20000, 64), (40000, 32)):
        A = torch.randn((n, d), device=dev)
        ms = timed(lambda: S.kernel_matrix(S.KernelSpec("rbf", 1.0 / d), A), reps=2)
        print(f"kernel_matrix rbf float32 {n} x {n}, d={d} (tcgen05): {ms:.2f} ms = {n * n / ms * 1e3:.3e} kernel values/s, {n * n * 4 / ms / 1e6:.0f} GB/s written", flush=True)
    A = torch.randn((6000, 64), device=dev, dtype=torch.float64)
    ms = timed(lambda: S.kernel_matrix(S.KernelSpec("rbf", 1.0 / 64), A), reps=2)
    print(f"kernel_matrix rbf float64 6000 x 6000, d=64 (FMA kernel): {ms:.2f} ms = {36e6 / ms * 1e3:.3e} kernel values/s", flush=True)
