"""K7 (kernel mat-vec on the tensor cores) at configs[4] size: 1M queries x 500k support rows, d = 32, and d = 64/128."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200.solver import KernelSpec, kernel_matvec
g = torch.Generator().manual_seed(5)
for d, nsv, nq in ((32, 500_000, 1_000_000), (64, 200_000, 400_000), (128, 100_000, 200_000)):
    Xa = torch.randn(nsv, d, generator=g).cuda(); Xb = torch.randn(nq, d, generator=g).cuda()
    coef = (torch.rand(nsv, generator=g, dtype=torch.float64) + 0.1).cuda()
    for mt in ("1", "2"):
        os.environ["B2SVM_TC_MT"] = mt
        out = kernel_matvec(KernelSpec("rbf", 1.0 / d), Xa, coef, Xb); torch.cuda.synchronize()
        t0 = time.perf_counter(); out2 = kernel_matvec(KernelSpec("rbf", 1.0 / d), Xa, coef, Xb); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"d={d} {nq} x {nsv} MT={mt}: {dt*1e3:.1f} ms = {nq*nsv/dt:.3e} kernel evaluations/s  checksum {float(out2[:1000].sum()):.9e}", flush=True)
    del os.environ["B2SVM_TC_MT"]
