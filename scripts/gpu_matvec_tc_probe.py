"""Where does the tcgen05 mat-vec spend its time?  (B2SVM_TC_DEBUG stubs out one stage at a time.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200.solver import KernelSpec, kernel_matvec

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(5)
na, nb, d = 200_000, 400_000, int(os.environ.get("D", "32"))
Xa = torch.randn(na, d, generator=g).to(dev)
Xb = torch.randn(nb, d, generator=g).to(dev)
coef = torch.rand(na, generator=g, dtype=torch.float64).to(dev) + 0.1
spec = KernelSpec("rbf", gamma=1.0 / d)
os.environ["B2SVM_MATVEC_TC"] = "2"
for dbg in (0, 1, 2, 3, 6, 7):
    os.environ["B2SVM_TC_DEBUG"] = str(dbg)
    kernel_matvec(spec, Xa, coef, Xb); torch.cuda.synchronize()
    t0 = time.perf_counter(); kernel_matvec(spec, Xa, coef, Xb); torch.cuda.synchronize()
    print(f"debug={dbg}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
os.environ["B2SVM_MATVEC_TC"] = "0"
kernel_matvec(spec, Xa, coef, Xb); torch.cuda.synchronize()
t0 = time.perf_counter(); kernel_matvec(spec, Xa, coef, Xb); torch.cuda.synchronize()
print(f"ffma: {(time.perf_counter()-t0)*1e3:.1f} ms")
