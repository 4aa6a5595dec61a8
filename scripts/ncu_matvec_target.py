"""One large K7 call (400k queries x 200k support rows, d = 32, rbf) for an ncu capture of matvec_tc_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysvm_b200.solver import KernelSpec, kernel_matvec
g = torch.Generator().manual_seed(5)
d = int(os.environ.get("D", 32))
Xa = torch.randn(200_000, d, generator=g).cuda()
Xb = torch.randn(400_000, d, generator=g).cuda()
coef = (torch.rand(200_000, generator=g, dtype=torch.float64) + 0.1).cuda()
for _ in range(2):
    out = kernel_matvec(KernelSpec("rbf", 1.0 / d), Xa, coef, Xb)
torch.cuda.synchronize()
print("ok", float(out[:8].sum()))
