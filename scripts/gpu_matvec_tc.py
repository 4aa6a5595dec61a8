"""K7 (kernel mat-vec) on the tensor cores vs the FFMA kernel vs a float64 torch reference.

    python scripts/gpu_matvec_tc.py            # accuracy at several shapes + timing at the C5 shape
Each variant is selected through B2SVM_MATVEC_TC (0 = FFMA kernel, 2 = tcgen05 forced); the env is read per call.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pysvm_b200.solver import KernelSpec, kernel_matvec


def truth(kind, gamma, coef0, degree, Xa, coef, Xb, block=8192):
    Xa64, c64 = Xa.double(), coef.double()
    out = torch.empty(Xb.shape[0], dtype=torch.float64, device=Xb.device)
    for s in range(0, Xb.shape[0], block):
        q = Xb[s:s + block].double()
        dot = q @ Xa64.T
        if kind == "rbf":
            k = torch.exp(-gamma * ((q * q).sum(1)[:, None] + (Xa64 * Xa64).sum(1)[None, :] - 2 * dot).clamp_min(0))
        elif kind == "poly":
            k = (gamma * dot + coef0) ** degree
        elif kind == "sigmoid":
            k = torch.tanh(gamma * dot + coef0)
        else:
            k = dot
        out[s:s + block] = k @ c64
    return out


def run(mode, spec, Xa, coef, Xb, reps=1):
    os.environ["B2SVM_MATVEC_TC"] = str(mode)
    out = kernel_matvec(spec, Xa, coef, Xb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = kernel_matvec(spec, Xa, coef, Xb)
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    worst = 0.0
    for (na, nb, d, kind) in [(1000, 777, 32, "rbf"), (4096, 4096, 64, "rbf"), (3333, 2000, 7, "rbf"), (2048, 1500, 33, "poly"),
                              (2048, 1500, 48, "sigmoid"), (5000, 129, 20, "linear"), (130, 5000, 64, "rbf"),
                              (3000, 2500, 100, "rbf"), (2000, 3000, 128, "rbf"), (1500, 1500, 96, "poly")]:
        Xa = torch.randn(na, d, generator=g).to(dev)
        Xb = torch.randn(nb, d, generator=g).to(dev)
        coef = torch.randn(na, generator=g, dtype=torch.float64).to(dev)
        coef[torch.rand(na, generator=g).to(dev) < 0.4] = 0.0
        gamma = 1.0 / d
        spec = KernelSpec(kind, gamma=gamma, coef0=0.5 if kind in ("poly", "sigmoid") else 0.0, degree=3)
        ref = truth(kind, gamma, spec.coef0, 3, Xa, coef, Xb)
        scale = ref.abs().max().item()
        o_tc, _ = run(2, spec, Xa, coef, Xb)
        o_ff, _ = run(0, spec, Xa, coef, Xb)
        e_tc = (o_tc - ref).abs().max().item() / scale
        e_ff = (o_ff - ref).abs().max().item() / scale
        worst = max(worst, e_tc)
        print(f"na={na} nb={nb} d={d} {kind:8s} rel err: tcgen05 {e_tc:.2e}   ffma {e_ff:.2e}", flush=True)
    # C5 shape: 1M queries x 500k support rows, d = 32
    na, nb, d = 500_000, 1_000_000, 32
    Xa = torch.randn(na, d, generator=g).to(dev)
    Xb = torch.randn(nb, d, generator=g).to(dev)
    coef = torch.rand(na, generator=g, dtype=torch.float64).to(dev) + 0.1
    spec = KernelSpec("rbf", gamma=1.0 / d)
    o_tc, t_tc = run(2, spec, Xa, coef, Xb, reps=2)
    o_ff, t_ff = run(0, spec, Xa, coef, Xb, reps=1)
    pairs = na * nb
    rel = ((o_tc - o_ff).abs().max() / o_ff.abs().max()).item()
    print(f"C5 decision 1M x 500k d=32: tcgen05 {t_tc*1e3:.1f} ms ({pairs/t_tc/1e12:.2f} Tpair/s, "
          f"{3*2*pairs*d/t_tc/1e12:.1f} TFLOP/s tf32 issued)   ffma {t_ff*1e3:.1f} ms   max rel diff {rel:.2e}")
    d = 64
    Xa = torch.randn(200_000, d, generator=g).to(dev)
    Xb = torch.randn(400_000, d, generator=g).to(dev)
    coef = torch.rand(200_000, generator=g, dtype=torch.float64).to(dev) + 0.1
    spec = KernelSpec("rbf", gamma=1.0 / d)
    o_tc, t_tc = run(2, spec, Xa, coef, Xb, reps=2)
    o_ff, t_ff = run(0, spec, Xa, coef, Xb, reps=1)
    rel = ((o_tc - o_ff).abs().max() / o_ff.abs().max()).item()
    print(f"400k x 200k d=64: tcgen05 {t_tc*1e3:.1f} ms   ffma {t_ff*1e3:.1f} ms   max rel diff {rel:.2e}")
    d = 128
    Xa = torch.randn(100_000, d, generator=g).to(dev)
    Xb = torch.randn(200_000, d, generator=g).to(dev)
    coef = torch.rand(100_000, generator=g, dtype=torch.float64).to(dev) + 0.1
    spec = KernelSpec("rbf", gamma=1.0 / d)
    o_tc, t_tc = run(2, spec, Xa, coef, Xb, reps=2)
    o_ff, t_ff = run(0, spec, Xa, coef, Xb, reps=1)
    rel = ((o_tc - o_ff).abs().max() / o_ff.abs().max()).item()
    print(f"200k x 100k d=128 (C3 decision): tcgen05 {t_tc*1e3:.1f} ms   ffma {t_ff*1e3:.1f} ms   max rel diff {rel:.2e}")
    assert worst < 2e-6, worst
    print("OK")


if __name__ == "__main__":
    main()
