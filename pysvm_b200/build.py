"""Build libb2svm.so in-tree with nvcc for sm_100a (B200).

    python -m pysvm_b200.build [--force] [--verbose]

The shared library lands in ``pysvm_b200/_lib/`` (git-ignored, shipped to the GPU box by gpurun).
nvcc cross-compiles without a GPU; cudart is linked statically so the library loads (and its
symbols can be checked) on a box without a driver.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libb2svm.so")
SOURCES = ["b2svm_api.cu", "b2svm_dense.cu", "b2svm_matvec_tc.cu", "b2svm_host.cu", "b2svm_launch_f32_plain.cu", "b2svm_launch_f32_generic.cu",
           "b2svm_launch_f64_plain.cu", "b2svm_launch_f64_generic.cu"]
HEADERS = ["b2svm_device.cuh", "b2svm_smo.cuh", "b2svm_launch.cuh", "b2svm_internal.h", os.path.join("..", "..", "include", "b2svm.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # fused multiply-adds are explicit fma()/fmaf() calls only
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",   # host code too: b2svm_host_std must round like numpy (no FMA contraction on aarch64 / -march=native)
    "-Xptxas", "-v",
]
if os.environ.get("B2SVM_TIMELINE"):      # per-phase clock stamps inside the persistent kernel (scripts/gpu_timeline.py)
    NVCC_FLAGS.append("-DB2SVM_TIMELINE")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libb2svm.digest")
    digest = source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    objs = []
    t0 = time.time()
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    print(f"built {LIB} in {time.time() - t0:.1f}s")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
