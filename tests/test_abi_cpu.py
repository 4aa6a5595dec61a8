"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads without a GPU, exports
every symbol include/b2svm.h declares, and the host-side logic around it is right."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pysvm_b200 import _native as N
    from pysvm_b200.build import build
    build()
    lib = N.load()
    header = open(os.path.join(ROOT, "include", "b2svm.h")).read()
    declared = set(re.findall(r"\b(b2svm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2svm_abi_version() == N.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the ABI structs: sizes pinned so a header edit cannot drift silently."""
    from pysvm_b200 import _native as N
    assert ctypes.sizeof(N.KernelDesc) == 32
    assert ctypes.sizeof(N.ProblemDesc) == 176
    assert ctypes.sizeof(N.SolveStats) == 104


def test_errors_are_reported_not_swallowed():
    from pysvm_b200 import _native as N
    lib = N.load()
    h = ctypes.c_void_p()
    rc = lib.b2svm_create(None, ctypes.byref(h))
    assert rc == -1 and b"null" in lib.b2svm_last_error()
    d = N.ProblemDesc()
    d.struct_bytes = 7
    assert lib.b2svm_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"size mismatch" in lib.b2svm_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pysvm_b200.svm.svc import BiKernelSVC
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BiKernelSVC().fit(np.zeros((4, 2)), np.array([0, 1, 0, 1]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysvm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), os.path.join(dirpath, f)


def test_nu_initial_alpha_matches_sequential_fill():
    from pysvm_b200.solver import nu_initial_alpha
    from oracle.smo_oracle import nu_fill
    rng = np.random.default_rng(0)
    for t in (3.0, 7.5, 284.5, 0.3, 1000.0, 221.0):
        y = np.where(rng.random(700) < 0.4, 1.0, -1.0)
        np.testing.assert_array_equal(nu_initial_alpha(y, t), nu_fill(y, t))


def test_oneclass_initial_alpha_matches_oracle():
    from pysvm_b200.svm.oc_svm import oneclass_initial_alpha
    from oracle.smo_oracle import oneclass_alpha0
    for nu, l in ((0.1, 200), (0.5, 2000), (0.37, 101), (0.999, 10)):
        np.testing.assert_array_equal(oneclass_initial_alpha(nu, l), oneclass_alpha0(nu, l))
    with pytest.raises(UnboundLocalError):
        oneclass_initial_alpha(0.001, 100)


def test_estimator_signatures_match_reference_defaults():
    """Constructor kwargs and defaults of SURVEY 8(b) surface 1."""
    import inspect
    import pysvm_b200 as P
    sig = lambda c: {k: v.default for k, v in inspect.signature(c.__init__).parameters.items() if k != "self"}
    assert sig(P.KernelSVC) == dict(C=1., kernel='rbf', degree=3, gamma='scale', coef0=0., max_iter=1000, rff=False,
                                    D=1000, tol=1e-5, cache_size=256, multiclass='ovr', n_jobs=None)
    assert sig(P.NuSVC) == dict(nu=0.5, kernel='rbf', degree=3, gamma='scale', coef0=0., max_iter=1000, rff=False,
                                D=1000, tol=1e-5, cache_size=256, multiclass='ovr', n_jobs=None)
    assert sig(P.KernelSVR) == dict(C=1., eps=0., kernel='rbf', degree=3, gamma='scale', coef0=0., max_iter=1000,
                                    rff=False, D=1000, tol=1e-5, cache_size=256)
    assert sig(P.NuSVR) == dict(C=1., nu=0.5, kernel='rbf', degree=3, gamma='scale', coef0=0., max_iter=1000,
                                rff=False, D=1000, tol=1e-5, cache_size=256)
    assert sig(P.OneClassSVM) == dict(nu=0.5, kernel='rbf', degree=3, gamma='scale', coef0=0., max_iter=1000,
                                      rff=False, D=1000, tol=1e-5, cache_size=256)
    assert sig(P.LinearSVC) == dict(C=1., max_iter=1000, tol=1e-5, cache_size=256, multiclass='ovr', n_jobs=None)
    assert sig(P.LinearSVR) == dict(C=1., eps=0., max_iter=1000, tol=1e-5, cache_size=256)
    from sklearn.base import clone
    est = clone(P.KernelSVC(C=3., multiclass="ovo"))
    assert est.C == 3. and est.multiclass == "ovo"


def test_host_std_is_bit_identical_to_numpy():
    """gamma='scale' = 1 / (n_features * X.std()) (svc.py:212-217): the threaded host helper must reproduce numpy's
    pairwise float32 / float64 reduction to the last bit, and hand back a scalar of X's dtype."""
    from pysvm_b200 import _native as N
    from pysvm_b200.solver import host_std
    import ctypes as C
    lib = N.load()
    rng = np.random.default_rng(7)
    for shape, dt in [((7, 3), np.float32), ((129, 1), np.float32), ((1000, 37), np.float32), ((4096, 30), np.float64),
                      ((100001, 13), np.float32), ((70000, 17), np.float32), ((30000, 64), np.float64),
                      ((2100000, 1), np.float32)]:
        X = (rng.standard_normal(shape) * 3 + 1).astype(dt)
        out = C.c_double()
        N.check(lib.b2svm_host_std(X.ctypes.data, X.size, N.F64 if dt == np.float64 else N.F32, C.byref(out)))
        assert dt(out.value) == X.std(), (shape, dt)
        got = host_std(X)
        assert got == X.std() and type(got) is type(X.std())
    Xf = np.asfortranarray(rng.standard_normal((1500, 900)).astype(np.float32))       # not C-contiguous: numpy's own path
    assert host_std(Xf) == Xf.std()
    assert host_std(Xf[::2]) == Xf[::2].std()


def test_package_generators_equal_the_oracles():
    """bench.py's own arm draws its inputs from pysvm_b200.synth (the oracle is only the checker / the CPU baseline):
    the arrays must be the ones the fixtures and the oracle-side tests were made from."""
    from oracle import smo_oracle as O
    from pysvm_b200 import synth
    Xa, ya = synth.classification(n=3000, d=128, seed=0)
    Xb, yb = O.synth_classification(n=3000, d=128, seed=0)
    assert np.array_equal(Xa, Xb) and np.array_equal(ya, yb)
    Xa, za = synth.regression(n=2000, d=64, seed=0)
    Xb, zb = O.synth_regression(n=2000, d=64, seed=0)
    assert np.array_equal(Xa, Xb) and np.array_equal(za, zb)
    assert np.array_equal(synth.oneclass(n=1000, d=32, seed=1), O.synth_oneclass(n=1000, d=32, seed=1))


def test_host_std_every_small_length_and_block_edges():
    """The pairwise tree has special cases below 8 elements, at the 128-element block size and at every halving
    (sub-trees are rounded down to multiples of 8): sweep all short lengths and a few awkward long ones."""
    from pysvm_b200 import _native as N
    import ctypes as C
    lib = N.load()
    rng = np.random.default_rng(11)
    lengths = list(range(1, 300)) + [1023, 1024, 1025, 4097, 65535, 65537, 131071, 262145, 1000003]
    for dt, code in ((np.float32, N.F32), (np.float64, N.F64)):
        for n in lengths:
            x = (rng.standard_normal(n) * 2 - 0.5).astype(dt)
            out = C.c_double()
            N.check(lib.b2svm_host_std(x.ctypes.data, n, code, C.byref(out)))
            assert dt(out.value) == x.std(), (dt, n)


def test_polled_exchange_units_are_read_by_single_128_bit_loads():
    """The in-kernel exchange relies on {payload, tag} travelling in ONE 16-byte memory transaction.  ptxas narrows a
    vector load whose components are partly dead into scalar loads (this happened: a stale payload then passed the tag
    check once in ~10^4 iterations), so every polled unit keeps all four words alive.  Guard it in the SASS: no scalar
    or 64-bit STRONG.SYS global load may exist in the persistent kernels (the abort flag is a generic LD, not LDG)."""
    import shutil
    import subprocess
    from pysvm_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build()
    checked = 0
    for name in os.listdir(build.LIBDIR):
        if not (name.startswith("b2svm_launch_") and name.endswith(".o")):
            continue
        sass = subprocess.run([cuobjdump, "-sass", os.path.join(build.LIBDIR, name)], capture_output=True, text=True).stdout
        narrowed = [ln for ln in sass.splitlines() if "LDG.E.STRONG.SYS " in ln or "LDG.E.64.STRONG.SYS " in ln]
        assert not narrowed, (name, narrowed[:3])
        assert sass.count("LDG.E.128.STRONG.SYS") > 0 and sass.count("STG.E.128.STRONG.SYS") > 0, name
        checked += 1
    assert checked == 4
