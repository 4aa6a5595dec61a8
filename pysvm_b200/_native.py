"""ctypes binding of libb2svm.so (include/b2svm.h).  There is no CPU fallback: if the CUDA
library is missing or fails, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libb2svm.so")
ABI_VERSION = 1
MAX_ITER = 1 << 29          # B2SVM_MAX_ITER: iterations per b2svm_solve call (29-bit epoch tag in the exchange records)

F32, F64 = 0, 1
KERNEL_IDS = {"linear": 0, "poly": 1, "rbf": 2, "sigmoid": 3, "precomputed": 4}
SOLVER_C, SOLVER_NU = 0, 1

EXPORTS = [
    "b2svm_abi_version", "b2svm_last_error", "b2svm_create", "b2svm_destroy", "b2svm_init_state",
    "b2svm_init_gradient", "b2svm_select", "b2svm_update", "b2svm_kernel_column", "b2svm_solve",
    "b2svm_rho", "b2svm_rho_b", "b2svm_kernel_matvec", "b2svm_rff_transform", "b2svm_rff_decision",
    "b2svm_solve_batched", "b2svm_mailbox_export", "b2svm_mailbox_attach", "b2svm_prepare", "b2svm_bias_partials", "b2svm_release_cache",
    "b2svm_host_std", "b2svm_solve_ranks_on_one_gpu", "b2svm_rff_transform_f32", "b2svm_kernel_matrix", "b2svm_weighted_rows", "b2svm_select_wss3",
]


class B2SVMError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b2svm error {code}: {message}")
        self.code = code


class KernelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("gamma", C.c_double), ("coef0", C.c_double),
                ("degree", C.c_double)]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p),
        ("n_rows", C.c_int64), ("n_features", C.c_int32), ("x_dtype", C.c_int32),
        ("X", C.c_void_p), ("ldx", C.c_int64), ("n_vars", C.c_int64),
        ("variant", C.c_int32), ("_pad0", C.c_int32),
        ("y", C.c_void_p), ("alpha", C.c_void_p), ("neg_y_grad", C.c_void_p),
        ("C", C.c_double), ("tol", C.c_double), ("kernel", KernelDesc),
        ("cache_bytes", C.c_uint64), ("max_ctas", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("_pad1", C.c_int32), ("row_offset", C.c_int64), ("n_rows_global", C.c_int64),
    ]


class SolveStats(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int64), ("converged", C.c_int32), ("n_ctas", C.c_int32),
        ("last_i", C.c_int64), ("last_j", C.c_int64),
        ("last_delta_i", C.c_double), ("last_delta_j", C.c_double),
        ("cold_sweeps", C.c_int64), ("cache_hits", C.c_int64), ("cache_misses", C.c_int64),
        ("algorithmic_bytes", C.c_double), ("device_ms", C.c_float), ("launches", C.c_int32),
        ("cache_stores", C.c_int64), ("cache_evictions", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """Load libb2svm.so and declare prototypes.  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pysvm_b200.build` (nvcc, sm_100a). "
            "pysvm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.b2svm_abi_version.restype = C.c_int
    lib.b2svm_last_error.restype = C.c_char_p
    lib.b2svm_create.argtypes = [P(ProblemDesc), P(vp)]
    lib.b2svm_destroy.argtypes = [vp]
    lib.b2svm_init_state.argtypes = [vp, vp]
    lib.b2svm_init_gradient.argtypes = [vp, vp]
    lib.b2svm_select.argtypes = [vp, P(i64), P(i64)]
    lib.b2svm_select_wss3.argtypes = [vp, P(i64), P(i64)]
    lib.b2svm_update.argtypes = [vp, i64, i64, P(dbl), P(dbl)]
    lib.b2svm_kernel_column.argtypes = [vp, i64, vp]
    lib.b2svm_solve.argtypes = [vp, i64, P(SolveStats)]
    lib.b2svm_rho.argtypes = [vp, P(dbl)]
    lib.b2svm_rho_b.argtypes = [vp, P(dbl), P(dbl)]
    lib.b2svm_kernel_matvec.argtypes = [i32, vp, P(KernelDesc), i32, i32, vp, i64, i64, vp, i64, i64, vp, vp]
    lib.b2svm_rff_transform.argtypes = [i32, vp, i32, vp, i64, i64, i32, vp, vp, i32, vp]
    lib.b2svm_rff_transform_f32.argtypes = [i32, vp, vp, i64, i64, i32, vp, vp, i32, vp, i64]
    lib.b2svm_kernel_matrix.argtypes = [i32, vp, P(KernelDesc), i32, i32, vp, i64, i64, vp, i64, i64, vp, i64]
    lib.b2svm_weighted_rows.argtypes = [i32, vp, i32, vp, i64, i64, i32, vp, vp]
    lib.b2svm_rff_decision.argtypes = [i32, vp, i32, vp, i64, i64, i32, vp, vp, i32, vp, vp]
    lib.b2svm_solve_batched.argtypes = [P(vp), i32, i64, P(SolveStats)]
    lib.b2svm_solve_ranks_on_one_gpu.argtypes = [P(vp), i32, i64, P(SolveStats)]
    lib.b2svm_mailbox_export.argtypes = [vp, vp]
    lib.b2svm_mailbox_attach.argtypes = [vp, vp]
    lib.b2svm_prepare.argtypes = [vp]
    lib.b2svm_bias_partials.argtypes = [vp, P(dbl)]
    lib.b2svm_release_cache.argtypes = [vp]
    lib.b2svm_host_std.argtypes = [vp, i64, i32, P(dbl)]
    for name in EXPORTS:
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        if name != "b2svm_last_error":
            fn.restype = C.c_int
    if lib.b2svm_abi_version() != ABI_VERSION:
        raise ImportError("libb2svm.so ABI version mismatch; rebuild with `python -m pysvm_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise B2SVMError(rc, load().b2svm_last_error().decode(errors="replace"))
