"""Estimators with the reference's names (``pysvm/svm/__init__.py``), solved by the device SMO path.

Classification: :mod:`.svc`; regression (2l-variable "mirror" problems): :mod:`.svr`; one-class: :mod:`.oc_svm`.
The ``Bi*`` binary building blocks are exported too: ``bench.py`` and the parity tests drive them directly."""
from .oc_svm import OneClassSVM
from .svc import BiKernelSVC, BiLinearSVC, BiNuSVC, KernelSVC, LinearSVC, NuSVC
from .svr import KernelSVR, LinearSVR, NuSVR

__all__ = [
    "LinearSVC", "KernelSVC", "NuSVC", "BiLinearSVC", "BiKernelSVC", "BiNuSVC",
    "LinearSVR", "KernelSVR", "NuSVR",
    "OneClassSVM",
]
