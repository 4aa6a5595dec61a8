"""pysvm_b200 -- B200-native SMO solver behind PySVM's estimator / solver API.

Same public names as the reference package (``pysvm/__init__.py:3-6``)."""
from .svm import LinearSVC, LinearSVR, KernelSVC, KernelSVR, NuSVC, NuSVR, OneClassSVM

__all__ = ["LinearSVC", "LinearSVR", "KernelSVC", "KernelSVR", "NuSVC", "NuSVR", "OneClassSVM"]
