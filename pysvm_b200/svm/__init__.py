from .svc import LinearSVC, KernelSVC, NuSVC, BiLinearSVC, BiKernelSVC, BiNuSVC
from .svr import LinearSVR, KernelSVR, NuSVR
from .oc_svm import OneClassSVM
