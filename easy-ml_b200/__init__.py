"""easy-ml_b200: B200-native dense-contraction path for Easy ML (f32/f64), over the C ABI of
include/easyml_b200.h.  Import as `importlib.import_module("easy-ml_b200")` or via the root shim
`import easy_ml_b200`.

Layout: csrc/ (CUDA kernels + the C ABI), lib/ (built libeasyml_b200.so), _ffi.py (ctypes binding),
host.py + distributions.py (Python mirror of the reference operator API and of the GEMM callers), host/ (C++ mirror), rust/ (the `cuda` feature shim).
"""
from . import _ffi  # raises ImportError when the CUDA library is not built: there is no fallback
from ._ffi import EmlError, NoDeviceError, BadArgError, StateError, OPS, lib
from .host import (Derivatives, Einsum, InconsistentDimensionLengthError, Matrix, Panic, PinnedBuffer, Record, RecordMatrix,
                   RecordTensor, RecordedStep, Tensor, WengertList, sgd_update, wait_arrival)
from .device import DeviceMatrix, DeviceTensor
from .distributions import (Gaussian, MultivariateGaussian, MultivariateGaussianError, MultivariateGaussianTensor,
                            cholesky_decomposition)


def device_count():
    """eml_init: number of CUDA devices; raises NoDeviceError when there is none (no CPU fallback)."""
    import ctypes

    n = ctypes.c_int()
    _ffi.check(lib.eml_init(ctypes.byref(n)))
    return n.value


def kernel_launches():
    return int(lib.eml_kernel_launches())


def last_kernel():
    return lib.eml_last_kernel().decode()
