"""ctypes binding of libeasyml_b200.so -- the C ABI declared in include/easyml_b200.h.

There is no fallback of any kind: if the shared library is missing this module raises at import,
and every compute call returns EML_ERR_NO_DEVICE (raised as NoDeviceError) without a B200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeasyml_b200.so")

EML_OK, EML_ERR_NO_DEVICE, EML_ERR_BAD_ARG, EML_ERR_OOM, EML_ERR_CUDA, EML_ERR_NCCL, EML_ERR_STATE = (
    0, -1, -2, -3, -4, -5, -6)
EML_F32, EML_F64 = 0, 1

OPS = dict(NEG=0, SIN=1, COS=2, EXP=3, LN=4, SQRT=5, ADD_C=6, SUB_C=7, MUL_C=8, DIV_C=9, POW_C=10, C_SUB=11,
           C_DIV=12, C_POW=13, ADD=32, SUB=33, MUL=34, DIV=35, POW=36)


class EmlError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[eml {code}] {message}")
        self.code, self.message = code, message


class NoDeviceError(EmlError):
    pass


class BadArgError(EmlError):
    pass


class StateError(EmlError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C easy-ml_b200/csrc`). There is no CPU fallback for the f32/f64 path.")

lib = C.CDLL(LIB_PATH)

i64, vp, cint = C.c_int64, C.c_void_p, C.c_int
pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "eml_init": (cint, [C.POINTER(cint)]),
    "eml_shutdown": (None, []),
    "eml_last_error": (C.c_char_p, []),
    "eml_version": (cint, []),
    "eml_kernel_launches": (i64, []),
    "eml_last_kernel": (C.c_char_p, []),
    "eml_alloc": (cint, [C.POINTER(vp), C.c_size_t]),
    "eml_free": (cint, [vp]),
    "eml_alloc_async": (cint, [C.POINTER(vp), C.c_size_t, vp]),
    "eml_free_async": (cint, [vp, C.c_size_t, vp]),
    "eml_upload": (cint, [vp, vp, C.c_size_t]),
    "eml_download": (cint, [vp, vp, C.c_size_t]),
    "eml_sync": (cint, []),
    "eml_gemm_f64_sharded": (cint, [cint, vp, vp, vp, i64, i64, i64, i64, i64, i64]),
    "eml_gemm_f32_sharded": (cint, [cint, vp, vp, vp, i64, i64, i64, i64, i64, i64]),
    "eml_gemm_f64_scatter_dev": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64, cint, C.POINTER(vp), cint, vp]),
    "eml_gemm_f64_sliced_dev": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64, cint, C.POINTER(cint), vp]),
    "eml_ipc_export": (cint, [vp, vp]),
    "eml_ipc_open": (cint, [vp, C.POINTER(vp)]),
    "eml_ipc_close": (cint, [vp]),
    "eml_copy_async": (cint, [vp, vp, C.c_size_t, vp]),
    "eml_einsum_n_f32": (cint, [cint, C.POINTER(vp), pi64, vp, cint, pi64, pi64, cint, pi64, pi64]),
    "eml_einsum_n_f64": (cint, [cint, C.POINTER(vp), pi64, vp, cint, pi64, pi64, cint, pi64, pi64]),
    "eml_covariance_f32": (cint, [vp, i64, i64, i64, cint, vp]),
    "eml_covariance_f64": (cint, [vp, i64, i64, i64, cint, vp]),
    "eml_covariance_f32_dev": (cint, [vp, i64, i64, cint, vp, vp]),
    "eml_covariance_f64_dev": (cint, [vp, i64, i64, cint, vp, vp]),
    "eml_measure_fp64_tensor_peak": (cint, [C.POINTER(C.c_double)]),
    "eml_host_alloc": (cint, [C.POINTER(vp), C.c_size_t]),
    "eml_host_free": (cint, [vp]),
    "eml_tape_create": (cint, [cint, C.POINTER(vp)]),
    "eml_tape_destroy": (cint, [vp]),
    "eml_tape_len": (cint, [vp, pi64]),
    "eml_tape_clear": (cint, [vp]),
    "eml_tape_variables": (cint, [vp, vp, i64, i64, pi64]),
    "eml_tape_constants": (cint, [vp, vp, i64, i64, pi64]),
    "eml_tape_variables_dev": (cint, [vp, vp, i64, i64, pi64]),
    "eml_tape_constants_dev": (cint, [vp, vp, i64, i64, pi64]),
    "eml_tape_from_existing": (cint, [vp, vp, vp, i64, i64, cint, pi64]),
    "eml_tape_range": (cint, [vp, i64, i64, i64, i64, i64, cint, pi64]),
    "eml_tape_append_nullary": (cint, [vp, pi64]),
    "eml_tape_append_unary": (cint, [vp, i64, C.c_double, pi64]),
    "eml_tape_append_binary": (cint, [vp, i64, C.c_double, i64, C.c_double, pi64]),
    "eml_tape_derivatives_at": (cint, [vp, i64, C.POINTER(vp)]),
    "eml_tape_reset": (cint, [vp, i64]),
    "eml_val_release": (cint, [vp, i64]),
    "eml_val_shape": (cint, [vp, i64, pi64, pi64, C.POINTER(cint)]),
    "eml_val_numbers": (cint, [vp, i64, vp]),
    "eml_val_numbers_async": (cint, [vp, i64, vp, C.POINTER(vp)]),
    "eml_arrival_wait": (cint, [vp]),
    "eml_tape_sync": (cint, [vp]),
    "eml_tape_capture_begin": (cint, [vp]),
    "eml_tape_capture_end": (cint, [vp, C.POINTER(vp)]),
    "eml_graph_launch": (cint, [vp, C.POINTER(vp)]),
    "eml_graph_destroy": (cint, [vp]),
    "eml_graph_profile": (cint, [vp, cint, C.c_char_p, i64]),
    "eml_val_indexes": (cint, [vp, i64, vp]),
    "eml_val_device_ptr": (cint, [vp, i64, C.POINTER(vp)]),
    "eml_tape_matmul": (cint, [vp, i64, i64, pi64]),
    "eml_tape_unary": (cint, [vp, cint, C.c_double, i64, pi64]),
    "eml_tape_binary": (cint, [vp, cint, i64, i64, pi64]),
    "eml_tape_derivatives": (cint, [vp, i64, i64, i64, C.POINTER(vp)]),
    "eml_derivs_free": (cint, [vp]),
    "eml_derivs_len": (cint, [vp, pi64]),
    "eml_derivs_at": (cint, [vp, i64, C.POINTER(C.c_double)]),
    "eml_derivs_at_val": (cint, [vp, i64, vp]),
    "eml_derivs_at_val_dev": (cint, [vp, i64, vp]),
    "eml_derivs_to_host": (cint, [vp, vp]),
    "eml_tape_sgd_update": (cint, [vp, i64, vp, C.c_double]),
}
for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
    SIGNATURES.update({
        f"eml_gemm_{sfx}": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64]),
        f"eml_gemm_exact_{sfx}": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64]),
        f"eml_contract_{sfx}": (cint, [vp, vp, vp, i64, i64, i64, i64, pi64, pi64, pi64]),
        f"eml_einsum2_{sfx}": (cint, [vp, i64, vp, i64, vp, cint, pi64, pi64, pi64, cint, pi64, pi64, pi64]),
        f"eml_dot_{sfx}": (cint, [vp, i64, vp, i64, i64, vp]),
        f"eml_contract_{sfx}_dev": (cint, [vp, vp, vp, i64, i64, i64, i64, pi64, pi64, pi64, cint, vp]),
        f"eml_gemm_{sfx}_dev": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64, vp]),
        f"eml_gemm_exact_{sfx}_dev": (cint, [vp, vp, vp, i64, i64, i64, i64, i64, i64, vp]),
        f"eml_matmul_bwd_{sfx}_dev": (cint, [vp, vp, vp, vp, vp, i64, i64, i64, vp]),
        f"eml_unary_{sfx}_dev": (cint, [cint, ct, vp, vp, vp, i64, vp]),
        f"eml_binary_{sfx}_dev": (cint, [cint, vp, vp, vp, vp, vp, i64, vp]),
        f"eml_chain_{sfx}_dev": (cint, [vp, vp, vp, i64, vp]),
        f"eml_sgd_{sfx}_dev": (cint, [vp, vp, ct, i64, vp]),
        f"eml_reorder_{sfx}_dev": (cint, [vp, vp, cint, pi64, pi64, vp]),
    })

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here == the library does not export what the header declares
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    return lib.eml_last_error().decode(errors="replace")


def check(rc):
    if rc == EML_OK:
        return
    msg = last_error()
    if rc == EML_ERR_NO_DEVICE:
        raise NoDeviceError(rc, msg)
    if rc == EML_ERR_BAD_ARG:
        raise BadArgError(rc, msg)
    if rc == EML_ERR_STATE:
        raise StateError(rc, msg)
    raise EmlError(rc, msg)


def arr3(v):
    return (C.c_int64 * 3)(*[int(x) for x in v])


def arrn(v):
    v = [int(x) for x in v]
    return (C.c_int64 * max(len(v), 1))(*v)
