"""eml_gemm_*_sharded: one call from one host thread, HOST buffers, the GPUs of this box (BASELINE config 5 behind the
drop-in boundary; replaces matrix_view_multiplication, reference src/matrices/operations.rs:165-196, for large operands).

Rows of C are independent units, so there is no cross-GPU reduction and the sharded result must carry the BITS of the
single-GPU entry point -- for every GPU count the box offers, pageable and page-locked caller memory, ragged shapes
(slabs that are partly or wholly empty, leading dimensions larger than the rows)."""
import ctypes as C

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import BadArgError, check, lib

pytestmark = pytest.mark.gpu


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gemm(a, b, n_gpus=None, lda=None, ldb=None):
    sfx = "f32" if a.dtype == np.float32 else "f64"
    m, k, n = a.shape[0], a.shape[1] if lda is None else None, b.shape[1]
    c = np.empty((a.shape[0], b.shape[1]), a.dtype)
    args = (ptr(a), ptr(b), ptr(c), a.shape[0], b.shape[1], a.shape[1], a.strides[0] // a.itemsize,
            b.strides[0] // b.itemsize, b.shape[1])
    if n_gpus is None:
        check(getattr(lib, f"eml_gemm_{sfx}")(*args))
    else:
        check(getattr(lib, f"eml_gemm_{sfx}_sharded")(n_gpus, *args))
    return c


def gpu_counts():
    n = eml.device_count()
    return [g for g in (1, 2, 3, 4, 8) if g <= n]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(2304, 1536, 2048), (700, 300, 1000), (100, 64, 40), (4096, 8192, 512)])
def test_sharded_host_gemm_has_the_bits_of_one_gpu(dtype, shape):
    m, k, n = shape
    r = np.random.default_rng(m + k + n)
    a = r.uniform(-1, 1, (m, k)).astype(dtype)
    b = r.uniform(-1, 1, (k, n)).astype(dtype)
    want = gemm(a, b)
    bound = (1e-5 if dtype == np.float32 else 1e-12) * (np.abs(a[:3]).astype(np.float64) @ np.abs(b).astype(np.float64))
    assert np.all(np.abs(want[:3].astype(np.float64) - a[:3].astype(np.float64) @ b.astype(np.float64)) <= bound)
    for g in gpu_counts():
        got = gemm(a, b, n_gpus=g)
        if dtype == np.float64:
            assert np.array_equal(got, want), f"{g} GPUs: bits differ from the single-GPU call"
        else:
            # the f32 tensor-core kernel chooses its deterministic split-K from the number of tiles of a launch, which a
            # row slab changes: both within the contract
            full = (1e-5) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
            assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= 2 * full + 1e-30)
        assert np.array_equal(gemm(a, b, n_gpus=g), got), f"{g} GPUs: not bit-identical run to run"


def test_sharded_host_gemm_views_with_leading_dimensions_and_pinned_memory():
    r = np.random.default_rng(5)
    m, k, n = 1500, 1100, 900
    big_a = r.uniform(-1, 1, (m, k + 24))
    big_b = r.uniform(-1, 1, (k, n + 8))
    a, b = big_a[:, :k], big_b[:, :n]            # row-major views: lda = k + 24, ldb = n + 8
    want = np.ascontiguousarray(a) @ np.ascontiguousarray(b)
    hp = C.c_void_p()
    check(lib.eml_host_alloc(C.byref(hp), m * n * 8))
    try:
        pinned_c = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(m * n,)).reshape(m, n)
        for g in gpu_counts():
            pinned_c[:] = 0
            check(lib.eml_gemm_f64_sharded(g, ptr(a), ptr(b), hp, m, n, k, k + 24, n + 8, n))
            bound = 1e-12 * (np.abs(a) @ np.abs(b))
            assert np.all(np.abs(pinned_c - want) <= bound)
    finally:
        del pinned_c
        check(lib.eml_host_free(hp))


def test_sharded_host_gemm_refuses_more_gpus_than_the_box_has():
    a = np.ones((8, 8))
    c = np.empty((8, 8))
    with pytest.raises(BadArgError):
        check(lib.eml_gemm_f64_sharded(eml.device_count() + 1, ptr(a), ptr(a), ptr(c), 8, 8, 8, 8, 8, 8))
    with pytest.raises(BadArgError):
        check(lib.eml_gemm_f64_sharded(0, ptr(a), ptr(a), ptr(c), 8, 8, 8, 8, 8, 8))
