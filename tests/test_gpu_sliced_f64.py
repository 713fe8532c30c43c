"""f64 GEMM through error-free int8 slices on the INT8 tensor cores (eml_gemm_f64_sliced_dev, csrc/gemm_sliced_f64.cu).

The contract of the path is |C - ref| <= 1e-12 sum_k |a_ik b_kj| per element.  The sliced kernel's truncation is relative
to row / column maxima, so a guard pass must prove the contract for the operands at hand or hand the call to the DMMA
kernel.  Checked here: the contract on ordinary data (against an extended-precision host product and against the
oracle), exactness on integer data, and that every way of defeating the scaling is caught by the guard."""
import ctypes as C

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

vp = lambda t: C.c_void_p(t.data_ptr())


def sliced(a, b, slices=0):
    m, k = a.shape
    n = b.shape[1]
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float64)
    path = C.c_int(-1)
    check(lib.eml_gemm_f64_sliced_dev(vp(a), vp(b), vp(c), m, n, k, a.stride(0), b.stride(0), c.stride(0), slices,
                                      C.byref(path), None))
    torch.cuda.synchronize()
    return c, path.value


def worst_ratio(c, a, b, rows):
    """max over the sampled rows of |c - exact| / (1e-12 sum_k |a b|), exact in long double on the host"""
    ah = a[rows].cpu().numpy().astype(np.longdouble)
    bh = b.cpu().numpy().astype(np.longdouble)
    err = np.abs(c[rows].cpu().numpy().astype(np.longdouble) - ah @ bh)
    return float((err / (1e-12 * (np.abs(ah) @ np.abs(bh)) + 1e-300)).max())


def uniform(g, *shape):
    return torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) * 2 - 1


@pytest.mark.parametrize("shape", [(512, 768, 640), (1000, 1300, 900), (257, 300, 2050), (2048, 2048, 2048)])
def test_contract_on_ordinary_data(orc, shape):
    m, n, k = shape
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a, b = uniform(g, m, k), uniform(g, k, n)
    c, path = sliced(a, b, 8)
    assert path == 8 and eml.last_kernel() == "sliced_int8_f64"
    rows = [0, 1, m // 2, m - 1]
    assert worst_ratio(c, a, b, rows) < 1e-2      # two orders of magnitude inside the contract
    # and against the oracle's reference-order product on one row
    ref = orc.matrix_mul(a[m - 1:m].cpu().numpy(), b.cpu().numpy(), threads=4)
    bound = 1e-12 * (np.abs(a[m - 1:m].cpu().numpy()) @ np.abs(b.cpu().numpy()))
    assert np.all(np.abs(c[m - 1:m].cpu().numpy() - ref) <= bound)
    # deterministic
    assert torch.equal(c, sliced(a, b, 8)[0])
    # slices = 0: as few as the guard can prove -- seven for data like this, still far inside the contract
    c7, auto = sliced(a, b)
    assert auto == 7 and worst_ratio(c7, a, b, rows) < 1e-1


def test_rows_and_columns_of_very_different_magnitude_are_scaled_independently():
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(1024, 1024, generator=g, device="cuda", dtype=torch.float64) * \
        torch.logspace(-120, 120, 1024, device="cuda", dtype=torch.float64)[:, None]
    b = uniform(g, 1024, 1024) * torch.logspace(60, -60, 1024, device="cuda", dtype=torch.float64)[None, :]
    c, path = sliced(a, b, 8)
    assert path == 8
    assert worst_ratio(c, a, b, [0, 500, 1023]) < 1e-2


def test_integer_valued_operands_are_exact():
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randint(-8, 9, (512, 640), generator=g, device="cuda").double()
    b = torch.randint(-8, 9, (640, 768), generator=g, device="cuda").double()
    c, path = sliced(a, b)
    assert path >= 7 and torch.equal(c, a @ b)
    a[3] = 0.0          # an all-zero row and column are exact and exempt from the guard
    b[:, 5] = 0.0
    c, path = sliced(a, b)
    assert path >= 7 and torch.equal(c, a @ b)


def test_guard_hands_unprovable_operands_to_the_dmma_kernel():
    g = torch.Generator(device="cuda").manual_seed(4)
    # column 0 of B only meets entries of A that are 1e-30 of their row maximum: the slices cannot represent them
    a = uniform(g, 512, 512)
    b = torch.zeros(512, 512, device="cuda", dtype=torch.float64)
    a[:, 0] = 1e-30
    b[0, :] = 1.0
    b[1:, 1:] = uniform(g, 511, 511)
    c, path = sliced(a, b)
    assert path == 0 and eml.last_kernel().startswith("dmma_f64")
    assert worst_ratio(c, a, b, [0, 255, 511]) < 1.0
    # non-finite values, extreme exponents, too few slices for the length of the sum, shapes below the kernel's range
    a = uniform(g, 512, 512)
    a[5, 7] = float("inf")
    assert sliced(a, uniform(g, 512, 512))[1] == 0
    a = uniform(g, 512, 512)
    a[9, 9] = float("nan")
    assert sliced(a, uniform(g, 512, 512))[1] == 0
    assert sliced(uniform(g, 512, 512) * 1e290, uniform(g, 512, 512))[1] == 0
    assert sliced(uniform(g, 512, 512), uniform(g, 512, 512), slices=6)[1] == 0
    c, path = sliced(uniform(g, 128, 512), uniform(g, 512, 512))
    assert path == 0 and not bool(torch.isnan(c).any())
    rc = lib.eml_gemm_f64_sliced_dev(vp(a), vp(a), vp(a), 512, 512, 512, 512, 512, 512, 11, None, None)
    assert rc != 0 and "6..10 slices" in lib.eml_last_error().decode()


def test_more_slices_never_hurt_and_the_default_beats_the_fp64_kernel():
    g = torch.Generator(device="cuda").manual_seed(6)
    a, b = uniform(g, 1024, 1024), uniform(g, 1024, 1024)
    rows = [0, 400, 1023]
    ratios = {}
    for s in (7, 8, 9, 10):
        c, path = sliced(a, b, s)
        assert path == s
        ratios[s] = worst_ratio(c, a, b, rows)
    assert ratios[7] < 1e-1 and ratios[8] < 1e-3 and ratios[9] < 1e-3 and ratios[10] < 1e-3
    cd = torch.empty_like(a)
    check(lib.eml_gemm_f64_dev(vp(a), vp(b), vp(cd), 1024, 1024, 1024, 1024, 1024, 1024, None))
    torch.cuda.synchronize()
    assert ratios[8] < worst_ratio(cd, a, b, rows)   # 55-bit slices + exact integer sums: tighter than FMA chains


def test_device_matrix_front_end():
    r = np.random.default_rng(12)
    a, b = r.uniform(-1, 1, (600, 700)), r.uniform(-1, 1, (700, 520))
    da, db = eml.DeviceMatrix.from_host(eml.Matrix(a)), eml.DeviceMatrix.from_host(eml.Matrix(b))
    prod, used = da.mul_sliced(db)
    assert used and prod.size() == (600, 520)
    exact = a.astype(np.longdouble) @ b.astype(np.longdouble)
    bound = 1e-12 * (np.abs(a).astype(np.longdouble) @ np.abs(b).astype(np.longdouble))
    assert np.all(np.abs(prod.to_host().data.astype(np.longdouble) - exact) <= 1e-2 * bound)
    assert np.allclose(prod.to_host().data, (da * db).to_host().data, rtol=0, atol=1e-12)
    with pytest.raises(eml.Panic, match="Mismatched Matrices, left is 600x700, right is 600x700"):
        da.mul_sliced(da)


@pytest.mark.parametrize("decades", [1, 3, 5, 7, 9, 12, 16])
@pytest.mark.parametrize("sparsity", [0.0, 0.6])
def test_guard_is_sound_on_wide_dynamic_ranges(decades, sparsity):
    """Magnitudes log-uniform over `decades` decades inside every row / column, optionally with most entries zero: the
    wider the spread, the less the row-maximum scaling can represent.  Whatever the guard decides, EVERY element of the
    result must meet the contract (checked against a long-double product of the whole matrix); and it must not refuse
    the benign cases."""
    m, n, k = 320, 288, 512
    r = np.random.default_rng(1000 * decades + int(10 * sparsity))
    def draw(shape):
        mag = 10.0 ** (-decades * r.random(shape))
        x = mag * r.choice([-1.0, 1.0], shape)
        if sparsity:
            x[r.random(shape) < sparsity] = 0.0
        return x
    a, b = draw((m, k)), draw((k, n))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    exact = a.astype(np.longdouble) @ b.astype(np.longdouble)
    bound = 1e-12 * (np.abs(a).astype(np.longdouble) @ np.abs(b).astype(np.longdouble))
    for slices in (0, 8, 10):
        c, path = sliced(ta, tb, slices)
        err = np.abs(c.cpu().numpy().astype(np.longdouble) - exact)
        assert np.all(err <= bound), (decades, sparsity, slices, path, float((err / (bound + 1e-300)).max()))
        if decades <= 3 and sparsity == 0.0:
            assert path >= 7, "a benign spread must not be refused"
