"""Host-side logic of the Python mirror that needs no GPU: the restated Cholesky factorisation and Box-Muller
draws of easy-ml_b200/distributions.py against the oracle's literal loops, on the reference's fixtures
(tests/linear_algebra.rs:175-240, tests/distributions.rs:10-46,48-220), and the no-device behaviour of the product
path (it must refuse, not fall back)."""
import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200 import distributions


def test_cholesky_restatement_matches_oracle_and_goldens(orc, golden):
    for g in golden["cholesky"]:
        dtype = np.float32 if g["exact"] else np.float64
        a = np.array(g["a"], dtype=dtype)
        low = distributions.cholesky_decomposition(eml.Matrix(a))
        assert np.array_equal(low.data, orc.cholesky(a)), g["cite"]
        if g["exact"]:
            assert np.array_equal(low.data, np.array(g["expect"], dtype=dtype))
    r = np.random.default_rng(4)
    for n, dtype in ((17, np.float32), (40, np.float64)):
        x = r.standard_normal((3 * n, n))
        spd = (x.T @ x / (3 * n) + 0.1 * np.eye(n)).astype(dtype)
        assert np.array_equal(distributions.cholesky_decomposition(spd), orc.cholesky(spd))
    assert distributions.cholesky_decomposition(eml.Matrix(np.array([[1.0, 2.0], [2.0, 1.0]]))) is None
    # (a - sum) * (1 / L[j][j]), two roundings (src/linear_algebra.rs:1093-1096): 5 * fl(1/3) != fl(5/3)
    low = distributions.cholesky_decomposition(np.array([[9.0, 5.0], [5.0, 10.0]]))
    assert low[1, 0].hex() == float.fromhex("0x1.aaaaaaaaaaaaap+0").hex() and low[1, 0] != 5.0 / 3.0
    assert distributions.cholesky_decomposition(np.zeros((2, 3))) is None


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_box_muller_restatement_consumes_the_source_like_the_reference(orc, golden, dtype):
    source = golden["multivariate_gaussian"]["source"]
    # Gaussian::draw, tests/distributions.rs:10-46 style: even and odd sample counts, exhausted source
    for count in (1, 2, 5, 12):
        got = eml.Gaussian(1.0, 2.0, dtype).draw(iter(source), count)
        want = orc.gaussian_draw(1.0, 2.0, iter(source), count, dtype)
        assert np.array_equal(got, want)
    assert eml.Gaussian(0, 1, dtype).draw(iter(source[:5]), 6) is None
    # row-wise draws of the multivariate form: N odd -> N + 1 numbers per row, the surplus sample dropped
    rows = distributions.Gaussian(0, 1, dtype)._draw_rows(distributions._Source(source, dtype), 25, 3)
    it = iter(source)
    want = np.array([orc.gaussian_draw(0, 1, it, 3, dtype) for _ in range(25)])
    assert np.array_equal(rows, want)
    assert distributions.Gaussian(0, 1, dtype)._draw_rows(distributions._Source(source[:99], dtype), 25, 3) is None


def test_product_path_refuses_to_run_without_a_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(eml.NoDeviceError):
        eml.Matrix([[1.0, 2.0]]) * eml.Matrix([[3.0], [4.0]])
    with pytest.raises(eml.NoDeviceError):
        eml.MultivariateGaussian(eml.Matrix([[0.0], [0.0]]), eml.Matrix(np.eye(2))).draw([0.5] * 8, 2)
    with pytest.raises(eml.NoDeviceError):
        eml.DeviceMatrix.from_host(eml.Matrix([[1.0]]))
