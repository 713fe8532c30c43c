"""the multivariate Gaussian draw (src/distributions.rs:484-540) is M matrix-vector products
`mean + L * z_m` in the reference; the host mirror stacks them into ONE product on the drop-in GEMM.  Checked
against the oracle's literal per-sample loop on the reference's own fixture (tests/distributions.rs:48-220) and
on seeded inputs, within the north star's bound tol * sum_k |L[n,k] z[m,k]|."""
import numpy as np
import pytest

import easy_ml_b200 as eml

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def bound_for(orc, want, mean, cov, dtype):
    """tol * (sum_k |L[n,k] z[m,k]| + |mean[n]|) per drawn element, z recovered from the oracle's samples."""
    low = orc.cholesky(cov).astype(np.float64)
    mean = mean.astype(np.float64)
    z = np.linalg.solve(low, (want.astype(np.float64) - mean).T).T
    return TOL[dtype] * (np.abs(z) @ np.abs(low).T + np.abs(mean)) + np.finfo(dtype).tiny


def test_reference_fixture_matrix_and_tensor_forms(orc, golden):
    g = golden["multivariate_gaussian"]
    mean = np.array(g["mean"])
    cov = np.array(g["covariance"])
    want = orc.multivariate_gaussian_draw(mean, cov, g["source"], g["max_samples"])
    gaussian = eml.MultivariateGaussian(eml.Matrix(mean.reshape(3, 1)), eml.Matrix(cov))
    samples = gaussian.draw(iter(g["source"]), g["max_samples"])          # a plain Python iterator, as in Rust
    assert samples.size() == (25, 3)
    assert np.all(np.abs(samples.data - want) <= bound_for(orc, want, mean, cov, np.float64))
    # the reference's assertions (:175-193)
    for f in range(3):
        feature_mean = samples.data[:, f].sum() / g["max_samples"]
        sd = np.sqrt(cov[f, f])
        assert mean[f] - 0.5 * sd < feature_mean < mean[f] + 0.5 * sd
    # Tensor form gives the same numbers (:196-219) and refuses equal dimension names (:497-499)
    tensor_gaussian = eml.MultivariateGaussianTensor(eml.Tensor([("means", 3)], mean),
                                                     eml.Tensor([("rows", 3), ("columns", 3)], cov))
    drawn = tensor_gaussian.draw(g["source"], g["max_samples"], "samples", "features")
    assert drawn.shape() == [("samples", 25), ("features", 3)]
    assert np.array_equal(drawn.data, samples.data)
    assert tensor_gaussian.draw(g["source"], g["max_samples"], "x", "x") is None
    # exhausted source / covariance that is not positive definite -> None (:346-368)
    assert gaussian.draw(g["source"][:99], g["max_samples"]) is None
    assert eml.MultivariateGaussian(eml.Matrix(mean.reshape(3, 1)),
                                    eml.Matrix(np.array([[1.0, 2, 0], [2, 1, 0], [0, 0, 1]]))).draw(g["source"], 2) is None


def test_constructor_panics():
    with pytest.raises(eml.Panic, match="Mean must be a column vector"):
        eml.MultivariateGaussian(eml.Matrix(np.zeros((1, 3))), eml.Matrix(np.eye(3)))
    with pytest.raises(eml.Panic, match="Supplied 'covariance' matrix is not square"):
        eml.MultivariateGaussian(eml.Matrix(np.zeros((3, 1))), eml.Matrix(np.zeros((3, 2))))
    with pytest.raises(eml.Panic, match="Means must be same length as covariance matrix"):
        eml.MultivariateGaussian(eml.Matrix(np.zeros((2, 1))), eml.Matrix(np.eye(3)))
    with pytest.raises(eml.MultivariateGaussianError, match="Covariance matrix is not square"):
        eml.MultivariateGaussianTensor(eml.Tensor([("m", 3)], np.zeros(3)), eml.Tensor([("r", 3), ("c", 2)], np.zeros((3, 2))))
    with pytest.raises(eml.MultivariateGaussianError, match="Mean vector has a different length"):
        eml.MultivariateGaussianTensor(eml.Tensor([("m", 2)], np.zeros(2)), eml.Tensor([("r", 3), ("c", 3)], np.eye(3)))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("samples,features", [(7, 5), (64, 33), (300, 96)])
def test_batched_draw_matches_per_sample_loop(orc, dtype, samples, features):
    r = np.random.default_rng(1000 + samples + features)
    x = r.standard_normal((4 * features, features))
    cov = (x.T @ x / (4 * features) + 0.1 * np.eye(features)).astype(dtype)   # symmetric positive definite
    mean = r.uniform(-3, 3, features).astype(dtype)
    per_row = features + (features & 1)
    source = r.uniform(1e-6, 1.0, samples * per_row).astype(dtype)
    want = orc.multivariate_gaussian_draw(mean, cov, source, samples)
    got = eml.MultivariateGaussian(eml.Matrix(mean.reshape(-1, 1)), eml.Matrix(cov)).draw(source, samples)
    assert np.all(np.abs(got.data.astype(np.float64) - want.astype(np.float64)) <= bound_for(orc, want, mean, cov, dtype))


def test_large_draw_is_one_tensor_core_product(orc):
    """4096 samples of a 512-dimensional Gaussian: one 4096 x 512 x 512 product (the reference would run 4096
    matrix-vector products); sampled rows against the oracle loop."""
    samples, features = 4096, 512
    r = np.random.default_rng(77)
    x = r.standard_normal((2 * features, features)).astype(np.float32)
    cov = (x.T @ x / (2 * features) + 0.5 * np.eye(features, dtype=np.float32)).astype(np.float32)
    mean = r.uniform(-1, 1, features).astype(np.float32)
    source = r.uniform(1e-6, 1.0, samples * features).astype(np.float32)
    before = eml.kernel_launches()
    got = eml.MultivariateGaussian(eml.Matrix(mean.reshape(-1, 1)), eml.Matrix(cov)).draw(source, samples)
    # (two split/pack passes, the product, at most one split-K reduction, the early-exit non-finite repair)
    assert eml.last_kernel() == "tf32x3_tcgen05" and eml.kernel_launches() - before <= 5
    assert got.size() == (samples, features)
    low = orc.cholesky(cov)
    rows = [0, 1, 2047, 4095]
    for m in rows:
        want = orc.multivariate_gaussian_draw(mean, cov, source[m * features:(m + 1) * features], 1)[0]
        z = eml.Gaussian(0, 1, np.float32).draw(source[m * features:(m + 1) * features], features)
        bound = 1e-5 * (np.abs(low).astype(np.float64) @ np.abs(z).astype(np.float64) + np.abs(mean))
        assert np.all(np.abs(got.data[m].astype(np.float64) - want.astype(np.float64)) <= bound)
    # the sample covariance of the draw approaches the requested covariance (statistical sanity, loose)
    centred = got.data - got.data.mean(axis=0)
    sample_cov = centred.T @ centred / samples
    assert np.abs(sample_cov - cov).max() < 0.25
