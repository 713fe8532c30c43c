// host_test.cpp -- the reference's own known-answer tests for the hot path, replayed through the C++ mirror
// (easyml.hpp) and the C ABI.  Exit code 0 = all passed.  Without a GPU it checks that the product path fails
// loudly (BackendError: no CUDA device), i.e. that there is no CPU fallback, and exits 3.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>

#include "easyml.hpp"

using namespace easyml;

static int failures = 0;
#define EXPECT(cond)                                                            \
    do {                                                                        \
        if (!(cond)) {                                                          \
            std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                         \
        }                                                                       \
    } while (0)

template <class F>
static std::string panic_text(F&& f) {
    try {
        f();
    } catch (const Panic& p) {
        return p.what();
    }
    return "<no panic>";
}

template <class T>
static void matrix_goldens() {
    // tests/matrices.rs:161-166
    Matrix<T> a(3, 2, {1, 2, 3, 4, 5, 6}), b(2, 3, {1, 2, 3, 4, 5, 6});
    EXPECT((a * b) == Matrix<T>(3, 3, {9, 12, 15, 19, 26, 33, 29, 40, 51}));
    // tests/matrices.rs:168-173 (#[should_panic]); message src/matrices/operations.rs:176-183
    EXPECT(panic_text([&] { (void)(a * a); }) ==
           "Mismatched Matrices, left is 3x2, right is 3x2, * is only defined for MxN * NxL");
    // tests/linear_algebra.rs:157-170: covariance of column features (exact: the data is +-1)
    Matrix<T> feat(4, 3, {1, 1, -1, -1, -1, 1, -1, -1, 1, 1, 1, -1});
    EXPECT(feat.covariance_column_features() == Matrix<T>(3, 3, {1, 1, -1, 1, 1, -1, -1, -1, 1}));
    // tests/linear_algebra.rs:183-190: L * L^T == A exactly
    Matrix<T> l(3, 3, {5, 0, 0, 3, 3, 0, -1, 1, 3}), lt(3, 3, {5, 3, -1, 0, 3, 1, 0, 0, 3});
    EXPECT((l * lt) == Matrix<T>(3, 3, {25, 15, -5, 15, 18, 0, -5, 0, 11}));
}

template <class T>
static void tensor_goldens() {
    // src/tensors/operations.rs:521-546
    Tensor<T> a({{"r", 2}, {"c", 3}}, {1, 2, 3, 4, 5, 6}), b({{"r", 3}, {"c", 2}}, {10, 11, 12, 13, 14, 15});
    EXPECT((a * b) == Tensor<T>({{"r", 2}, {"c", 2}}, {76, 82, 184, 199}));
    // :1655-1708: output takes the left's first and the right's second dimension name
    Tensor<T> c({{"a", 3}, {"b", 2}}, {1, 2, 3, 4, 5, 6});
    EXPECT((a * c) == Tensor<T>({{"r", 2}, {"b", 2}}, {22, 28, 49, 64}));
    // :485-491 and :14-18,492-501
    Tensor<T> bad({{"a", 2}, {"b", 3}}, {1, 2, 3, 4, 5, 6});
    EXPECT(panic_text([&] { (void)(a * bad); }) ==
           "Mismatched tensors, left is [(\"r\", 2), (\"c\", 3)], right is [(\"a\", 2), (\"b\", 3)], * is only defined "
           "for MxN * NxL dimension lengths");
    Tensor<T> rc({{"rows", 2}, {"columns", 3}}, {1, 2, 3, 4, 5, 6}), cr({{"columns", 3}, {"rows", 2}}, {1, 2, 3, 4, 5, 6});
    EXPECT(panic_text([&] { (void)(rc * cr); }).rfind("Matrix multiplication of tensors with shapes left "
                                                       "[(\"rows\", 2), (\"columns\", 3)] and right [(\"columns\", 3), "
                                                       "(\"rows\", 2)] would create duplicate dimension names", 0) == 0);
    // Tensor<T,1>::scalar_product
    Tensor<T> u({{"x", 3}}, {1, 2, 3}), v({{"x", 3}}, {4, 5, 6});
    EXPECT(u.scalar_product(v) == T(32));
}

template <class T>
static void einsum_goldens() {
    // tests/einsum.rs:165-187 (ij,jk->ijk) and :189-206 (ij,kl->il): not GEMMs, bit-exact generic kernel
    Tensor<T> x({{"i", 2}, {"j", 3}}, {0, 10, 20, 1, 11, 21});
    Tensor<T> y({{"j", 3}, {"k", 4}}, {0, 10, 20, 30, 1, 11, 21, 31, 2, 12, 22, 32});
    auto ijk = Einsum::with_2(x, y).to({"i", "j", "k"});
    EXPECT(ijk.shape() == (Shape{{"i", 2}, {"j", 3}, {"k", 4}}));
    EXPECT(ijk.data()[4] == T(10) && ijk.data()[5] == T(110) && ijk.data()[11] == T(640));
    Tensor<T> y2({{"k", 3}, {"l", 4}}, y.data());
    EXPECT(Einsum::with_2(x, y2).to({"i", "l"}) == Tensor<T>({{"i", 2}, {"l", 4}}, {90, 990, 1890, 2790, 99, 1089, 2079, 3069}));
    // src/tensors/einsum.rs:58-72 doctest: ab,bc->ac == w * x
    Tensor<T> w({{"a", 4}, {"b", 2}}, {0, 1, 2, 3, 4, 5, 6, 7}), z({{"b", 2}, {"c", 3}}, {0, 1, 2, 3, 4, 5});
    EXPECT(Einsum::with_2(w, z).to({"a", "c"}) == (w * z));
    // tests/einsum.rs:76-88: ab,cb->ac == x * y^T
    Tensor<T> p({{"a", 2}, {"b", 3}}, {0, 10, 20, 1, 11, 21}), q({{"c", 2}, {"b", 3}}, {0, 10, 20, 1, 11, 21});
    Tensor<T> qt({{"b", 3}, {"c", 2}}, {0, 1, 10, 11, 20, 21});
    EXPECT(Einsum::with_2(p, q).to({"a", "c"}) == (p * qt));
    // tests/einsum.rs:133-163: batched matmul == per-batch *
    std::vector<T> xb(3 * 2 * 5), yb(3 * 5 * 3);
    for (size_t i = 0; i < xb.size(); i++) xb[i] = T((int)(i % 7) - 3);
    for (size_t i = 0; i < yb.size(); i++) yb[i] = T((int)(i % 5) - 2);
    Tensor<T> X({{"batch", 3}, {"a", 2}, {"b", 5}}, xb), Y({{"batch", 3}, {"b", 5}, {"c", 3}}, yb);
    auto Z = Einsum::with_2(X, Y).to({"batch", "a", "c"});
    for (int bt = 0; bt < 3; bt++) {
        Tensor<T> xs({{"a", 2}, {"b", 5}}, std::vector<T>(xb.begin() + bt * 10, xb.begin() + bt * 10 + 10));
        Tensor<T> ys({{"b", 5}, {"c", 3}}, std::vector<T>(yb.begin() + bt * 15, yb.begin() + bt * 15 + 15));
        auto zs = xs * ys;
        EXPECT(std::equal(zs.data().begin(), zs.data().end(), Z.data().begin() + bt * 6));
    }
    // tests/einsum.rs:208-226: three operands ab,cb,bd->ac; :13-27 one operand ab->ba
    Tensor<T> m23({{"a", 2}, {"b", 3}}, {0, 10, 20, 1, 11, 21}), n23({{"c", 2}, {"b", 3}}, {0, 10, 20, 1, 11, 21});
    Tensor<T> m34({{"b", 3}, {"d", 4}}, {0, 10, 20, 30, 1, 11, 21, 31, 2, 12, 22, 32});
    EXPECT(Einsum::with_3(m23, n23, m34).to({"a", "c"}) == Tensor<T>({{"a", 2}, {"c", 2}}, {33600, 35600, 35600, 37792}));
    EXPECT(Einsum::with_1(m23).to({"b", "a"}) == Tensor<T>({{"b", 3}, {"a", 2}}, {0, 1, 10, 11, 20, 21}));
    // Err(InconsistentDimensionLengthError)
    bool threw = false;
    try {
        (void)Einsum::with_2(x, y).to({"i", "nope"});
    } catch (const InconsistentDimensionLengthError& e) {
        threw = e.dimension == "nope" && !e.lengths[0] && !e.lengths[1];
    }
    EXPECT(threw);
}

template <class T>
static void record_goldens() {
    // src/differentiation/container_record/mod.rs:2718-2811: 4x3 * 3x4, numbers, every derivative, tape length
    WengertList<T> list;
    std::vector<T> av{1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 5, 2}, bv{1, 4, 7, 0, 2, 5, 8, 5, 3, 6, 9, 2};
    auto A = RecordTensor<T>::variables(list, Tensor<T>({{"r", 4}, {"c", 3}}, av));
    auto B = RecordTensor<T>::variables(list, Tensor<T>({{"x", 3}, {"y", 4}}, bv));
    EXPECT(list.len() == 24);
    EXPECT(A.indexes().front() == 0 && A.indexes().back() == 11 && B.indexes().front() == 12);
    auto Cm = A * B;
    EXPECT(Cm.shape() == (Shape{{"r", 4}, {"y", 4}}));
    EXPECT(Cm.numbers().data() == (std::vector<T>{14, 32, 50, 16, 32, 77, 122, 37, 50, 122, 194, 58, 16, 37, 58, 29}));
    // tape length of the scalar path: 16 outputs x (3 multiplies + 2 adds)   (mod.rs:2805-2810)
    EXPECT(list.len() == 24 + 16 * 5);
    // Index of C[i,j] = base + (i*N+j)*(2K-1) + 2K-2   (the last add of its chain)
    auto ci = Cm.indexes();
    EXPECT(ci[0] == 24 + 4 && ci[5] == 24 + 5 * 5 + 4 && ci[15] == 24 + 15 * 5 + 4);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            auto d = Cm.derivatives_for(i, j);
            auto dA = d.at_tensor(A), dB = d.at_tensor(B);
            for (int r = 0; r < 4; r++)
                for (int k = 0; k < 3; k++) EXPECT(dA.data()[r * 3 + k] == (r == i ? bv[k * 4 + j] : T(0)));
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 4; c++) EXPECT(dB.data()[k * 4 + c] == (c == j ? av[i * 3 + k] : T(0)));
            EXPECT(d.at(ci[i * 4 + j]) == T(1) && d.len() == list.len());
        }
    // container ops around the matmul (functions.rs rules): d/dx of sum-free chain y = exp(-x) at one element
    auto E = (-A).exp();
    auto de = E.derivatives_for(1, 2).at_tensor(A);
    EXPECT(std::fabs(de.data()[5] + std::exp(-T(6))) < T(1e-6) && de.data()[4] == T(0));
    EXPECT(panic_text([&] { (void)(A + B); }).rfind("Record containers must have the same shape", 0) == 0);
    // constants carry no history
    auto K = RecordTensor<T>::constants(Tensor<T>({{"x", 3}, {"y", 4}}, bv), list);
    EXPECT(!K.has_history());
    EXPECT(panic_text([&] { (void)K.derivatives_for(0, 0); }) == "Record has no WengertList to find derivatives from");
    list.clear();
    EXPECT(list.len() == 0);
    A.reset();
    EXPECT(list.len() == 12 && A.indexes().front() == 0);
}

// tests/distributions.rs:48-220: the multivariate Gaussian draw on the reference's fixed random source, and
// tests/linear_algebra.rs:175-188: Cholesky of the 3x3 example (exact)
static void distribution_goldens() {
    Matrix<float> spd(3, 3, {25, 15, -5, 15, 18, 0, -5, 0, 11});
    auto low = cholesky_decomposition(spd);
    EXPECT(low && *low == Matrix<float>(3, 3, {5, 0, 0, 3, 3, 0, -1, 1, 3}));
    EXPECT(!cholesky_decomposition(Matrix<double>(2, 2, {1, 2, 2, 1})));
    // src/linear_algebra.rs:1093-1096: (a - sum) * (1 / L[j][j]) -- two roundings: 5 * fl(1/3) = 0x1.aaaaaaaaaaaaap+0,
    // one ulp below fl(5/3)
    auto two = cholesky_decomposition(Matrix<double>(2, 2, {9, 5, 5, 10}));
    EXPECT(two && two->get(1, 0) == 0x1.aaaaaaaaaaaaap+0 && two->get(1, 0) != 5.0 / 3.0);
    const std::vector<double> numbers = {0.8040186166230938, 0.580253222290779, 0.3807224296769691, 0.21493968506238081, 0.7492549315762678, 0.3307385278792596, 0.7428234020730242, 0.924520495434979, 0.7706864587077074, 0.9606610369139641, 0.22464359232335274, 0.7572331492368785, 0.42525179149566283, 0.669874551931497, 0.4900294220194159, 0.8419224030915755, 0.2152677317521281, 0.9047234091130423, 0.5552287542435432, 0.0845469814310511, 0.41907285936702277, 0.20341459428573838, 0.348744633661872, 0.5312939758141078, 0.2277672193216529, 0.1688203334261118, 0.8382370210884449, 0.019565698056365877, 0.6201664569519008, 0.7479491404823113, 0.484013003448045, 0.7347758377654283, 0.336498328409339, 0.9071544418215385, 0.6634427419789561, 0.5844436681111158, 0.003789929995079211, 0.15610119074707995, 0.2150253132088742, 0.45046274535042086, 0.2120243842318259, 0.9372863373271367, 0.30736703756192063, 0.3679450747723736, 0.6542627475645721, 0.46319586776390764, 0.8711072275933169, 0.2902645933689898, 0.9720242533821568, 0.8461652295777833, 0.543197353963871, 0.170283015830482, 0.04273291417037717, 0.7186183438469571, 0.6060819803673965, 0.92429551494178, 0.9596189866410694, 0.9415763505896844, 0.264649901252882, 0.6987701655198049, 0.17563937503343774, 0.46796285074389776, 0.31485784595214716, 0.6719786444284983, 0.10138451740090293, 0.4985694635269784, 0.3525591176422236, 0.3939537113644429, 0.045029903348206446, 0.15755561321672773, 0.254414485279737, 0.7848580636066318, 0.05721098529487523, 0.7601928733446881, 0.6764204954104394, 0.7442849656738886, 0.28162126043898295, 0.21093997385733432, 0.9399064250864566, 0.7977861215591273, 0.5374688753828565, 0.8851215426025814, 0.41642367914454037, 0.2959207016044503, 0.2121000397454158, 0.4868594413558851, 0.4571656286244179, 0.13484646900784636, 0.4443762480787943, 0.7939780466365716, 0.7067786522378445, 0.6152187449980677, 0.4519149484716358, 0.7900479010107251, 0.9689655611261592, 0.8700246722278424, 0.3335030618321242, 0.20847389033467123, 0.7034874950351062, 0.8874968748931777};
    size_t next = 0;
    auto source = [&]() -> std::optional<double> {
        if (next >= numbers.size()) return std::nullopt;
        return numbers[next++];
    };
    const double mean[3] = {0.0, 10.0, -10.0}, var[3] = {1.0, 2.0, 1.0};
    MultivariateGaussian<double> gaussian(Matrix<double>(3, 1, {0.0, 10.0, -10.0}),
                                          Matrix<double>(3, 3, {1.0, 0.5, -0.1, 0.5, 2.0, 0.9, -0.1, 0.9, 1.0}));
    auto samples = gaussian.draw(source, 25);  // N = 3 is odd: 4 numbers per sample, all 100 consumed
    EXPECT(samples && samples->rows() == 25 && samples->columns() == 3 && next == 100);
    if (samples)
        for (int f = 0; f < 3; f++) {  // :175-193: each feature's sample mean within half a standard deviation
            double sum = 0;
            for (int m = 0; m < 25; m++) sum += samples->get(m, f);
            EXPECT(sum / 25 < mean[f] + 0.5 * std::sqrt(var[f]) && sum / 25 > mean[f] - 0.5 * std::sqrt(var[f]));
        }
    EXPECT(!gaussian.draw(source, 1));  // exhausted source -> None
    EXPECT(panic_text([&] { MultivariateGaussian<double> g(Matrix<double>(1, 3, {0, 0, 0}), Matrix<double>(3, 3, std::vector<double>(9, 1.0))); }) ==
           "Mean must be a column vector");
}

// Device-resident operator chains: same numbers as the host-operand operators, same panics
template <class T>
static void device_matrix_checks() {
    Matrix<T> a(3, 2, {1, 2, 3, 4, 5, 6}), b(2, 3, {1, 2, 3, 4, 5, 6}), c(3, 3, {1, 0, 0, 0, 1, 0, 0, 0, 1});
    DeviceMatrix<T> da(a), db(b), dc(c);
    EXPECT((da * db).to_host() == a * b);                                            // tests/matrices.rs:161-166
    EXPECT(((da * db) + dc).to_host() == Matrix<T>(3, 3, {10, 12, 15, 19, 27, 33, 29, 40, 52}));
    EXPECT((da * db - dc).transpose().to_host() == Matrix<T>(3, 3, {8, 19, 29, 12, 25, 40, 15, 33, 50}));
    EXPECT((-da).to_host() == Matrix<T>(3, 2, {-1, -2, -3, -4, -5, -6}));
    EXPECT((da * T(2) + T(1)).to_host() == Matrix<T>(3, 2, {3, 5, 7, 9, 11, 13}));
    EXPECT((da / T(2) - T(1)).to_host() == Matrix<T>(3, 2, {T(-0.5), 0, T(0.5), 1, T(1.5), 2}));
    EXPECT(da.transpose().to_host() == Matrix<T>(2, 3, {1, 3, 5, 2, 4, 6}));
    EXPECT(panic_text([&] { (void)(da * da); }) ==
           "Mismatched Matrices, left is 3x2, right is 3x2, * is only defined for MxN * NxL");
    EXPECT(panic_text([&] { (void)(da + db); }) ==
           "Mismatched matrices, left is 3x2, right is 2x3, + is only defined for MxN + MxN");
}

// f64 product through int8 slices: exact on small integers, falls back (slices_used = 0) below the kernel's range
static void sliced_product_checks() {
    const int n = 320;
    std::vector<double> a((size_t)n * n), b((size_t)n * n);
    for (int i = 0; i < n * n; i++) {
        a[i] = (double)((i * 7 + 3) % 17 - 8);
        b[i] = (double)((i * 5 + 1) % 13 - 6);
    }
    Matrix<double> A(n, n, a), B(n, n, b);
    DeviceMatrix<double> dA(A), dB(B);
    int used = -1;
    Matrix<double> fast = dA.mul_sliced(dB, 0, &used).to_host();
    EXPECT(used >= 7);
    EXPECT(fast == (dA * dB).to_host());  // integer data: both kernels are exact
    Matrix<double> s1(2, 2, {1, 2, 3, 4});
    DeviceMatrix<double> d1(s1);
    EXPECT(d1.mul_sliced(d1, 0, &used).to_host() == Matrix<double>(2, 2, {7, 10, 15, 22}) && used == 0);
}

int main() {
    int n = 0;
    if (eml_init(&n) != EML_OK) {
        // no device: the product path must refuse to run (no CPU fallback), with the library's message
        bool loud = false;
        try {
            Matrix<double> a(1, 1, {1.0});
            (void)(a * a);
        } catch (const BackendError& e) {
            loud = e.code == EML_ERR_NO_DEVICE;
            std::printf("no CUDA device: product path failed loudly as required: %s\n", e.what());
        }
        return loud ? 3 : 1;
    }
    matrix_goldens<float>();
    matrix_goldens<double>();
    tensor_goldens<float>();
    tensor_goldens<double>();
    einsum_goldens<float>();
    einsum_goldens<double>();
    record_goldens<float>();
    record_goldens<double>();
    distribution_goldens();
    device_matrix_checks<float>();
    device_matrix_checks<double>();
    sliced_product_checks();
    if (failures) {
        std::fprintf(stderr, "%d check(s) failed\n", failures);
        return 1;
    }
    std::printf("host_test: all reference known-answer tests passed through the C++ mirror (%lld kernel launches)\n",
                (long long)eml_kernel_launches());
    return 0;
}
