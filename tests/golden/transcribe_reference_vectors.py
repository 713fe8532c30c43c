#!/usr/bin/env python
"""Writes tests/golden/reference_vectors.json.

The reference (Skeletonxf/easy-ml) is Rust and cannot be built or imported in this image, so the golden
vectors are the literal inputs/outputs of the reference's OWN tests for the hot path, transcribed here
with the file:line they come from (paths relative to /root/reference).  Nothing is computed: every
"expect" below is a literal that appears in the reference's test source (integer arithmetic written out
in the source, e.g. `1 * 10 + 2 * 12 + 3 * 14`, is evaluated).  Run: python tests/golden/transcribe_reference_vectors.py
"""
import json
import os


def randomish(shape):  # tests/einsum.rs:9-11  Tensor::from_fn(shape, |[x, y]| (x + (10 * y)) as f32)
    return [[float(x + 10 * y) for y in range(shape[1])] for x in range(shape[0])]


V = {"matmul": [], "matmul_panic": [], "tensor_matmul": [], "tensor_matmul_panic": [], "einsum2": [],
     "einsum2_equiv": [], "record_matmul": [], "xor_nn": {}}

# ---- A2 Matrix * Matrix ------------------------------------------------------------------------
V["matmul"].append({
    "cite": "tests/matrices.rs:161-166",
    "a": [[1, 2], [3, 4], [5, 6]], "b": [[1, 2, 3], [4, 5, 6]],
    "expect": [[9, 12, 15], [19, 26, 33], [29, 40, 51]]})
V["matmul"].append({
    "cite": "tests/linear_algebra.rs:183-190 (f32: L * L^T == A exactly)",
    "a": [[5.0, 0.0, 0.0], [3.0, 3.0, 0.0], [-1.0, 1.0, 3.0]],
    "b": [[5.0, 3.0, -1.0], [0.0, 3.0, 1.0], [0.0, 0.0, 3.0]],
    "expect": [[25.0, 15.0, -5.0], [15.0, 18.0, 0.0], [-5.0, 0.0, 11.0]]})
V["matmul"].append({
    "cite": "src/differentiation/container_record/mod.rs:2816-2865 (numbers of RecordMatrix product)",
    "a": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0], [0.0, 5.0, 2.0]],
    "b": [[1.0, 4.0, 7.0, 0.0], [2.0, 5.0, 8.0, 5.0], [3.0, 6.0, 9.0, 2.0]],
    "expect": [[14.0, 32.0, 50.0, 16.0], [32.0, 77.0, 122.0, 37.0], [50.0, 122.0, 194.0, 58.0],
               [16.0, 37.0, 58.0, 29.0]]})
V["matmul_panic"].append({
    "cite": "tests/matrices.rs:168-173 (#[should_panic]); message src/matrices/operations.rs:176-183",
    "a_shape": [3, 2], "b_shape": [3, 2],
    "message": "Mismatched Matrices, left is 3x2, right is 3x2, * is only defined for MxN * NxL"})

# ---- A3 Tensor<T,2> * Tensor<T,2> --------------------------------------------------------------
V["tensor_matmul"].append({
    "cite": "src/tensors/operations.rs:521-546",
    "a": [[1, 2, 3], [4, 5, 6]], "a_names": ["r", "c"],
    "b": [[10, 11], [12, 13], [14, 15]], "b_names": ["r", "c"],
    "expect": [[1 * 10 + 2 * 12 + 3 * 14, 1 * 11 + 2 * 13 + 3 * 15],
               [4 * 10 + 5 * 12 + 6 * 14, 4 * 11 + 5 * 13 + 6 * 15]],
    "expect_names": ["r", "c"]})
V["tensor_matmul"].append({
    "cite": "src/tensors/operations.rs:1655-1708 (all 16 owned/ref/view combinations agree)",
    "a": [[1, 2, 3], [4, 5, 6]], "a_names": ["r", "c"],
    "b": [[1, 2], [3, 4], [5, 6]], "b_names": ["a", "b"],
    "expect": [[1 * 1 + 2 * 3 + 3 * 5, 1 * 2 + 2 * 4 + 3 * 6], [4 * 1 + 5 * 3 + 6 * 5, 4 * 2 + 5 * 4 + 6 * 6]],
    "expect_names": ["r", "b"]})
V["tensor_matmul"].append({
    "cite": "tests/linear_algebra.rs:192-201 (f32 tensor L * L^T == A exactly)",
    "a": [[5.0, 0.0, 0.0], [3.0, 3.0, 0.0], [-1.0, 1.0, 3.0]], "a_names": ["r", "c"],
    "b": [[5.0, 3.0, -1.0], [0.0, 3.0, 1.0], [0.0, 0.0, 3.0]], "b_names": ["r", "c"],
    "expect": [[25.0, 15.0, -5.0], [15.0, 18.0, 0.0], [-5.0, 0.0, 11.0]], "expect_names": ["r", "c"]})
V["tensor_matmul"].append({
    "cite": "src/differentiation/container_record/mod.rs:2721-2775 (numbers of RecordTensor product)",
    "a": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0], [0.0, 5.0, 2.0]], "a_names": ["r", "c"],
    "b": [[1.0, 4.0, 7.0, 0.0], [2.0, 5.0, 8.0, 5.0], [3.0, 6.0, 9.0, 2.0]], "b_names": ["r", "c"],
    "expect": [[14.0, 32.0, 50.0, 16.0], [32.0, 77.0, 122.0, 37.0], [50.0, 122.0, 194.0, 58.0],
               [16.0, 37.0, 58.0, 29.0]], "expect_names": ["r", "c"]})
V["tensor_matmul_panic"].append({
    "cite": "src/tensors/operations.rs:14-18,492-501 (rows x columns * columns x rows is forbidden)",
    "a_names": ["rows", "columns"], "a_shape": [2, 3], "b_names": ["columns", "rows"], "b_shape": [3, 2],
    "message_prefix": "Matrix multiplication of tensors with shapes left [(\"rows\", 2), (\"columns\", 3)] and right "
                      "[(\"columns\", 3), (\"rows\", 2)] would create duplicate dimension names as the shape "
                      "[(\"rows\", 2), (\"rows\", 2)]."})
V["tensor_matmul_panic"].append({
    "cite": "src/tensors/operations.rs:485-491",
    "a_names": ["r", "c"], "a_shape": [2, 3], "b_names": ["a", "b"], "b_shape": [2, 3],
    "message_prefix": "Mismatched tensors, left is [(\"r\", 2), (\"c\", 3)], right is [(\"a\", 2), (\"b\", 3)], "
                      "* is only defined for MxN * NxL dimension lengths"})

# ---- A4 Einsum2 --------------------------------------------------------------------------------
V["einsum2"].append({
    "cite": "tests/einsum.rs:165-187 (ij,jk->ijk)",
    "a": randomish([2, 3]), "a_names": ["i", "j"], "b": randomish([3, 4]), "b_names": ["j", "k"],
    "out_names": ["i", "j", "k"],
    "expect": [[[0.0, 0.0, 0.0, 0.0], [10.0, 110.0, 210.0, 310.0], [40.0, 240.0, 440.0, 640.0]],
               [[0.0, 10.0, 20.0, 30.0], [11.0, 121.0, 231.0, 341.0], [42.0, 252.0, 462.0, 672.0]]]})
V["einsum2"].append({
    "cite": "tests/einsum.rs:189-206 (ij,kl->il)",
    "a": randomish([2, 3]), "a_names": ["i", "j"], "b": randomish([3, 4]), "b_names": ["k", "l"],
    "out_names": ["i", "l"],
    "expect": [[90.0, 990.0, 1890.0, 2790.0], [99.0, 1089.0, 2079.0, 3069.0]]})
# equivalences the reference asserts between Einsum and `*` (both sides recomputed by the test)
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:60-74 (ab,cb->ac == x * y^T, matrix-vector)",
    "a": randomish([2, 3]), "a_names": ["a", "b"], "b": randomish([1, 3]), "b_names": ["c", "b"],
    "out_names": ["a", "c"], "equiv": "a @ b.T"})
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:76-88 (ab,cb->ac == x * y^T)",
    "a": randomish([2, 3]), "a_names": ["a", "b"], "b": randomish([2, 3]), "b_names": ["c", "b"],
    "out_names": ["a", "c"], "equiv": "a @ b.T"})
V["einsum2_equiv"].append({
    "cite": "src/tensors/einsum.rs:58-72 doctest (ab,bc->ac == w * x)",
    "a": [[float(a * 2 + b) for b in range(2)] for a in range(4)], "a_names": ["a", "b"],
    "b": [[float(b * 3 + c) for c in range(3)] for b in range(2)], "b_names": ["b", "c"],
    "out_names": ["a", "c"], "equiv": "a @ b"})
V["einsum2_equiv"].append({
    "cite": "src/tensors/einsum.rs:74-79 doctest (ab,ac->bc == w^T * z)",
    "a": [[float(a * 2 + b) for b in range(2)] for a in range(4)], "a_names": ["a", "b"],
    "b": [[float(a * 2 + c) for c in range(3)] for a in range(4)], "b_names": ["a", "c"],
    "out_names": ["b", "c"], "equiv": "a.T @ b"})
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:133-163 (batch a b,batch b c->batch a c == per-batch *)",
    "a": [[[float(i + 10 * j + 2 * k) for k in range(5)] for j in range(2)] for i in range(3)],
    "a_names": ["batch", "a", "b"],
    "b": [[[float(i + 10 * j + 2 * k) for k in range(3)] for j in range(5)] for i in range(3)],
    "b_names": ["batch", "b", "c"],
    "out_names": ["batch", "a", "c"], "equiv": "batched a @ b"})
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:90-98 (a,a-> == scalar_product)",
    "a": [0.0, 1.0, 2.0], "a_names": ["a"], "b": [0.0, 1.0, 2.0], "b_names": ["a"], "out_names": [],
    "equiv": "dot"})
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:100-119 (ab,ab-> and ab,ab->ab == elementwise)",
    "a": randomish([2, 3]), "a_names": ["a", "b"], "b": randomish([2, 3]), "b_names": ["a", "b"],
    "out_names": ["a", "b"], "equiv": "a * b"})
V["einsum2_equiv"].append({
    "cite": "tests/einsum.rs:121-131 (a,b->ab outer product)",
    "a": [0.0, 1.0, 2.0], "a_names": ["a"], "b": [0.0, 10.0, 20.0, 30.0, 40.0], "b_names": ["b"],
    "out_names": ["a", "b"], "equiv": "outer"})

# ---- A6-A9 RecordTensor / RecordMatrix matmul --------------------------------------------------
V["record_matmul"].append({
    "cite": "src/differentiation/container_record/mod.rs:2718-2811 and :2813-2903",
    "dtype": "f64",
    "a": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0], [0.0, 5.0, 2.0]],
    "b": [[1.0, 4.0, 7.0, 0.0], [2.0, 5.0, 8.0, 5.0], [3.0, 6.0, 9.0, 2.0]],
    "expect": [[14.0, 32.0, 50.0, 16.0], [32.0, 77.0, 122.0, 37.0], [50.0, 122.0, 194.0, 58.0],
               [16.0, 37.0, 58.0, 29.0]],
    "asserts": ["numbers equal the scalar Record path", "every dC[i,j]/dA and dC[i,j]/dB equal the scalar path",
                "tape lengths equal"]})

# ---- end-to-end tape semantics: XOR network ----------------------------------------------------
V["xor_nn"] = {
    "cite": "tests/differentiation_nn.rs:23-120 (f32, ReLU, lr 0.1, 300 epochs, asserts losses[299] < 0.02)",
    "w1": [[0.1, -0.1, -0.1], [0.5, 0.4, -0.1], [0.3, 0.5, 0.2]], "w2": [[0.3], [0.1], [-0.4]],
    "inputs": [[0.0, 0.0, 1.0], [0.0, 1.0, 1.0], [1.0, 0.0, 1.0], [1.0, 1.0, 1.0]],
    "labels": [0.0, 1.0, 1.0, 0.0], "learning_rate": 0.1, "epochs": 300, "final_loss_below": 0.02}

# ---- next row: covariance ---------------------------------------------------------
V["covariance"] = [
    {"cite": "tests/linear_algebra.rs:157-170 (covariance_column_features::<f32>, assert_eq!: exact)",
     "x": [[1, 1, -1], [-1, -1, 1], [-1, -1, 1], [1, 1, -1]], "feature_axis": 1,
     "expect": [[1, 1, -1], [1, 1, -1], [-1, -1, 1]], "exact": True},
    {"cite": "src/linear_algebra.rs:739-763 doctest (tensor.covariance(\"features\"), output printed to 3 decimals)",
     "x": [[1.0, 0.0, 0.5], [1.2, -1.0, 0.4], [1.8, -1.2, 0.7], [0.9, 0.1, 0.3], [0.7, 0.5, 0.6]], "feature_axis": 1,
     "names": ["samples", "features"], "feature_dimension": "features",
     "expect": [[0.142, -0.226, 0.026], [-0.226, 0.438, -0.022], [0.026, -0.022, 0.020]], "exact": False,
     "decimals": 3}]

# ---- next row: one-operand and 3..6-operand einsum --------------------------------
def from_fn(shape, f):  # Tensor::from_fn over a row-major index walk
    import itertools
    import functools
    flat = [float(f(*ix)) for ix in itertools.product(*[range(n) for n in shape])]
    def nest(vals, dims):
        if len(dims) == 1:
            return vals
        step = functools.reduce(lambda a, b: a * b, dims[1:], 1)
        return [nest(vals[i * step:(i + 1) * step], dims[1:]) for i in range(dims[0])]
    return nest(flat, list(shape))


W_ = from_fn((4, 2), lambda w, x: w * 2 + x)            # tests/einsum.rs:232, 254, 283, 312
X_ = from_fn((2, 3), lambda x, y: x * 3 + y)
Y_ = from_fn((2, 3, 2), lambda x, y, z: x * 6 + y * 2 + z)
Z_ = from_fn((4, 2), lambda w, z: w * 2 + z)
NEGX_ = from_fn((2, 3), lambda x, y: -(x * 3 + y))
V["einsum_n"] = [
    {"cite": "tests/einsum.rs:13-27 (transpose ab->ba == index_by([b, a]))", "operands": [randomish((3, 2))],
     "names": [["a", "b"]], "out_names": ["b", "a"], "equiv": "x0.T"},
    {"cite": "tests/einsum.rs:29-37 (summation ab-> == iter().sum())", "operands": [randomish((2, 3))],
     "names": [["a", "b"]], "out_names": [], "equiv": "sequential_sum(x0)"},
    {"cite": "tests/einsum.rs:39-47 (column_sum ab->b)", "operands": [randomish((3, 2))], "names": [["a", "b"]],
     "out_names": ["b"], "equiv": "x0.sum(axis=0)"},
    {"cite": "tests/einsum.rs:49-57 (row_sum ab->a)", "operands": [randomish((3, 2))], "names": [["a", "b"]],
     "out_names": ["a"], "equiv": "x0.sum(axis=1)"},
    {"cite": "tests/einsum.rs:208-226 (ab,cb,bd->ac)", "operands": [randomish((2, 3)), randomish((2, 3)), randomish((3, 4))],
     "names": [["a", "b"], ["c", "b"], ["b", "d"]], "out_names": ["a", "c"],
     "expect": [[33600.0, 35600.0], [35600.0, 37792.0]]},
    {"cite": "tests/einsum.rs:228-247 (wx,xy,xyz,wz->xz)", "operands": [W_, X_, Y_, Z_],
     "names": [["w", "x"], ["x", "y"], ["x", "y", "z"], ["w", "z"]], "out_names": ["x", "z"],
     "expect": [[560.0, 884.0], [6800.0, 9408.0]]},
    {"cite": "tests/einsum.rs:249-275 (wx,xy,xyz,wz,zyx->xz with .named)", "operands": [W_, X_, Y_, Z_, Y_],
     "names": [["w", "x"], ["x", "y"], ["x", "y", "z"], ["w", "z"], ["z", "y", "x"]], "out_names": ["x", "z"],
     "expect": [[2016.0, 8432.0], [24752.0, 90384.0]]},
    {"cite": "tests/einsum.rs:277-304 (wx,xy,xyz,wz,xyz,xy->xyz)", "operands": [W_, X_, Y_, Z_, Y_, NEGX_],
     "names": [["w", "x"], ["x", "y"], ["x", "y", "z"], ["w", "z"], ["x", "y", "z"], ["x", "y"]],
     "out_names": ["x", "y", "z"],
     "expect": [[[0.0, 0.0], [-224.0, -612.0], [-3584.0, -6800.0]],
                [[-22032.0, -37044.0], [-69632.0, -108864.0], [-170000.0, -254100.0]]]},
    {"cite": "tests/einsum.rs:306-319 (wx,xy,xyz,wz,xyz,xy->)", "operands": [W_, X_, Y_, Z_, Y_, NEGX_],
     "names": [["w", "x"], ["x", "y"], ["x", "y", "z"], ["w", "z"], ["x", "y", "z"], ["x", "y"]], "out_names": [],
     "expect": -672892.0},
]

# ---- next row: GEMM callers -- Cholesky and the multivariate Gaussian draw (mean + L * z) ----
V["cholesky"] = [
    {"cite": "tests/linear_algebra.rs:175-188 (f32, assert_eq!: exact; L * L^T == A exactly)",
     "a": [[25.0, 15.0, -5.0], [15.0, 18.0, 0.0], [-5.0, 0.0, 11.0]],
     "expect": [[5.0, 0.0, 0.0], [3.0, 3.0, 0.0], [-1.0, 1.0, 3.0]], "exact": True},
    {"cite": "tests/linear_algebra.rs:206-240 (f64, sum of absolute differences < 0.0001)",
     "a": [[18.0, 22.0, 54.0, 42.0], [22.0, 70.0, 86.0, 62.0], [54.0, 86.0, 174.0, 134.0], [42.0, 62.0, 134.0, 106.0]],
     "expect": [[4.24264, 0.0, 0.0, 0.0], [5.18545, 6.56591, 0.0, 0.0], [12.72792, 3.04604, 1.64974, 0.0],
                [9.89949, 1.62455, 1.84971, 1.39262]], "exact": False, "abs_sum_tolerance": 0.0001}]
V["multivariate_gaussian"] = {
    "cite": "tests/distributions.rs:48-220 (f64; 100 source numbers, N = 3 rounds to 4 per sample => 25 samples; each "
            "feature's sample mean within half a standard deviation of its mean; Tensor draw == Matrix draw)",
    "mean": [0.0, 10.0, -10.0],
    "covariance": [[1.0, 0.5, -0.1], [0.5, 2.0, 0.9], [-0.1, 0.9, 1.0]],
    "max_samples": 25,
    "source": [
    0.8040186166230938, 0.580253222290779, 0.3807224296769691, 0.21493968506238081, 0.7492549315762678,
    0.3307385278792596, 0.7428234020730242, 0.924520495434979, 0.7706864587077074, 0.9606610369139641,
    0.22464359232335274, 0.7572331492368785, 0.42525179149566283, 0.669874551931497, 0.4900294220194159,
    0.8419224030915755, 0.2152677317521281, 0.9047234091130423, 0.5552287542435432, 0.0845469814310511,
    0.41907285936702277, 0.20341459428573838, 0.348744633661872, 0.5312939758141078, 0.2277672193216529,
    0.1688203334261118, 0.8382370210884449, 0.019565698056365877, 0.6201664569519008, 0.7479491404823113,
    0.484013003448045, 0.7347758377654283, 0.336498328409339, 0.9071544418215385, 0.6634427419789561,
    0.5844436681111158, 0.003789929995079211, 0.15610119074707995, 0.2150253132088742, 0.45046274535042086,
    0.2120243842318259, 0.9372863373271367, 0.30736703756192063, 0.3679450747723736, 0.6542627475645721,
    0.46319586776390764, 0.8711072275933169, 0.2902645933689898, 0.9720242533821568, 0.8461652295777833,
    0.543197353963871, 0.170283015830482, 0.04273291417037717, 0.7186183438469571, 0.6060819803673965,
    0.92429551494178, 0.9596189866410694, 0.9415763505896844, 0.264649901252882, 0.6987701655198049,
    0.17563937503343774, 0.46796285074389776, 0.31485784595214716, 0.6719786444284983, 0.10138451740090293,
    0.4985694635269784, 0.3525591176422236, 0.3939537113644429, 0.045029903348206446, 0.15755561321672773,
    0.254414485279737, 0.7848580636066318, 0.05721098529487523, 0.7601928733446881, 0.6764204954104394,
    0.7442849656738886, 0.28162126043898295, 0.21093997385733432, 0.9399064250864566, 0.7977861215591273,
    0.5374688753828565, 0.8851215426025814, 0.41642367914454037, 0.2959207016044503, 0.2121000397454158,
    0.4868594413558851, 0.4571656286244179, 0.13484646900784636, 0.4443762480787943, 0.7939780466365716,
    0.7067786522378445, 0.6152187449980677, 0.4519149484716358, 0.7900479010107251, 0.9689655611261592,
    0.8700246722278424, 0.3335030618321242, 0.20847389033467123, 0.7034874950351062, 0.8874968748931777,
    ]}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
with open(out, "w") as f:
    json.dump(V, f, indent=1)
print("wrote", out)
