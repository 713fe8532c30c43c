"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, fails loudly without a device, and the host mirror validates like the reference BEFORE the FFI."""
import os
import re

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "easyml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eml_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) > 60
    for s in syms:
        assert hasattr(_ffi.lib, s), f"{s} declared in include/easyml_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_ffi.SIGNATURES) == syms


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "easyml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def has_gpu():
    import ctypes

    n = ctypes.c_int()
    return _ffi.lib.eml_init(ctypes.byref(n)) == 0


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_device_is_a_hard_error():
    with pytest.raises(eml.NoDeviceError):
        eml.device_count()
    with pytest.raises(eml.NoDeviceError):
        eml.Matrix([[1.0, 2.0]]) * eml.Matrix([[1.0], [2.0]])
    with pytest.raises(eml.NoDeviceError):
        eml.Tensor([("a", 3)], [1.0, 2.0, 3.0]).scalar_product(eml.Tensor([("a", 3)], [1.0, 2.0, 3.0]))
    with pytest.raises(eml.NoDeviceError):
        eml.WengertList(np.float32)
    x = eml.Tensor([("a", 2), ("b", 2)], [1.0, 2.0, 3.0, 4.0], dtype=np.float32)
    with pytest.raises(eml.NoDeviceError):
        eml.Einsum.with_2(x, x).named(["a", "b"], ["b", "c"]).to(["a", "c"])
    assert "no CPU fallback" in _ffi.last_error()


def test_host_validation_matches_reference_panics(golden):
    for g in golden["matmul_panic"]:
        with pytest.raises(eml.Panic) as e:
            eml.Matrix(np.zeros(g["a_shape"])) * eml.Matrix(np.zeros(g["b_shape"]))
        assert str(e.value) == g["message"]
    for g in golden["tensor_matmul_panic"]:
        a = eml.Tensor(list(zip(g["a_names"], g["a_shape"])), np.zeros(g["a_shape"]))
        b = eml.Tensor(list(zip(g["b_names"], g["b_shape"])), np.zeros(g["b_shape"]))
        with pytest.raises(eml.Panic) as e:
            a * b
        assert str(e.value).startswith(g["message_prefix"])


def test_host_validation_agrees_with_oracle_messages(orc):
    # the oracle restates the reference's panic text; the host mirror must produce the same strings
    cases = [((2, 3), ["r", "c"], (2, 3), ["a", "b"]), ((2, 3), ["rows", "columns"], (3, 2), ["columns", "rows"])]
    for ashape, an, bshape, bn in cases:
        with pytest.raises(orc.ReferencePanic) as want:
            orc.tensor_mul(np.zeros(ashape), an, np.zeros(bshape), bn)
        with pytest.raises(eml.Panic) as got:
            eml.Tensor(list(zip(an, ashape)), np.zeros(ashape)) * eml.Tensor(list(zip(bn, bshape)), np.zeros(bshape))
        assert str(got.value) == str(want.value)
    with pytest.raises(orc.ReferencePanic) as want:
        orc.vector_dot(np.zeros(3), "a", np.zeros(4), "a")
    with pytest.raises(eml.Panic) as got:
        eml.Tensor([("a", 3)], np.zeros(3)).scalar_product(eml.Tensor([("a", 4)], np.zeros(4)))
    assert str(got.value) == str(want.value)


def test_einsum_errors_match_oracle(orc):
    a = eml.Tensor([("a", 2), ("b", 3)], np.zeros((2, 3), np.float32))
    b = eml.Tensor([("b", 4), ("c", 5)], np.zeros((4, 5), np.float32))
    for out in (["a", "c"], ["a", "z"]):
        with pytest.raises(orc.InconsistentDimensionLengthError) as want:
            orc.einsum2(a.data, a.names(), b.data, b.names(), out)
        with pytest.raises(eml.InconsistentDimensionLengthError) as got:
            eml.Einsum.with_2(a, b).to(out)
        assert (got.value.dimension, got.value.lengths) == (want.value.dimension, want.value.lengths)


def test_non_float_types_are_not_this_paths_business():
    with pytest.raises(TypeError):
        eml.host._sfx(np.int32)


def test_rust_ffi_bindings_are_generated_from_the_header():
    """easy-ml_b200/rust/src/cuda/ffi.rs (the `cuda` feature's raw bindings) is generated from include/easyml_b200.h:
    the committed file must be up to date and declare every symbol the header does."""
    import importlib.util
    import re

    gen_path = os.path.join(ROOT, "easy-ml_b200", "rust", "gen_ffi.py")
    spec = importlib.util.spec_from_file_location("gen_ffi", gen_path)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(gen.OUT).read()
    header = open(os.path.join(ROOT, "include", "easyml_b200.h")).read()
    protos = list(gen.prototypes(header))
    assert len(protos) >= 70
    for _, name, _ in protos:
        assert re.search(rf"pub fn {name}\(", committed), f"{name} missing from ffi.rs"
    out = gen.OUT
    try:
        gen.OUT = out + ".check"
        gen.main()
        assert open(gen.OUT).read() == committed, "ffi.rs is stale: run python easy-ml_b200/rust/gen_ffi.py"
    finally:
        if os.path.exists(out + ".check"):
            os.remove(out + ".check")
        gen.OUT = out


def _rust_sources():
    base = os.path.join(ROOT, "easy-ml_b200", "rust")
    out = []
    for d, _, files in os.walk(base):
        out += [os.path.join(d, f) for f in files if f.endswith(".rs")]
    return sorted(out)


def _strip_rust(text):
    """Source with comments, string literals and char literals blanked (lifetimes such as 'a are kept)."""
    import re

    text = re.sub(r"/\*.*?\*/", lambda m: " " * len(m.group(0)), text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text, flags=re.S)
    text = re.sub(r"'(?:\\.|[^'\\])'", "' '", text)
    return text


def test_rust_shim_is_syntactically_sane_and_only_uses_symbols_that_exist():
    """There is no Rust toolchain in this image, so the shim cannot be compiled here.  What can be checked is checked:
    brackets balance in every .rs file, every `ffi::` function or constant the hand-written modules use is declared in
    the generated bindings, and every `crate::cuda::` item the dispatch sites call is defined by the module."""
    import re

    files = _rust_sources()
    assert any(f.endswith("dispatch_sites.rs") for f in files) and any(f.endswith("einsum.rs") for f in files)
    ffi = open(os.path.join(ROOT, "easy-ml_b200", "rust", "src", "cuda", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (eml_\w+)\(", ffi)) | set(re.findall(r"pub const (EML_\w+):", ffi)) | {
        "EmlVal", "EmlTape", "EmlDerivs", "EmlGraph"}
    defined = set()
    for f in files:
        if os.sep + "cuda" + os.sep in f:
            src = _strip_rust(open(f).read())
            defined |= set(re.findall(r"pub(?:\(crate\))?\s+(?:unsafe\s+)?(?:fn|struct|enum|mod|trait|const)\s+(\w+)", src))
    for f in files:
        src = _strip_rust(open(f).read())
        depth = {"(": 0, "[": 0, "{": 0}
        pairs = {")": "(", "]": "[", "}": "{"}
        for ch in src:
            if ch in depth:
                depth[ch] += 1
            elif ch in pairs:
                depth[pairs[ch]] -= 1
                assert depth[pairs[ch]] >= 0, f"{f}: unbalanced {ch}"
        assert all(v == 0 for v in depth.values()), f"{f}: unbalanced brackets {depth}"
        if f.endswith("ffi.rs"):
            continue
        for name in re.findall(r"(?<![:\w])ffi::(\w+)", src):  # (not core::ffi::c_int)
            assert name in declared, f"{f}: ffi::{name} is not in the generated bindings"
        for path in re.findall(r"crate::cuda::((?:\w+::)*\w+)", src):
            for part in path.split("::"):
                if part in ("Kind", "GENERIC", "F32", "F64") or part[0].isupper() and part in defined:
                    continue
                assert part in defined, f"{f}: crate::cuda::{path} -- `{part}` is not defined in src/cuda"


def test_static_library_is_built_and_exports_the_abi():
    """north_star: "a thin extern "C" layer over a CUDA static library built for sm_100a by build.rs" -- the Makefile
    archives the same objects as libeasyml_b200.a; every declared entry point must be defined in it."""
    import re
    import subprocess

    lib_a = os.path.join(ROOT, "easy-ml_b200", "lib", "libeasyml_b200.a")
    assert os.path.exists(lib_a), "run `make -C easy-ml_b200/csrc` (build() does)"
    header = open(os.path.join(ROOT, "include", "easyml_b200.h")).read()
    names = set(re.findall(r"^(?:int|void|int64_t|const char\*)\s+(eml_\w+)\(", header, flags=re.M))
    syms = subprocess.run(["nm", "--defined-only", lib_a], capture_output=True, text=True, check=True).stdout
    have = set(re.findall(r" T (eml_\w+)", syms))
    assert names <= have, sorted(names - have)
    build_rs = open(os.path.join(ROOT, "easy-ml_b200", "rust", "build.rs")).read()
    assert "static=easyml_b200" in build_rs and "dylib=easyml_b200" not in build_rs
