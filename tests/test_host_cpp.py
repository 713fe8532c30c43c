"""The C++ mirror of the reference operator API (easy-ml_b200/host/easyml.hpp) over the C ABI.

host_test replays the reference's own known-answer tests for the path (tests/matrices.rs:161-173,
src/tensors/operations.rs:521-546,1655-1708, tests/einsum.rs:60-206, container_record/mod.rs:2718-2811)
with the reference's panic messages.  On a box without a GPU it must fail loudly: no CPU fallback.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "easy-ml_b200", "host")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "easy-ml_b200", "csrc"), "-s", "-j", "8"])
    subprocess.check_call(["make", "-C", HOST, "-s", "all"])
    return os.path.join(HOST, "host_test")


def test_cpp_mirror_builds_and_refuses_to_run_without_a_device():
    exe = _build()
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the gpu-marked test runs the known-answer tests")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_known_answers_on_gpu():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference known-answer tests passed" in r.stdout
