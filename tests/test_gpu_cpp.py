"""GPU tests of the C++ layers above the C ABI (prebuilt by __graft_entry__.build(), run as subprocesses):

* tests/cpp/test_mirror.cpp   -- include/genfft_cuda/fft.h, the class mirror of genfft::FFT / RealFFT / DIT /
                                 FFTVert / FFT2D, driven with host pointers in the reference tests' structure;
* oracle/ref_plugin_test.cpp  -- the REFERENCE's own classes (compiled from /root/reference) with the CUDA
                                 factories of include/genfft_cuda/backend.h as template arguments, compared with the
                                 reference's native CPU back-end in the same process.
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} was not prebuilt")
    r = subprocess.run([path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout


def test_cpp_class_mirror():
    run(os.path.join(ROOT, "tests", "cpp", "_build", "test_mirror"))


def test_reference_classes_with_cuda_factories():
    run(os.path.join(ROOT, "oracle", "_ref", "ref_plugin_test"))
