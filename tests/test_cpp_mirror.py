"""Builds and runs the C++ mirror of the reference API (include/huff_coding.hpp) against libhuffb200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_huff_coding")


def _build():
    from huff_encoding_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_huff_coding.cpp")
    lib_dir = os.path.join(ROOT, "huff_encoding_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + lib_dir, "-lhuffb200",
                               "-Wl,-rpath," + lib_dir])


def test_cpp_mirror_host_part():
    _build()
    out = subprocess.run([EXE, "host"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu():
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
