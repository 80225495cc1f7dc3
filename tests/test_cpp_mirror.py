"""The compiled host side above the C ABI: include/ndarray_interp_b200.hpp (C++17 mirror of the
reference API).  CPU: the header and tests/cpp/test_mirror.cpp compile and link against the in-tree
library.  GPU: the program runs the reference's own test expectations through it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ndarray_interp_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_mirror.bin")


def _build():
    from ndarray_interp_b200.build import build
    build()
    gxx = shutil.which("g++")
    assert gxx, "g++ not found"
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"),
           "-o", EXE, "-L", PKG, "-lndi_b200", "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_mirror_runs_the_reference_expectations():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failures" in r.stdout
