"""The compiled host layer (quantr_b200/host/quantr.hpp): the reference's tests replayed from C++ through the C ABI."""
import os
import subprocess

import pytest

from helpers import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "reference_tests")


def build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "reference_tests.cpp"),
                    "-L" + os.path.join(ROOT, "quantr_b200"), "-lqsv", "-Wl,-rpath," + os.path.join(ROOT, "quantr_b200"),
                    "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True, capture_output=True)


def test_cpp_host_builder_and_validation():
    build()
    out = subprocess.run([EXE, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout


@pytest.mark.gpu
def test_cpp_host_reference_golden_vectors_on_device():
    build()
    out = subprocess.run([EXE, "device"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout
