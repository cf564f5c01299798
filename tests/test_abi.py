"""The C-ABI library loads and exports every symbol include/qsv.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT
from quantr_b200 import _ffi as F


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qsv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsv_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(F.EXPORTED_SYMBOLS)


def test_rust_shim_declares_the_same_exports():
    """rust/quantr-b200-shim is source-only (no rustc in the image): at least keep its extern "C" block in step with the header."""
    text = open(os.path.join(ROOT, "rust", "quantr-b200-shim", "src", "lib.rs")).read()
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n    }\n")]
    assert sorted(set(re.findall(r"pub fn (qsv_[a-z_0-9]+)\s*\(", block))) == declared_functions()


def test_library_exports_every_declared_symbol():
    lib = F.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    assert C.sizeof(F.QsvOp) == 56
    assert C.sizeof(F.QsvStats) == 72
    assert F.QsvOp.controls.offset == 16 and F.QsvOp.param.offset == 24 and F.QsvOp.matrix.offset == 40


def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device qsv_create must fail with a message."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    lib = F.load_library()
    h = C.c_void_p()
    rc = lib.qsv_create(C.byref(h), 3, 0)
    assert rc == 3 and not h.value
    assert b"no CPU fallback" in lib.qsv_last_error(None)
    import quantr_b200 as qb
    with pytest.raises(F.QsvError):
        qb.Circuit.new(2).add_gate(qb.Gate.H, 0).simulate()


def test_create_multi_validates_its_arguments_and_has_no_fallback(monkeypatch):
    """qsv_create_multi (one handle over several GPUs of the process): argument errors are reported before any device
    is touched; without a GPU it fails as loudly as qsv_create, through the reference-facing API (QSV_DEVICES) too."""
    lib = F.load_library()
    h = C.c_void_p()
    dev3 = (C.c_int32 * 3)(0, 1, 2)
    assert lib.qsv_create_multi(C.byref(h), 10, dev3, 3) == 1 and not h.value  # not a power of two
    assert b"power of two" in lib.qsv_last_error(None)
    dup = (C.c_int32 * 2)(1, 1)
    assert lib.qsv_create_multi(C.byref(h), 10, dup, 2) == 1 and b"twice" in lib.qsv_last_error(None)
    assert lib.qsv_create_multi(C.byref(h), 10, None, 2) == 1
    assert lib.qsv_create_multi(None, 10, dup, 2) == 1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    two = (C.c_int32 * 2)(0, 1)
    rc = lib.qsv_create_multi(C.byref(h), 10, two, 2)
    assert rc != 0 and not h.value
    import quantr_b200 as qb
    monkeypatch.setenv("QSV_DEVICES", "0,1")
    with pytest.raises(F.QsvError):
        qb.Circuit.new(8).add_gate(qb.Gate.H, 0).simulate()


def test_null_handles_are_rejected():
    lib = F.load_library()
    assert lib.qsv_init_basis(None, 0) == 1
    assert lib.qsv_destroy(None) == 0
    assert lib.qsv_plan_destroy(None) == 0


def test_plain_c_consumer(tmp_path):
    """include/qsv.h is a C header: a C99 program (what cgo or bindgen would see) compiles against it with -pedantic,
    links libqsv.so, finds the struct layout the bindings assume and drives the plan API and its error paths."""
    import subprocess
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.join(ROOT, "quantr_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-L" + libdir, "-lqsv", "-Wl,-rpath," + libdir,
                    "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout + out.stderr
