"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (as the checker or the timed CPU baseline); the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

CUSTOM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_double))


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.oracle_simulate_faithful.restype = C.c_int
        lib.oracle_simulate_faithful.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        lib.oracle_simulate_dense.restype = C.c_int
        lib.oracle_simulate_dense.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_int]
        for name in ("oracle_measure_all", "oracle_measure_all_cdf"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def simulate(n, ops, n_ops, register=None, *, mode="dense", threads=1, custom_callback=None):
    """Runs the oracle on a qsv_op array.  `register`: complex128[2^n] or None for |0..0>.

    mode "faithful": per-gate hash-map restatement of src/circuit/simulation.rs:64-135.
    mode "dense":    same arithmetic per group of coupled amplitudes.
    custom_callback(op_index, qubits:list[int]) -> complex array | None: called like the reference
    calls a Custom closure (once per basis state in faithful mode); default: the op's matrix/none_mask.
    """
    lib = load()
    amps = np.zeros(1 << n, dtype=np.complex128)
    if register is None:
        amps[0] = 1.0
    else:
        amps[:] = np.asarray(register, dtype=np.complex128)
    cb = None
    if custom_callback is not None:
        def _cb(ctx, op_index, qubits, k, out):
            res = custom_callback(op_index, [qubits[i] for i in range(k)])
            if res is None:
                return 0
            res = np.asarray(res, dtype=np.complex128)
            for t in range(1 << k):
                out[2 * t] = res[t].real
                out[2 * t + 1] = res[t].imag
            return 1
        cb = CUSTOM_FN(_cb)
    cb_ptr = C.cast(cb, C.c_void_p) if cb is not None else None
    ops_ptr = C.cast(ops, C.c_void_p)
    if mode == "faithful":
        rc = lib.oracle_simulate_faithful(n, ops_ptr, n_ops, _dp(amps), cb_ptr, None)
    else:
        rc = lib.oracle_simulate_dense(n, ops_ptr, n_ops, _dp(amps), cb_ptr, None, threads)
    if rc != 0:
        raise ValueError("oracle rejected the circuit")
    return amps


def measure_all(n, amps, uniforms, *, cdf=True):
    """SuperPosition::measure per uniform (super_positions.rs:332-342); UINT64_MAX = failed to collapse."""
    lib = load()
    a = np.ascontiguousarray(amps, dtype=np.complex128)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    out = np.empty(u.shape[0], dtype=np.uint64)
    fn = lib.oracle_measure_all_cdf if cdf else lib.oracle_measure_all
    fn(n, _dp(a), _dp(u), u.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out
