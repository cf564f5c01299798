"""ctypes binding of the C ABI in include/qsv.h (libqsv.so).

This is the reference-side binding a maintainer would write for the FFI seam; the Rust
equivalent (`extern "C"` block + safe wrapper) is shown in INTEGRATION.md.

There is no CPU fallback: if libqsv.so is missing or no CUDA device is usable the calls
raise.  Host-only entry points (qsv_plan_*) work without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QSV_LIB", os.path.join(_HERE, "libqsv.so"))  # QSV_LIB: developer override (A/B builds)

# gate kinds, src/circuit/gate.rs:19-106 in declaration order (include/qsv.h)
(GATE_ID, GATE_H, GATE_X, GATE_Y, GATE_Z, GATE_S, GATE_SDAG, GATE_T, GATE_TDAG, GATE_RX, GATE_RY, GATE_RZ,
 GATE_X90, GATE_Y90, GATE_MX90, GATE_MY90, GATE_PHASE, GATE_CR, GATE_CRK, GATE_CZ, GATE_CY, GATE_CNOT,
 GATE_SWAP, GATE_TOFFOLI, GATE_CUSTOM) = range(25)

ERR_NAMES = {0: "ok", 1: "invalid argument", 2: "out of memory", 3: "CUDA error", 4: "NCCL error",
             5: "unsupported", 6: "internal error"}
(ERR_INVALID_ARG, ERR_OUT_OF_MEMORY, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED, ERR_INTERNAL) = range(1, 7)
UINT64_MAX = (1 << 64) - 1


class QsvOp(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32),
        ("target", C.c_uint32),
        ("n_controls", C.c_uint32),
        ("reserved", C.c_uint32),
        ("controls", C.POINTER(C.c_uint32)),
        ("param", C.c_double),
        ("iparam", C.c_int32),
        ("reserved2", C.c_int32),
        ("matrix", C.POINTER(C.c_double)),
        ("none_mask", C.POINTER(C.c_uint8)),
    ]


class QsvStats(C.Structure):
    _fields_ = [
        ("n_gates", C.c_uint64),
        ("n_passes", C.c_uint64),
        ("n_rounds", C.c_uint64),
        ("n_kernel_launches", C.c_uint64),
        ("bytes_per_pass", C.c_uint64),
        ("n_exchanges", C.c_uint64),
        ("exchange_bytes", C.c_uint64),
        ("device_ms", C.c_double),
        ("exchange_ms", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class QsvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"qsv error {code} ({ERR_NAMES.get(code, '?')}): {message}")
        self.code = code
        self.message = message


_lib = None


def load_library():
    """Loads libqsv.so (built in-tree by `make lib` / __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make lib` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "quantr_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, i64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int64, C.c_size_t
    dp = C.POINTER(C.c_double)
    sigs = {
        "qsv_create": [C.POINTER(vp), u32, i32],
        "qsv_create_sharded": [C.POINTER(vp), u32, i32, i32, i32, vp, sz],
        "qsv_create_multi": [C.POINTER(vp), u32, C.POINTER(i32), i32],
        "qsv_nccl_unique_id": [vp, sz],
        "qsv_peer_export": [vp, vp, sz],
        "qsv_peer_import": [vp, vp, sz],
        "qsv_destroy": [vp],
        "qsv_set_option": [vp, C.c_char_p, i64],
        "qsv_get_info": [vp, C.c_char_p, C.POINTER(i64)],
        "qsv_init_basis": [vp, u64],
        "qsv_upload": [vp, dp, u64, u64],
        "qsv_download": [vp, dp, u64, u64],
        "qsv_gather": [vp, C.POINTER(u64), u64, dp],
        "qsv_apply": [vp, C.POINTER(QsvOp), sz, C.POINTER(QsvStats)],
        "qsv_plan_create": [C.POINTER(vp), u32, u32, C.POINTER(QsvOp), sz, u32, u32, i32],
        "qsv_plan_create_ex": [C.POINTER(vp), u32, u32, C.POINTER(QsvOp), sz, u32, u32, i32, C.POINTER(C.c_uint8), i32],
        "qsv_plan_destroy": [vp],
        "qsv_plan_num_steps": [vp, C.POINTER(sz)],
        "qsv_plan_get_step": [vp, sz, C.POINTER(i32), C.POINTER(u32), C.POINTER(C.c_uint8), sz],
        "qsv_plan_get_layout": [vp, i32, C.POINTER(C.c_uint8), sz],
        "qsv_get_layout": [vp, C.POINTER(C.c_uint8), sz],
        "qsv_plan_stats": [vp, C.POINTER(QsvStats)],
        "qsv_plan_initial_amplitudes": [vp, u64, dp, sz],
        "qsv_plan_serialize": [vp, vp, sz, C.POINTER(sz)],
        "qsv_run_plan": [vp, vp, C.POINTER(QsvStats)],
        "qsv_sample": [vp, dp, u64, C.POINTER(u64)],
        "qsv_norm_sqr": [vp, dp],
        "qsv_save": [vp, C.c_char_p],
        "qsv_load": [vp, C.c_char_p],
        "qsv_synchronize": [vp],
        "qsv_last_step_ms": [vp, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_size_t)],
        "qsv_device_pointer": [vp, C.POINTER(vp), C.POINTER(vp)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.qsv_last_error.argtypes = [vp]
    lib.qsv_last_error.restype = C.c_char_p
    lib.qsv_plan_last_error.argtypes = []
    lib.qsv_plan_last_error.restype = C.c_char_p
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "qsv_create", "qsv_create_sharded", "qsv_create_multi", "qsv_nccl_unique_id", "qsv_peer_export", "qsv_peer_import", "qsv_destroy", "qsv_last_error", "qsv_set_option",
    "qsv_get_info", "qsv_init_basis", "qsv_upload", "qsv_download", "qsv_gather", "qsv_apply", "qsv_plan_create",
    "qsv_plan_create_ex", "qsv_plan_num_steps", "qsv_plan_get_step", "qsv_plan_get_layout", "qsv_get_layout",
    "qsv_plan_destroy", "qsv_plan_initial_amplitudes", "qsv_plan_stats", "qsv_plan_serialize", "qsv_plan_last_error", "qsv_run_plan", "qsv_sample",
    "qsv_norm_sqr", "qsv_save", "qsv_load", "qsv_synchronize", "qsv_last_step_ms", "qsv_device_pointer",
]


def check(code, handle=None):
    if code != 0:
        lib = load_library()
        msg = lib.qsv_last_error(handle)
        raise QsvError(code, msg.decode() if msg else "")


def check_plan(code, lib=None):
    """`lib`: the library that built the plan (the host-emulation harness of the tests keeps its own error text)."""
    if code != 0:
        lib = lib or load_library()
        lib.qsv_plan_last_error.restype = C.c_char_p
        msg = lib.qsv_plan_last_error()
        raise QsvError(code, msg.decode() if msg else "")
