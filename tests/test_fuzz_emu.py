"""A fixed slice of the CPU fuzzer (tools/fuzz_emu.py): random circuits over every standard gate kind plus Custom closures
of six shapes (dense unitary, phased permutation, multi-controlled 2x2, partial with None, non-unitary, multi-controlled
flip on an arbitrary control pattern up to 10 wires), scheduled by plan.cpp and run through the host emulation of the
kernel's per-thread code: one device (uploaded register, |0..0>, folded prefix from a random basis state, forced tile
shapes, one gate per pass) and 2/4/8 emulated ranks (basis state, uploaded register, pipelined remaps), against the oracle.
The long campaigns (tens of thousands of seeds, also in TMA mode and at n up to 17) are run by hand with the tool."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_emu  # noqa: E402
from helpers import emu_lib  # noqa: E402


@pytest.mark.parametrize("first", [0, 20000, 20020])
def test_fuzz_slice(first):
    for seed in range(first, first + 20):
        fuzz_emu.one(seed)


def test_fuzz_slice_tma_mode():
    lib = emu_lib()
    lib.qsv_emu_set_tma_mode(1)
    try:
        for seed in range(30000, 30020):
            fuzz_emu.one(seed)
    finally:
        lib.qsv_emu_set_tma_mode(0)


def test_oracle_modes_agree_under_fuzz():
    """The checker checked: the faithful restatement (per-gate hash-map rebuild, first-assign / then-add, None overwrite;
    simulation.rs:64-135) and the dense one agree on the fuzzer's circuits, Custom closures with None results and
    non-unitary images included."""
    import numpy as np
    from helpers import encode_gates, orc
    for seed in range(90000, 90060):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(1, 9))
        c = fuzz_emu.build(rng, n, int(rng.integers(1, 60)), p_custom=0.4)
        enc = encode_gates(c.circuit_gates, n)
        reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        reg /= np.linalg.norm(reg)
        a = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
        b = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="faithful")
        assert np.max(np.abs(a - b)) <= 1e-11 * max(1.0, float(np.max(np.abs(a)))), seed


def test_plan_api_survives_garbage_ops():
    """qsv_plan_create on malformed qsv_op arrays (kinds out of range, wires beyond the register, repeated or missing
    controls, NaN / inf angles, reserved fields set, wrong control counts, absurd tile sizes): an error code, never a
    crash; whatever it accepts also executes."""
    import ctypes as C

    import numpy as np
    from helpers import F
    lib = emu_lib()
    rng = np.random.default_rng(1)
    codes = {}
    for _ in range(3000):
        n = int(rng.integers(1, 12))
        m = int(rng.integers(1, 6))
        ops = (F.QsvOp * m)()
        keep = []
        for i in range(m):
            o = ops[i]
            o.kind = int(rng.choice([rng.integers(0, 26), rng.integers(0, 2 ** 31)], p=[0.9, 0.1]))
            o.target = int(rng.choice([rng.integers(0, n), rng.integers(0, 40)], p=[0.8, 0.2]))
            nc = int(rng.choice([0, 1, 2, 3, rng.integers(0, 6)]))
            o.n_controls = nc
            if nc and rng.random() < 0.95:
                arr = (C.c_uint32 * nc)(*[int(rng.choice([rng.integers(0, n), rng.integers(0, 40)], p=[0.85, 0.15])) for _ in range(nc)])
                keep.append(arr)
                o.controls = arr
            o.param = float(rng.choice([rng.normal(), np.nan, np.inf, 1e308]))
            o.iparam = int(rng.integers(-5, 70))
            if o.iparam == 1:
                o.iparam = 0  # compact columns promise a buffer of a given size: not garbage the library could detect
            if rng.random() < 0.1:
                o.reserved = 1
            if (o.kind == F.GATE_CUSTOM or rng.random() < 0.05) and nc + 1 <= 6 and rng.random() < 0.9:
                mat = np.ascontiguousarray(rng.normal(size=(1 << (nc + 1), 1 << (nc + 1), 2)))
                keep.append(mat)
                o.matrix = mat.ctypes.data_as(C.POINTER(C.c_double))
                if rng.random() < 0.5:
                    nm = np.ascontiguousarray((rng.random(1 << (nc + 1)) < 0.5).astype(np.uint8))
                    keep.append(nm)
                    o.none_mask = nm.ctypes.data_as(C.POINTER(C.c_uint8))
        plan = C.c_void_p()
        nl = max(1, n - int(rng.integers(0, 3)))
        rc = lib.qsv_plan_create(C.byref(plan), n, nl, ops, m, int(rng.choice([0, 4, 6, 8, 13, 99])), int(rng.choice([0, 1, 3, 9])),
                                 int(rng.integers(0, 2)))
        codes[rc] = codes.get(rc, 0) + 1
        if rc == 0:
            if nl == n:
                amps = np.zeros(1 << lib.qsv_emu_alloc_qubits(plan), dtype=np.complex128)
                amps[0] = 1
                assert lib.qsv_emu_run_plan(plan, amps.ctypes.data_as(C.POINTER(C.c_double)), 0) == 0
            lib.qsv_plan_destroy(plan)
        else:
            assert rc in (F.ERR_INVALID_ARG, F.ERR_UNSUPPORTED) and lib.qsv_plan_last_error()
    assert codes.get(0, 0) > 20 and codes.get(F.ERR_INVALID_ARG, 0) > 1000
