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
