"""The pipelined TMA kernel's host-visible logic on the CPU (quantr_b200/csrc/pass_kernel_tma.cu, tma_tile.h):

* the tile -> tensor-map description (dimensions cut at the tile's bit segments, boxes, coordinates) and the 128-byte
  swizzle, through a software model of a tiled-mode box copy;
* the external-phase tables (two half-index tables per DIAG op, multiplied per tile);
* the scaled last round + box store of passes that do not store from registers;
* the fused basis initialisation (zero tiles written from a zeroed buffer).

The host emulation (tests/emu) walks the passes in "TMA mode" with the kernel's own __host__ __device__ code and is
compared with the CPU oracle; the GPU suite (tests/test_gpu_kernel_variants.py) forces the real kernel onto the same sizes.
"""
import ctypes as C

import numpy as np
import pytest

from golden import reference_vectors as rv
from helpers import (OracleCircuit, emu_lib, emu_simulate, emu_simulate_sharded, encode_gates, orc, qb, qft_circuit, qft_expected,
                     random_any_gate_circuit, random_layered_circuit, st)

G = qb.Gate


@pytest.fixture(autouse=True)
def tma_mode():
    lib = emu_lib()
    lib.qsv_emu_set_tma_mode(1)
    yield lib
    lib.qsv_emu_set_tma_mode(0)


@pytest.mark.parametrize("seed", range(12))
def test_random_circuits_through_tma_path(seed, tma_mode):
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(8, 15))
    c = random_any_gate_circuit(OracleCircuit, G, n, int(rng.integers(20, 120)), rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=2)
    walked = 0
    for tile_bits, low_bits in ((6, 3), (7, 3), (8, 4), (10, 3), (11, 3)):
        if tile_bits > n:
            continue
        out = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=low_bits)
        assert np.max(np.abs(out - ref)) < 1e-12, (n, tile_bits)
        walked += tma_mode.qsv_emu_tma_passes()
    assert walked > 0  # the TMA walk was really taken


def test_scattered_tiles_need_several_boxes(tma_mode):
    """Gates on far-apart wires: the tile has more bit segments than a tensor map has dimensions, so it is moved as
    several boxes whose upper bits fold into the last dimension's coordinate."""
    n = 16
    c = OracleCircuit.new(n)
    for rep in range(3):
        for w in (0, 2, 4, 6, 8, 10, 15):
            c.add_gate(G.Rx(0.1 + 0.2 * w + rep), w)
            c.add_gate(G.CRk(2 + rep, (w + 5) % n), w)
        c.add_gate(G.Toffoli(0, 8), 4)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=2)
    out = emu_simulate(n, enc, None, tile_bits=11, low_bits=3)
    assert np.max(np.abs(out - ref)) < 1e-12
    assert tma_mode.qsv_emu_tma_passes() > 0 and tma_mode.qsv_emu_tma_max_boxes() > 1


@pytest.mark.parametrize("n,tile_bits", [(13, 11), (14, 12), (17, 11)])
def test_qft_closed_form_through_tma_path(n, tile_bits, tma_mode):
    x = 0x12345 & ((1 << n) - 1)
    enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[x] = 1
    out = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=3)
    assert np.max(np.abs(out - qft_expected(n, x))) < 1e-13
    assert tma_mode.qsv_emu_tma_passes() >= 2


def test_custom_gates_and_none_rule_through_tma_path(tma_mode):
    c = rv.build_x3sudoko(OracleCircuit, G, st)
    enc = encode_gates(c.circuit_gates, 10)
    ref = orc.simulate(10, enc.ops, enc.n_ops, None, mode="dense")
    for tile_bits in (7, 8, 10):
        out = emu_simulate(10, enc, None, tile_bits=tile_bits, low_bits=3)
        assert np.max(np.abs(out - ref)) < 1e-12


def test_layered_circuit_through_tma_path(tma_mode):
    c = random_layered_circuit(OracleCircuit, G, 14, 6, seed=30)
    enc = encode_gates(c.circuit_gates, 14)
    ref = orc.simulate(14, enc.ops, enc.n_ops, None, mode="dense", threads=2)
    out = emu_simulate(14, enc, None, tile_bits=11, low_bits=3)
    assert np.max(np.abs(out - ref)) < 1e-12


@pytest.mark.parametrize("mode", [1, 2])
def test_fused_initialisation_through_tma_path(mode, tma_mode):
    lib = tma_mode
    lib.qsv_emu_run_plan_fused_init.restype = C.c_int
    lib.qsv_emu_run_plan_fused_init.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64, C.c_uint64, C.c_uint32]
    rng = np.random.default_rng(77 + mode)
    for n, tile_bits in ((12, 8), (13, 11), (14, 11)):
        for trial in range(2):
            x = int(rng.integers(0, 1 << n))
            c = qft_circuit(OracleCircuit, G, n) if trial == 0 else random_any_gate_circuit(OracleCircuit, G, n, 40, rng)
            enc = encode_gates(c.circuit_gates, n)
            reg = np.zeros(1 << n, dtype=np.complex128)
            reg[x] = 1.0
            ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=2)
            plan = qb.Plan(n, enc, tile_bits=tile_bits, low_bits=3, lib=lib)
            amps = np.full(1 << n, np.nan + 1j * np.nan, dtype=np.complex128)  # the first pass must not read the register
            assert lib.qsv_emu_run_plan_fused_init(plan.handle, amps.ctypes.data_as(C.POINTER(C.c_double)), 0, x, mode) == 0
            assert np.max(np.abs(amps - ref)) < 1e-12
            plan.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_plans_through_tma_path(world, tma_mode):
    """Rank bits enter the external-phase tables (table A carries them) and the control predicates."""
    rng = np.random.default_rng(world)
    n = 13
    c = random_any_gate_circuit(OracleCircuit, G, n, 80, rng)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=2)
    out, _, n_exch = emu_simulate_sharded(n, enc, world, tile_bits=8, low_bits=3)
    assert np.max(np.abs(out - ref)) < 1e-12
    x = 1234
    enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    out, _, n_exch = emu_simulate_sharded(n, enc, world, basis_index=x, tile_bits=8, low_bits=3)
    assert n_exch == 0 and np.max(np.abs(out - qft_expected(n, x))) < 1e-13  # the first stages are folded into the ranks' initial amplitudes
