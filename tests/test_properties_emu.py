"""Size-independent properties of the path (BASELINE north_star: the checks that still apply where no oracle can run),
exercised here on the scheduler + kernel emulation: a circuit followed by its inverse is the identity, the path is
linear in the register (non-unitary Custom gates included), unitary circuits keep the norm, sampling is monotone in the
uniform."""
import numpy as np
import pytest

from helpers import OracleCircuit, emu_simulate, emu_simulate_sharded, encode_gates, orc, qb, random_any_gate_circuit, st

G = qb.Gate

INVERSE = {"H": "H", "X": "X", "Y": "Y", "Z": "Z", "S": "Sdag", "Sdag": "S", "T": "Tdag", "Tdag": "T", "X90": "MX90",
           "MX90": "X90", "Y90": "MY90", "MY90": "Y90"}


def inverse_of(gate):
    """The gate that undoes `gate` (every standard gate of the reference has one in the set)."""
    name = gate.get_name() if hasattr(gate, "get_name") else None
    k = gate.kind
    F = qb._ffi
    simple = {getattr(F, "GATE_" + a.upper()): b for a, b in INVERSE.items()}
    if k in simple:
        return getattr(G, simple[k])
    if k == F.GATE_RX:
        return G.Rx(-gate.param)
    if k == F.GATE_RY:
        return G.Ry(-gate.param)
    if k == F.GATE_RZ:
        return G.Rz(-gate.param)
    if k == F.GATE_PHASE:
        return G.Phase(-gate.param)
    if k == F.GATE_CR:
        return G.CR(-gate.param, gate.controls[0])
    if k == F.GATE_CRK:
        return G.CR(-2.0 * np.pi / (2.0 ** gate.iparam), gate.controls[0])
    if k in (F.GATE_CZ, F.GATE_CNOT, F.GATE_SWAP, F.GATE_TOFFOLI):
        return gate
    if k == F.GATE_CY:
        return gate
    raise AssertionError(name)


@pytest.mark.parametrize("seed", range(6))
def test_circuit_then_inverse_is_identity(seed):
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(5, 13))
    c = random_any_gate_circuit(OracleCircuit, G, n, 80, rng)
    wires_gates = [(i % n, g) for i, g in enumerate(c.circuit_gates) if g.kind != qb._ffi.GATE_ID]
    full = OracleCircuit.new(n)
    for w, g in wires_gates:
        full.add_gate(g, w)
    for w, g in reversed(wires_gates):
        full.add_gate(inverse_of(g), w)
    enc = encode_gates(full.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    for tile_bits, low_bits in ((0, 0), (6, 2), (9, 3)):
        out = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=low_bits)
        assert np.max(np.abs(out - reg)) < 1e-12
    if n >= 6:
        out, _, _ = emu_simulate_sharded(n, enc, 4, register=reg, tile_bits=4, low_bits=1)
        assert np.max(np.abs(out - reg)) < 1e-12


@pytest.mark.parametrize("seed", range(4))
def test_path_is_linear_in_the_register(seed):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_emu
    rng = np.random.default_rng(7100 + seed)
    n = int(rng.integers(4, 11))
    c = fuzz_emu.build(rng, n, 60, p_custom=0.4)  # Custom closures with None results and non-unitary images included
    enc = encode_gates(c.circuit_gates, n)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    phi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    a, b = complex(rng.normal(), rng.normal()), complex(rng.normal(), rng.normal())
    lhs = emu_simulate(n, enc, a * psi + b * phi)
    rhs = a * emu_simulate(n, enc, psi) + b * emu_simulate(n, enc, phi)
    assert np.max(np.abs(lhs - rhs)) < 1e-10 * max(1.0, float(np.max(np.abs(lhs))))


@pytest.mark.parametrize("seed", range(4))
def test_unitary_circuits_keep_the_norm_and_sampling_is_monotone(seed):
    rng = np.random.default_rng(7200 + seed)
    n = int(rng.integers(4, 13))
    c = random_any_gate_circuit(OracleCircuit, G, n, 100, rng)
    enc = encode_gates(c.circuit_gates, n)
    out = emu_simulate(n, enc, None)
    assert abs(np.sum(np.abs(out) ** 2) - 1.0) < 1e-12
    u = np.sort(rng.random(500))
    idx = orc.measure_all(n, out, u)
    assert np.all(np.diff(idx.astype(np.int64)) >= 0)  # the inverse CDF of super_positions.rs:332-342 is monotone
