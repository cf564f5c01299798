"""Pins the CPU oracle (oracle/quantr_oracle.cpp) against the reference's own golden vectors.

Reference: src/circuit.rs:603-982 (22 vectors), tests/qft.rs:22-47, tests/grovers.rs:22-155.
"""
import numpy as np
import pytest

from golden import reference_vectors as rv
from helpers import FaithfulOracleCircuit, OracleCircuit, encode_gates, orc, qb, qft_circuit, qft_expected, st

G = qb.Gate


@pytest.mark.parametrize("vec", rv.VECTORS, ids=[v["name"] for v in rv.VECTORS])
@pytest.mark.parametrize("cls", [OracleCircuit, FaithfulOracleCircuit], ids=["dense", "faithful"])
def test_oracle_reproduces_reference_golden_vector(vec, cls):
    circuit = vec["build"](cls, G, st)
    amps = circuit.simulate().get_state().take().get_amplitudes()
    expect = np.array(vec["expect"])
    assert amps.shape == expect.shape
    assert np.max(np.abs(amps - expect)) < min(vec["tol"], 1e-12)  # the reference asserts tol; the goldens are exact


def test_grovers_3qubit_measure_all_thresholds():
    """tests/grovers.rs:62-69: 500 shots, |011> and |111> each > 200, everything else exactly 0."""
    sim = rv.build_grovers_3qubit(FaithfulOracleCircuit, G, st).simulate()
    bins = sim.measure_all(500).take()
    for state, count in bins.items():
        if state.to_string() in ("011", "111"):
            assert count > 200
        else:
            assert count == 0
    assert sum(bins.values()) == 500


def test_example_grovers_config1_probabilities():
    """examples/grovers.rs (BASELINE config 1): |amp|^2 = 0.5 on |110> and |111> (QUICK_START.md:85-86)."""
    amps = rv.build_example_grovers(OracleCircuit, G, st).simulate().get_state().take().get_amplitudes()
    p = np.abs(amps) ** 2
    assert abs(p[0b110] - 0.5) < 1e-12 and abs(p[0b111] - 0.5) < 1e-12


def test_x3sudoko_distribution():
    """tests/grovers.rs:143-152: the six solution prefixes dominate (p = 0.1077 each)."""
    sim = rv.build_x3sudoko(OracleCircuit, G, st).simulate()
    amps = sim.get_state().take().get_amplitudes()
    p = np.abs(amps) ** 2
    assert abs(p.sum() - 1.0) < 1e-12
    prefix = p.reshape(64, 16).sum(axis=1)
    sols = [int(s, 2) for s in rv.SUDOKU_SOLUTIONS]
    for s in sols:
        assert abs(prefix[s] - 0.1077) < 5e-4
    others = np.delete(prefix, sols)
    assert others.max() < 0.0062
    bins = sim.measure_all(5000).take()
    for state, count in bins.items():
        if state.to_string()[:6] in rv.SUDOKU_SOLUTIONS:
            assert count > 150
        else:
            assert count < 150


def test_faithful_and_dense_agree_on_x3sudoko_with_closure_callback():
    """The faithful oracle calls the Custom closure once per basis state, like simulation.rs:137-156."""
    circuit = rv.build_x3sudoko(OracleCircuit, G, st)
    gates = list(circuit.circuit_gates)
    enc = encode_gates(gates, 10)
    customs = [g for g in gates if g.kind != 0]

    def callback(op_index, qubits):
        gate = customs[op_index]
        res = gate.func(st.ProductState(qubits))
        return None if res is None else st.into_super_position(res).get_amplitudes()

    a = orc.simulate(10, enc.ops, enc.n_ops, None, mode="faithful", custom_callback=callback)
    b = orc.simulate(10, enc.ops, enc.n_ops, None, mode="dense")
    assert np.max(np.abs(a - b)) < 1e-15


def test_none_overwrite_rule():
    """SURVEY.md App. B.4 / simulation.rs:120-133: untouched states overwrite accumulated images."""
    def closure(prod):
        if prod.get_qubits()[0] == st.Qubit.Zero:
            return None
        return st.SuperPosition.new_with_amplitudes_unchecked([np.sqrt(0.5), np.sqrt(0.5)])

    for cls in (OracleCircuit, FaithfulOracleCircuit):
        c = cls.new(1)
        c.add_gate(G.H, 0).add_gate(G.Custom(closure, [], "N"), 0)
        amps = c.simulate().get_state().take().get_amplitudes()
        assert np.allclose(amps, [np.sqrt(0.5), 0.5], atol=1e-15)


def test_post_select_non_unitary_and_failed_collapse():
    """examples/post_select.rs:39-49: |0> -> sqrt(2)|0>, |1> -> 0; measure returns None when u >= total."""
    def post_select(prod):
        if prod.get_qubits()[0] == st.Qubit.Zero:
            return st.SuperPosition.new_with_amplitudes_unchecked([np.sqrt(2.0), 0.0])
        return st.SuperPosition.new_with_amplitudes_unchecked([0.0, 0.0])

    c = OracleCircuit.new(2)
    c.add_gate(G.H, 0).add_gate(G.H, 1).add_gate(G.Custom(post_select, [], "P"), 1)
    amps = c.simulate().get_state().take().get_amplitudes()
    assert np.allclose(amps, [np.sqrt(0.5), 0, np.sqrt(0.5), 0], atol=1e-15)
    half = np.array([0.5, 0, 0, 0], dtype=np.complex128)  # total probability 0.25
    idx = orc.measure_all(2, half, np.array([0.1, 0.2499, 0.25, 0.9]), cdf=False)
    assert list(idx[:2]) == [0, 0] and all(int(i) == (1 << 64) - 1 for i in idx[2:])
    assert list(orc.measure_all(2, half, np.array([0.1, 0.2499, 0.25, 0.9]), cdf=True)) == list(idx)


@pytest.mark.parametrize("n", [3, 5, 7, 10])
def test_qft_closed_form(n):
    """amp[y] = 2^{-n/2} exp(2 pi i x bitrev(y) / 2^n)."""
    x = 0xACE1 & ((1 << n) - 1)
    for cls in (OracleCircuit, FaithfulOracleCircuit):
        amps = qft_circuit(cls, G, n, x).simulate().get_state().take().get_amplitudes()
        assert np.max(np.abs(amps - qft_expected(n, x))) < 1e-13


def test_measure_strict_inequality_and_order():
    """super_positions.rs:332-342: first i with u < cumulative (strict)."""
    amps = np.array([0.5, 0.5, 0.5, 0.5], dtype=np.complex128)
    u = np.array([0.0, 0.2499999, 0.25, 0.5, 0.74, 0.75, 0.999999])
    assert list(orc.measure_all(2, amps, u, cdf=False)) == [0, 0, 1, 2, 2, 3, 3]
    assert list(orc.measure_all(2, amps, u, cdf=True)) == [0, 0, 1, 2, 2, 3, 3]
