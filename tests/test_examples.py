"""The reference's example programs that are not already golden vectors, replayed through the host mirror on the oracle and
on the scheduler + kernel emulation: examples/custom_gate.rs (CCC-not as a Custom gate), examples/generalised_control_not_gate.rs
(multicnot::<6>).  examples/qft.rs, grovers.rs and post_select.rs are covered by tests/golden and the parity suites."""
import numpy as np
import pytest

from helpers import EmuCircuit, FaithfulOracleCircuit, OracleCircuit, qb, st

G = qb.Gate
Z, O = st.Qubit.Zero, st.Qubit.One


def cccnot(input_state):  # examples/custom_gate.rs:48-66
    state = list(input_state.get_qubits())
    if state == [O, O, O, Z]:
        return st.ProductState.new([O] * 4).into_super_position()
    if state == [O, O, O, O]:
        return st.ProductState.new([O, O, O, Z]).into_super_position()
    return None


def multicnot(n):  # examples/generalised_control_not_gate.rs:53-68
    def f(input_state):
        copy_state = input_state.clone()
        if list(copy_state.get_qubits()) == [O] * n:
            copy_state.get_mut_qubits()[n - 1] = Z
            return copy_state.into_super_position()
        if list(copy_state.get_qubits()) == [O] * (n - 1) + [Z]:
            copy_state.get_mut_qubits()[n - 1] = O
            return copy_state.into_super_position()
        return None
    return f


@pytest.mark.parametrize("C", [OracleCircuit, FaithfulOracleCircuit, EmuCircuit])
def test_custom_gate_example(C, capsys):
    qc = C.new(4)
    qc.add_repeating_gate(G.X, [0, 1, 2]).add_gate(G.Custom(cccnot, [0, 1, 2], "X"), 3)  # custom_gate.rs:19-25
    qc.set_print_progress(True)
    simulated = qc.simulate()
    bins = simulated.measure_all(50).take()
    assert {k.to_string(): v for k, v in bins.items()} == {"1111": 50}
    amps = simulated.get_state().take().get_amplitudes()
    assert abs(amps[0b1111] - 1.0) < 1e-15 and np.count_nonzero(amps) == 1


@pytest.mark.parametrize("C", [OracleCircuit, FaithfulOracleCircuit, EmuCircuit])
def test_generalised_control_not_gate_example(C):
    n = 6
    qc = C.new(n)
    qc.add_repeating_gate(G.X, [0, 1, 2, 3, 4, 5]).add_gate(G.Custom(multicnot(n), [0, 1, 2, 3, 4], "X"), 5)  # :24-35
    simulated = qc.simulate()
    bins = simulated.measure_all(50).take()
    assert {k.to_string(): v for k, v in bins.items()} == {"111110": 50}
