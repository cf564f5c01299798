"""The reference's own golden state vectors, restated as data.

Every entry cites the reference test it comes from (paths relative to /root/reference).
`build(Circuit, Gate, st)` constructs the circuit through the public builder API exactly as
the reference test does; `expect` is the hand-computed register asserted there; `tol` is
the reference's own tolerance (ERROR_MARGIN).  Nothing here reads /root/reference at run time.
"""
import math

S2 = math.sqrt(0.5)  # std::f64::consts::FRAC_1_SQRT_2
PI = math.pi


# Custom closures used by the reference tests -----------------------------------------------------

def example_cnot(st):
    """src/circuit.rs:504-512"""
    def f(prod):
        q = prod.get_qubits()
        if q[0] == st.Qubit.Zero:
            return None
        if q[1] == st.Qubit.Zero:
            return st.SuperPosition.new_with_amplitudes([0, 0, 0, 1])
        return st.SuperPosition.new_with_amplitudes([0, 0, 1, 0])
    return f


def make_qft_closure(Circuit, Gate):
    """tests/qft.rs:51-67 — the closure itself builds and simulates a sub-circuit."""
    def qft(input_state):
        qubit_num = input_state.num_qubits()
        mini = Circuit.new(qubit_num)
        for pos in range(qubit_num):
            mini.add_gate(Gate.H, pos)
            for k in range(2, qubit_num - pos + 1):
                mini.add_gate(Gate.CRk(k, pos + k - 1), pos)
        mini.change_register(input_state)
        return mini.simulate().take_state().take()
    return qft


def make_multicnot(st, num_control):
    """tests/grovers.rs:157-172 (multicnot::<NUM_CONTROL>)"""
    def multicnot(input_state):
        copy_state = input_state.clone()
        q = copy_state.get_qubits()
        ones = [st.Qubit.One] * num_control
        almost = list(ones)
        almost[num_control - 1] = st.Qubit.Zero
        if list(q) == ones:
            copy_state.get_mut_qubits()[num_control - 1] = st.Qubit.Zero
            return copy_state
        if list(q) == almost:
            copy_state.get_mut_qubits()[num_control - 1] = st.Qubit.One
            return copy_state
        return None
    return multicnot


# Golden vectors ----------------------------------------------------------------------------------

def _v(name, ref, n, tol, build, expect):
    return {"name": name, "ref": ref, "n": n, "tol": tol, "build": build, "expect": [complex(x) for x in expect]}


def _swap_and_conjugate(C, G, st):
    c = C.new(2)
    c.add_gates([G.H, G.H]).add_gates([G.S, G.Sdag])
    return c


def _t_and_conjugate(C, G, st):
    c = C.new(2)
    c.add_gates([G.H, G.H]).add_gates([G.T, G.Tdag])
    return c


def _custom_gates(C, G, st):
    c = C.new(3)
    c.add_gate(G.H, 2).add_gate(G.Custom(example_cnot(st), [2], "cNot"), 1)
    return c


def _toffoli_gates(C, G, st):
    c = C.new(4)
    c.add_gate(G.X, 0).add_gate(G.H, 3).add_gate(G.Y, 3).add_gate(G.Toffoli(3, 0), 1)
    return c


def _three_pauli(C, G, st):
    c = C.new(4)
    c.add_gates([G.Z, G.Y, G.H, G.X])
    return c


def _hash_map_two(C, G, st):
    c = C.new(3)
    c.add_gates_with_positions({0: G.X, 2: G.H})
    return c


def _two_hadamard(C, G, st):
    c = C.new(2)
    c.add_gates([G.H, G.H])
    return c


def _two_rows(C, G, st):
    c = C.new(4)
    c.add_gates_with_positions({0: G.X}).add_gates_with_positions({3: G.X, 2: G.H})
    return c


def _cy_swap(C, G, st):
    c = C.new(4)
    c.add_repeating_gate(G.X, [1, 2]).add_gate(G.CY(2), 0).add_gate(G.Swap(3), 2).add_gate(G.CY(0), 3)
    return c


def _cz_swap(C, G, st):
    c = C.new(3)
    c.add_repeating_gate(G.X, [0, 2]).add_gate(G.Swap(1), 2).add_gate(G.CZ(1), 0)
    return c


def _cnot_simple(C, G, st):
    c = C.new(2)
    c.add_gate(G.H, 0).add_gate(G.CNot(1), 0)
    return c


def _cnot_flipped(C, G, st):
    c = C.new(2)
    c.add_gate(G.H, 0).add_gate(G.CNot(0), 1)
    return c


def _cnot_asym(C, G, st):
    c = C.new(4)
    c.add_gate(G.H, 1).add_gate(G.CNot(1), 3).add_gate(G.Y, 1)
    return c


def _hh_then(gate_fn):
    def build(C, G, st):
        c = C.new(2)
        c.add_gates([G.H, G.H])
        gate_fn(c, G)
        return c
    return build


def _cr(C, G, st):
    c = C.new(3)
    c.add_gates([G.X, G.X, G.X]).add_gate(G.CR(-PI * 0.5, 2), 1)
    return c


def _crk(C, G, st):
    c = C.new(3)
    c.add_gates([G.X, G.X, G.X]).add_gate(G.CRk(2, 2), 1)
    return c


def _custom_register(C, G, st):
    c = C.new(3)
    reg = st.ProductState.new_unchecked([st.Qubit.One, st.Qubit.Zero, st.Qubit.One])
    c.add_gate(G.X, 1).change_register(reg)
    return c


def _simple_qft(C, G, st):
    c = C.new(3)
    c.add_repeating_gate(G.X, [1, 2]).add_gate(G.Custom(make_qft_closure(C, G), [0, 1], "QFT"), 2)
    return c


def build_grovers_3qubit(C, G, st):
    """tests/grovers.rs:22-43"""
    c = C.new(3)
    c.add_repeating_gate(G.H, [0, 1, 2])
    c.add_gate(G.CZ(1), 2)
    (c.add_repeating_gate(G.H, [0, 1, 2]).add_repeating_gate(G.X, [0, 1, 2]).add_gate(G.H, 2)
      .add_gate(G.Toffoli(0, 1), 2).add_gate(G.H, 2).add_repeating_gate(G.X, [0, 1, 2]).add_repeating_gate(G.H, [0, 1, 2]))
    return c


def build_example_grovers(C, G, st):
    """examples/grovers.rs:21-37 (BASELINE config 1): the CZ sits on wire 0."""
    c = C.new(3)
    c.add_repeating_gate(G.H, [0, 1, 2])
    c.add_gate(G.CZ(1), 0)
    (c.add_repeating_gate(G.H, [0, 1, 2]).add_repeating_gate(G.X, [0, 1, 2]).add_gate(G.H, 2)
      .add_gate(G.Toffoli(0, 1), 2).add_gate(G.H, 2).add_repeating_gate(G.X, [0, 1, 2]).add_repeating_gate(G.H, [0, 1, 2]))
    return c


def build_x3sudoko(C, G, st):
    """tests/grovers.rs:75-139 — 10 qubits, five 4-wire and one 6-wire Custom multi-CNOT."""
    m4, m6 = make_multicnot(st, 4), make_multicnot(st, 6)
    qc = C.new(10)
    qc.add_repeating_gate(G.H, [0, 1, 2, 3, 4, 5]).add_gate(G.X, 8).add_gate(G.X, 9).add_gate(G.H, 9)
    for rep in range(2):
        for i in range(3):
            qc.add_gate(G.Toffoli(i, i + 3), 8)
        qc.add_gate(G.Custom(m4, [0, 1, 2], "X"), 6)
        for i in range(3):
            qc.add_gate(G.CNot(i), 6)
        qc.add_gate(G.Custom(m4, [3, 4, 5], "X"), 7)
        for i in range(3, 6):
            qc.add_gate(G.CNot(i), 7)
        if rep == 0:
            qc.add_gate(G.Custom(m4, [6, 7, 8], "X"), 9)
    (qc.add_repeating_gate(G.H, [0, 1, 2, 3, 4, 5]).add_repeating_gate(G.X, [0, 1, 2, 3, 4, 5]).add_gate(G.H, 5)
       .add_gate(G.Custom(m6, [0, 1, 2, 3, 4], "X"), 5).add_gate(G.H, 5)
       .add_repeating_gate(G.X, [0, 1, 2, 3, 4, 5]).add_repeating_gate(G.H, [0, 1, 2, 3, 4, 5]))
    return qc


SUDOKU_SOLUTIONS = ["001100", "001010", "010100", "010001", "100010", "100001"]  # tests/grovers.rs:145

Z = 0j
VECTORS = [
    _v("swap_and_conjugate_gates", "src/circuit.rs:604-614", 2, 1e-6, _swap_and_conjugate, [0.5, -0.5j, 0.5j, 0.5]),
    _v("t_and_conjugate_gates", "src/circuit.rs:616-626", 2, 1e-6, _t_and_conjugate,
       [0.5, complex(0.5 * S2, -0.5 * S2), complex(0.5 * S2, 0.5 * S2), 0.5]),
    _v("custom_gates", "src/circuit.rs:629-641", 3, 1e-6, _custom_gates, [S2, Z, Z, S2, Z, Z, Z, Z]),
    _v("toffoli_gates", "src/circuit.rs:644-658", 4, 1e-6, _toffoli_gates,
       [Z] * 8 + [-1j * S2, Z, Z, Z] + [Z, 1j * S2, Z, Z]),
    _v("runs_three_pauli_gates_with_hadamard", "src/circuit.rs:689-701", 4, 1e-6, _three_pauli,
       [Z] * 4 + [Z, 1j * S2, Z, 1j * S2] + [Z] * 8),
    _v("hash_map_with_two_gates", "src/circuit.rs:704-713", 3, 1e-6, _hash_map_two, [Z, Z, Z, Z, S2, S2, Z, Z]),
    _v("two_hadamard_gates_work", "src/circuit.rs:723-731", 2, 1e-6, _two_hadamard, [0.5, 0.5, 0.5, 0.5]),
    _v("add_two_rows_single_gates", "src/circuit.rs:734-748", 4, 1e-6, _two_rows, [Z] * 8 + [Z, S2, Z, S2] + [Z] * 4),
    _v("cy_and_swap_gates_work", "src/circuit.rs:751-768", 4, 1e-6, _cy_swap, [Z] * 12 + [1, Z, Z, Z]),
    _v("cz_and_swap_gates_work", "src/circuit.rs:771-786", 3, 1e-6, _cz_swap, [Z] * 6 + [-1, Z]),  # the reference lists 16 entries for this 3-qubit test; only the first 8 are compared
    _v("cnot_gate_simply_use_works", "src/circuit.rs:789-802", 2, 1e-6, _cnot_simple, [S2, Z, S2, Z]),
    _v("cnot_gate_simply_flipped", "src/circuit.rs:805-818", 2, 1e-6, _cnot_flipped, [S2, Z, Z, S2]),
    _v("cnot_gate_extended_control_works_asymmetric", "src/circuit.rs:821-836", 4, 1e-6, _cnot_asym,
       [Z, -1j * S2, Z, Z, 1j * S2, Z, Z, Z] + [Z] * 8),
    _v("rx_gate", "src/circuit.rs:848-860", 2, 1e-6, _hh_then(lambda c, G: c.add_gate(G.Rx(PI), 0)), [-0.5j] * 4),
    _v("ry_gate", "src/circuit.rs:863-875", 2, 1e-6, _hh_then(lambda c, G: c.add_gate(G.Ry(PI), 0)), [-0.5, -0.5, 0.5, 0.5]),
    _v("rz_gate", "src/circuit.rs:878-890", 2, 1e-6, _hh_then(lambda c, G: c.add_gate(G.Rz(PI), 0)), [-0.5j, -0.5j, 0.5j, 0.5j]),
    _v("global_gate", "src/circuit.rs:893-905", 2, 1e-6, _hh_then(lambda c, G: c.add_gate(G.Phase(PI), 0)), [0.5j] * 4),
    _v("x90_and_mx90_gate", "src/circuit.rs:908-921", 2, 1e-6,
       _hh_then(lambda c, G: c.add_gate(G.MX90, 0).add_gate(G.X90, 1)), [0.5] * 4),
    _v("y90_and_my90_gate", "src/circuit.rs:924-937", 2, 1e-6,
       _hh_then(lambda c, G: c.add_gate(G.MY90, 0).add_gate(G.Y90, 1)), [-0.5, 0.5, 0.5, -0.5]),
    _v("cr_gate", "src/circuit.rs:940-952", 3, 1e-6, _cr, [Z] * 7 + [-1j]),
    _v("crk_gate", "src/circuit.rs:955-967", 3, 1e-6, _crk, [Z] * 7 + [1j]),
    _v("custom_register", "src/circuit.rs:970-982", 3, 1e-6, _custom_register, [Z] * 7 + [1]),
    _v("simple_qft", "tests/qft.rs:22-47", 3, 1e-8, _simple_qft,
       [S2 * 0.5, -S2 * 0.5, -0.5j * S2, 0.5j * S2, complex(-0.25, 0.25), complex(0.25, -0.25), complex(0.25, 0.25),
        complex(-0.25, -0.25)]),
    _v("grovers_3qubit", "tests/grovers.rs:22-60", 3, 1e-8, build_grovers_3qubit, [Z, Z, Z, -S2, Z, Z, Z, -S2]),
]

# circuit layout golden (src/circuit.rs:535-575): flat gate vector after push_multi_gates
LAYOUT_EXPECT = ["Id", "Id", "H", "CNot(2)", "Id", "Id", "Id", "CNot(0)", "Id", "Id", "H", "Id", "Toffoli(1, 2)", "Id", "Id",
                 "Id", "Id", "CNot(0)"]
