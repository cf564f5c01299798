"""Host-side logic that stays on the host in the reference too: builder layout, validation,
state types, gate-list encoding.  Reference: src/circuit.rs:124-341 and its tests :516-597."""
import numpy as np
import pytest

from golden import reference_vectors as rv
from helpers import encode_gates, qb, st
from quantr_b200 import _ffi as F

G, Circuit, QuantrError = qb.Gate, qb.Circuit, qb.QuantrError


def test_pushes_multi_gates():  # src/circuit.rs:535-551
    c = Circuit.new(3)
    c.add_gates([G.CNot(2), G.CNot(0), G.H]).add_gates([G.Toffoli(1, 2), G.H, G.CNot(0)])
    assert [repr(g) for g in c.get_gates()] == rv.LAYOUT_EXPECT


def test_pushes_multi_gates_using_vec():  # src/circuit.rs:554-575
    c = Circuit.new(3)
    c.add_gates_with_positions({2: G.H, 0: G.CNot(2), 1: G.CNot(0)})
    c.add_gates_with_positions({2: G.CNot(0), 0: G.Toffoli(1, 2), 1: G.H})
    assert [repr(g) for g in c.get_gates()] == rv.LAYOUT_EXPECT


def test_single_multi_gate_column_is_left_alone():
    c = Circuit.new(3)
    c.add_gate(G.CNot(0), 2)
    assert [repr(g) for g in c.get_gates()] == ["Id", "Id", "CNot(0)"]


@pytest.mark.parametrize("bad", [
    lambda c: c.add_gates([G.Id, G.Custom(rv.example_cnot(st), [1], "X"), G.Id]),  # circuit.rs:516-523
    lambda c: c.add_gates([G.CNot(0), G.Id, G.Id]),  # :525-532
    lambda c: c.add_gates_with_positions({2: G.H, 0: G.CNot(0), 1: G.CNot(0)}),  # :577-585
    lambda c: c.add_gates_with_positions({2: G.H, 0: G.CNot(2), 1: G.CNot(3)}),  # :587-597
    lambda c: c.add_gate(G.Custom(rv.example_cnot(st), [0], "NonAscii†"), 1),  # :839-845
    lambda c: c.add_gates([G.H, G.H]),
    lambda c: c.add_gate(G.H, 3),
    lambda c: c.add_gate(G.Toffoli(1, 1), 0),
])
def test_builder_rejects(bad):
    with pytest.raises(QuantrError):
        bad(Circuit.new(3))


def test_catches_repeating_positions():  # circuit.rs:715-720
    with pytest.raises(QuantrError):
        Circuit.new(4).add_repeating_gate(G.X, [0, 1, 1, 3])


def test_custom_register_wrong_dimension():  # circuit.rs:984-991
    c = Circuit.new(3)
    with pytest.raises(QuantrError):
        c.add_gate(G.X, 1).change_register(st.ProductState.new_unchecked([st.Qubit.One, st.Qubit.Zero]))


def test_zero_qubit_circuit_rejected():
    with pytest.raises(QuantrError):
        Circuit.new(0)


def test_product_state_bit_conventions():  # product_states.rs:275-320
    p = st.ProductState.binary_basis(5, 4)
    assert p.to_string() == "0101" and p.comp_basis() == 5
    assert st.ProductState.new([st.Qubit.One, st.Qubit.Zero]).comp_basis() == 2
    p.insert_qubits([st.Qubit.One, st.Qubit.Zero], [0, 3])
    assert p.to_string() == "1100"
    assert st.Qubit.One.kronecker_prod(st.Qubit.Zero).kronecker_prod(st.Qubit.One).to_string() == "101"
    with pytest.raises(QuantrError):
        st.ProductState.new([])
    with pytest.raises(QuantrError):
        p.invert_digit(4)
    assert p.invert_digit(1).to_string() == "1000"


def test_super_position_validation():  # super_positions.rs:403-544
    with pytest.raises(QuantrError):
        st.SuperPosition.new_with_amplitudes([1, 0, 0])
    with pytest.raises(QuantrError):
        st.SuperPosition.new_with_amplitudes([0.5, 0.5])
    sp = st.SuperPosition.new_with_amplitudes([0, 1j, 0, 0])
    assert sp.get_num_qubits() == 2 and sp.get_dimension() == 4
    assert sp.get_amplitude(1) == 1j and sp.get_amplitude(4) is None
    assert sp.get_amplitude_from_state(st.ProductState.binary_basis(1, 2)) == 1j
    assert sp.to_hash_map() == {st.ProductState.binary_basis(1, 2): 1j}
    items = list(sp)
    assert len(items) == 4 and items[1][0].to_string() == "01"  # zeros included, super_position_iter.rs:23-37
    hp = st.SuperPosition.new_with_hash_amplitudes({st.ProductState.binary_basis(2, 2): 1.0})
    assert hp.get_amplitudes()[2] == 1.0
    assert st.SuperPosition.new(2).get_amplitudes()[0] == 1.0


def test_super_position_reference_tests():  # super_positions.rs:403-544 replayed one by one
    r = float(np.sqrt(0.5))
    P = st.ProductState.new_unchecked
    Z, O = st.Qubit.Zero, st.Qubit.One
    amps = [0, r, -1j * r, 0]
    sp = st.SuperPosition.new_unchecked(2).set_amplitudes(amps)
    assert sp.get_amplitude_from_state(P([Z, O])) == r  # retrieve_amplitude_from_state
    assert sp.get_amplitude(2) == -1j * r  # retrieve_amplitude_from_list_pos
    states = {P([Z, O]): r, P([O, Z]): -1j * r}
    assert np.array_equal(st.SuperPosition.new_with_amplitudes(amps).get_amplitudes(),
                          st.SuperPosition.new_with_hash_amplitudes(states).get_amplitudes())  # sets_amplitude_from_states
    with pytest.raises(QuantrError, match="The first state has product dimension of 2, whilst the state, .101>, found as a key"):
        st.SuperPosition.new_with_hash_amplitudes({P([Z, O]): r, P([O, Z, O]): -1j * r})  # ..._wrong_dimension
    with pytest.raises(QuantrError, match="does not equal 1. That is, the superpositon does not conserve probability"):
        st.SuperPosition.new_with_hash_amplitudes({P([Z, O]): r, P([O, Z]): -0.5j * r})  # ..._breaks_probability
    with pytest.raises(QuantrError, match="Unable to retreive product state"):
        sp.get_amplitude_from_state(P([Z, O, O]))  # catches_retrieve_amplitude_from_wrong_state
    assert sp.get_amplitude(4) is None  # catches_retrieve_amplitude_from_wrong_list_pos (None.unwrap() panics)
    with pytest.raises(QuantrError, match="does not conserve probability"):
        st.SuperPosition.new_unchecked(2).set_amplitudes([0, 0.5, 0, -0.5j])  # catches_super_position_breaking_conservation
    with pytest.raises(QuantrError, match="has length 2, when it should have length 4"):
        st.SuperPosition.new_unchecked(2).set_amplitudes([1, 0])
    # the probability check comes before the length check (super_positions.rs:68-82)
    with pytest.raises(QuantrError, match="must be of the form 2..n"):
        st.SuperPosition.new_with_amplitudes([1, 0, 0])
    with pytest.raises(QuantrError, match="does not conserve probability"):
        st.SuperPosition.new_with_amplitudes([0.5, 0.5, 0.5])
    with pytest.raises(QuantrError, match="An empty HashMap was given"):
        st.SuperPosition.new_with_hash_amplitudes({})


def test_super_position_set_amplitudes_from_states_and_hash_map_margin():  # super_positions.rs:270-323
    sp = st.SuperPosition.new(2)
    assert sp.set_amplitudes_from_states({st.ProductState.new([st.Qubit.Zero, st.Qubit.One]): 1.0}) is sp
    assert np.array_equal(sp.get_amplitudes(), [0, 1, 0, 0])  # the doc test at :262-268
    with pytest.raises(QuantrError, match="An empty HashMap was given"):
        sp.set_amplitudes_from_states({})
    with pytest.raises(QuantrError, match="The first state has product dimension of 2"):
        sp.set_amplitudes_from_states({st.ProductState.binary_basis(1, 3): 1.0})
    with pytest.raises(QuantrError, match="does not equal 1"):
        sp.set_amplitudes_from_states({st.ProductState.binary_basis(1, 2): 0.5})
    assert st.SuperPosition.new(2).to_hash_map() == {st.ProductState.binary_basis(0, 2): 1.0}  # the doc test at :308-313
    # amplitudes whose square is within 1e-6 of zero are left out of the map (equal_within_error)
    tiny = st.SuperPosition.new_with_amplitudes_unchecked([1.0, 5e-4, 2e-3, 0.0])
    assert set(k.to_string() for k in tiny.to_hash_map()) == {"00", "10"}


def test_super_position_measure_rule():  # super_positions.rs:332-342: strict `<` on the running sum, None when it falls short
    sp = st.SuperPosition.new_with_amplitudes_unchecked([0.5, 0.5j, -0.5, 0.5])
    assert [sp.measure(u).to_string() for u in (0.0, 0.2499, 0.25, 0.6, 0.9999)] == ["00", "00", "01", "10", "11"]
    short = st.SuperPosition.new_with_amplitudes_unchecked([0.5, 0.0, 0.0, 0.5])
    assert short.measure(0.49).to_string() == "11" and short.measure(0.5) is None
    qb.seed(11)
    draws = [sp.measure().to_string() for _ in range(400)]
    assert all(0.15 < draws.count(k) / 400 < 0.35 for k in ("00", "01", "10", "11"))


def test_encoding_walks_gate_vector_like_simulation_rs():
    """simulation.rs:37-56: flat position -> wire = position mod n, Id skipped, order kept."""
    c = Circuit.new(3)
    c.add_gates([G.CNot(2), G.CNot(0), G.H]).add_gate(G.Rx(0.25), 1).add_gate(G.CRk(3, 0), 2)
    enc = encode_gates(c.get_gates(), 3)
    got = [(enc.ops[i].kind, enc.ops[i].target, [enc.ops[i].controls[j] for j in range(enc.ops[i].n_controls)])
           for i in range(enc.n_ops)]
    assert got == [(F.GATE_H, 2, []), (F.GATE_CNOT, 0, [2]), (F.GATE_CNOT, 1, [0]), (F.GATE_RX, 1, []), (F.GATE_CRK, 2, [0])]
    assert enc.ops[3].param == 0.25 and enc.ops[4].iparam == 3


def test_custom_expansion_matrix_and_none_mask():
    """Host expansion of a Custom closure: column s = closure(binary_basis(s))."""
    from quantr_b200.circuit import expand_custom
    m, none = expand_custom(G.Custom(rv.example_cnot(st), [2], "cNot"))
    assert list(none) == [1, 1, 0, 0]
    assert m[3, 2] == 1 and m[2, 3] == 1 and np.count_nonzero(m) == 2

    def bad(prod):
        return st.SuperPosition.new_with_amplitudes_unchecked([1, 0, 0, 0])
    with pytest.raises(QuantrError):
        expand_custom(G.Custom(bad, [], "bad"))


def test_structured_custom_gates_are_lowered_to_controlled_ops():
    """Permutation-with-phase closures (multi-controlled gates, tests/grovers.rs:157-172 multicnot::<N>) do not need the
    dense Custom round: x3sudoko schedules without one, a 13-wire multi-CNOT closure is passed as compact columns and acts
    like the reference's closure, and a 21-wire one - far beyond the 13-wire limit of dense Custom gates - is accepted."""
    import numpy as np
    from golden import reference_vectors as rv
    from helpers import EmuCircuit, OracleCircuit, emu_lib, emu_simulate, encode_gates, orc, qb
    from quantr_b200 import states as st
    G = qb.Gate
    lib = emu_lib()
    enc = encode_gates(rv.build_x3sudoko(OracleCircuit, G, st).circuit_gates, 10)
    desc = qb.Plan(10, enc, tile_bits=8, low_bits=3, lib=lib).describe()
    assert all(r["type"] == 0 for p in desc["passes"] for r in p["rounds"])  # no dense round
    ref = orc.simulate(10, enc.ops, enc.n_ops, None, mode="dense")
    assert np.max(np.abs(emu_simulate(10, enc, None, tile_bits=8, low_bits=3) - ref)) < 1e-12

    def multicnot(prod):  # the reference's multicnot::<N>: flips the last qubit when every other one is |1>
        q = prod.get_qubits()
        if all(b == st.Qubit.One for b in q[:-1]):
            amps = np.zeros(1 << len(q), dtype=np.complex128)
            idx = (1 << len(q)) - 2 if q[-1] == st.Qubit.One else (1 << len(q)) - 1
            amps[idx] = 1.0
            return st.SuperPosition.new_with_amplitudes_unchecked(amps)
        return None

    n = 13
    for pattern in ((1 << n) - 2, ((1 << n) - 2) & ~0b1000, 0b101):  # all controls set / one control clear / most clear
        c = OracleCircuit.new(n)
        for w in range(n):
            if (pattern >> (n - 1 - w)) & 1:
                c.add_gate(G.X, w)
        c.add_gate(G.H, 3)
        c.add_gate(G.Custom(multicnot, list(range(n - 1)), "mcx"), n - 1)  # 13 wires: compact columns
        enc = encode_gates(c.circuit_gates, n)
        assert enc.ops[enc.n_ops - 1].iparam == 1
        reg = np.zeros(1 << n, dtype=np.complex128)
        reg[0] = 1.0
        out = emu_simulate(n, enc, reg, tile_bits=8, low_bits=3)
        # expected: X pattern, H on wire 3, then the target flips on the branch where every control is 1
        exp = np.zeros(1 << n, dtype=np.complex128)
        b3 = 1 << (n - 1 - 3)
        for sign_idx, amp in ((pattern & ~b3, np.sqrt(0.5)), (pattern | b3, np.sqrt(0.5) * (-1.0 if (pattern & b3) else 1.0))):
            idx = sign_idx
            if (idx >> 1) == (1 << (n - 1)) - 1:
                idx ^= 1
            exp[idx] += amp
        assert np.max(np.abs(out - exp)) < 1e-14

    # 21 wires (20 controls): built directly as a qsv_op with compact columns (2 x 2^21 amplitudes)
    import ctypes as C
    from quantr_b200 import _ffi as F
    k, n = 21, 22
    dim = 1 << k
    none = np.ones(dim, dtype=np.uint8)
    none[dim - 2] = none[dim - 1] = 0
    cols = np.zeros((2, dim), dtype=np.complex128)
    cols[0, dim - 1] = 1.0  # |1..10> -> |1..11>
    cols[1, dim - 2] = 1.0
    ops = (F.QsvOp * 1)()
    ctrl = (C.c_uint32 * (k - 1))(*range(k - 1))
    ops[0].kind, ops[0].target, ops[0].n_controls, ops[0].controls, ops[0].iparam = F.GATE_CUSTOM, k - 1, k - 1, ctrl, 1
    ops[0].matrix = cols.ctypes.data_as(C.POINTER(C.c_double))
    ops[0].none_mask = none.ctypes.data_as(C.POINTER(C.c_uint8))
    enc = qb.EncodedOps(ops, [ctrl, cols, none], [])
    enc.n_ops = 1
    plan = qb.Plan(n, enc, lib=lib)
    d = plan.describe()
    assert d["n_lowered_ops"] == 1 and plan.stats()["n_passes"] == 1


def test_headline_plans_keep_their_shape():
    """Guards the schedules the measured numbers in DESIGN.md 6 belong to (host-only: plans are built, nothing runs):
    QFT-33 from a basis state = 17 stages folded into the initial amplitudes + one write-only pass + one pass over the
    contiguous low bits; QFT-36 on 8 ranks the same per rank and no remap; Grover-36 on 8 ranks = 3 remaps, each of
    which the pipelined exchange can run against both of its neighbouring passes."""
    import ctypes as C
    from helpers import emu_lib, encode_gates, qft_circuit
    from workloads import grover_circuit
    import quantr_b200 as qb
    lib = emu_lib()
    for n, nl in ((33, 33), (36, 33)):
        enc = encode_gates(qft_circuit(qb.Circuit, qb.Gate, n).get_gates(), n)
        plan = qb.Plan(n, enc, n_local=nl, free_layout=True, lib=lib)
        d = plan.describe()
        assert d["prefix_local_bits"] == 17 and d["prefix_ops"] == 2 * (n - nl + 17) - 1
        assert [k for k, _ in plan.steps()] == ["pass", "pass"]
        assert [len(p["rounds"]) for p in d["passes"]] == [2, 2]
        assert d["passes"][1]["tile"] == list(range(11))  # the contiguous low bits
        assert plan.layout(False) == list(range(n)) and plan.layout(True) == list(range(n))
        plan.close()
    c, _info = grover_circuit(qb.Circuit, qb.Gate, 36, iterations=1)
    plan = qb.Plan(36, encode_gates(c.get_gates(), 36), n_local=33, free_layout=True, lib=lib)
    steps = plan.steps()
    exchanges = [i for i, (k, _) in enumerate(steps) if k != "pass"]
    assert len(exchanges) == 3 and len(steps) - len(exchanges) <= 12
    assert any(r["type"] == 2 for p in plan.describe()["passes"] for r in p["rounds"])  # the Toffoli chains as permutation rounds
    lib.qsv_emu_overlap_group.restype = C.c_int
    sliceable = (C.c_uint8 * len(steps))(*[1 if k == "pass" else 0 for k, _ in steps])
    for i in exchanges:
        out = (C.c_uint32 * 6)()
        assert lib.qsv_emu_overlap_group(plan.handle, i, sliceable, 2, out) == 1
        assert out[0] == 1 and out[1] == 1 and out[2] == 2  # both neighbours sliced, four slices
    plan.close()
