"""Host scheduler + blob encoding + per-thread arithmetic, checked without a GPU.

The plan is produced by the product's scheduler (quantr_b200/csrc/plan.cpp) and executed by
tests/emu/qsv_emu.cpp, which drives the kernel's own __host__ __device__ per-thread functions
sequentially.  The checker is the CPU oracle.
"""
import numpy as np
import pytest

from golden import reference_vectors as rv
from helpers import (EmuCircuit, OracleCircuit, emu_simulate, encode_gates, orc, qb, qft_circuit, qft_expected,
                     random_any_gate_circuit, random_layered_circuit, st)
from quantr_b200 import _ffi as F

G = qb.Gate


@pytest.mark.parametrize("vec", rv.VECTORS, ids=[v["name"] for v in rv.VECTORS])
def test_emulated_plan_reproduces_golden_vector(vec):
    amps = vec["build"](EmuCircuit, G, st).simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps - np.array(vec["expect"]))) < 1e-12


@pytest.mark.parametrize("seed", range(12))
def test_random_circuits_match_oracle_over_tile_configs(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 14))
    c = random_any_gate_circuit(OracleCircuit, G, n, int(rng.integers(1, 120)), rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    for tile_bits, low_bits, fuse in [(0, 0, True), (6, 2, True), (8, 3, True), (13, 3, True), (5, 1, False), (12, 3, False)]:
        out = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=low_bits, fuse=fuse)
        assert np.max(np.abs(out - ref)) < 1e-12, (n, tile_bits, low_bits, fuse)


def test_layered_config3_generator_small():
    """BASELINE config 3's generator at a CPU-checkable size."""
    c = random_layered_circuit(OracleCircuit, G, 12, 6, seed=30)
    enc = encode_gates(c.circuit_gates, 12)
    ref = orc.simulate(12, enc.ops, enc.n_ops, None, mode="dense")
    out, desc = emu_simulate(12, enc, None, tile_bits=8, low_bits=3, describe=True)
    assert np.max(np.abs(out - ref)) < 1e-12
    assert abs(np.sum(np.abs(out) ** 2) - 1.0) < 1e-12
    assert len(desc["passes"]) < desc["n_gates"]  # fused


def test_x3sudoko_with_custom_gates():
    c = rv.build_x3sudoko(OracleCircuit, G, st)
    enc = encode_gates(c.circuit_gates, 10)
    ref = orc.simulate(10, enc.ops, enc.n_ops, None, mode="dense")
    for tile_bits in (0, 7, 8):
        out = emu_simulate(10, enc, None, tile_bits=tile_bits, low_bits=2)
        assert np.max(np.abs(out - ref)) < 1e-12


def test_none_overwrite_and_non_unitary_custom():
    def closure(prod):
        if prod.get_qubits()[0] == st.Qubit.Zero:
            return None
        return st.SuperPosition.new_with_amplitudes_unchecked([np.sqrt(0.5), np.sqrt(0.5)])

    c = EmuCircuit.new(1)
    c.add_gate(G.H, 0).add_gate(G.Custom(closure, [], "N"), 0)
    assert np.allclose(c.simulate().get_state().take().get_amplitudes(), [np.sqrt(0.5), 0.5], atol=1e-15)


@pytest.mark.parametrize("n,tile_bits,low_bits,expect_passes", [(13, 8, 3, 2), (16, 12, 3, 2), (33, 12, 3, 4), (33, 13, 3, 3), (36, 12, 3, 4)])
def test_qft_pass_counts(n, tile_bits, low_bits, expect_passes):
    """QFT-n: n H + n(n-1)/2 CRk fuse into ceil-ish(n / (T - L)) passes (SURVEY.md 7.2 hard part 1)."""
    c = qft_circuit(OracleCircuit, G, n)
    enc = encode_gates(c.circuit_gates, n)
    plan = qb.Plan(n, enc, tile_bits=tile_bits, low_bits=low_bits)
    stats = plan.stats()
    assert stats["n_gates"] == n + n * (n - 1) // 2
    assert stats["n_passes"] == expect_passes
    assert stats["bytes_per_pass"] == 32 << n
    desc = plan.describe()
    assert desc["n_lowered_ops"] <= 2 * n  # all CRk of one target merged into one diagonal


@pytest.mark.parametrize("n", [5, 9, 13])
def test_qft_closed_form_through_plan(n):
    x = 0x12345 & ((1 << n) - 1)
    amps = qft_circuit(EmuCircuit, G, n, x).simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps - qft_expected(n, x))) < 1e-13


def test_sharded_plan_rank_bits_feed_controls_and_phases():
    """Top log2(P) index bits live in the rank id: controls and diagonal phases on them must still act."""
    n, g = 10, 2
    rng = np.random.default_rng(5)
    c = OracleCircuit.new(n)
    for _ in range(60):
        wires = rng.permutation(n)
        t = int(wires[0]) if wires[0] >= g else int(wires[1] if wires[1] >= g else wires[2])
        ctrl = [int(w) for w in wires if w != t][:2]
        r = rng.integers(0, 6)
        gate = [G.H, G.Rx(0.3), G.CNot(ctrl[0]), G.Toffoli(ctrl[0], ctrl[1]), G.CRk(3, ctrl[0]), G.CZ(ctrl[0])][r]
        if r in (4, 5) and ctrl[0] < g:
            # diagonal gates may sit on global wires as long as nothing non-diagonal targets them
            pass
        c.add_gate(gate, t)
    c.add_gate(G.Rz(0.7), 0).add_gate(G.Z, 1).add_gate(G.CRk(2, 5), 1)  # diagonals targeting global wires
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    nl = n - g
    for rank in range(1 << g):
        shard = reg[rank << nl:(rank + 1) << nl]
        out = emu_simulate(n, enc, shard, tile_bits=6, low_bits=2, n_local=nl, rank=rank)
        assert np.max(np.abs(out - ref[rank << nl:(rank + 1) << nl])) < 1e-12


def test_plan_errors():
    enc = encode_gates(qft_circuit(OracleCircuit, G, 6).circuit_gates, 6)
    plan = qb.Plan(6, enc, n_local=4)  # H on a rank bit: the plan inserts a global-qubit remap
    assert any(kind == "exchange" for kind, _ in plan.steps())
    c = OracleCircuit.new(6)
    def increment(prod):  # a cyclic shift of the 64 basis states: dense as far as the scheduler is concerned
        q = prod.get_qubits()
        v = 0
        for b in q:
            v = (v << 1) | (1 if b == st.Qubit.One else 0)
        amps = np.zeros(1 << len(q), dtype=np.complex128)
        amps[(v + 1) % (1 << len(q))] = 1.0
        return st.SuperPosition.new_with_amplitudes_unchecked(amps)

    c.add_gate(G.Custom(increment, [0, 1, 2, 3, 4], "wide"), 5)
    with pytest.raises(F.QsvError) as e:
        qb.Plan(6, encode_gates(c.circuit_gates, 6), n_local=4)  # a dense 6-wire Custom gate cannot be made local on 4 bits
    assert e.value.code == 5
    c = OracleCircuit.new(6)
    c.add_gate(G.Custom(lambda p: None, [0, 1, 2, 3, 4], "nothing"), 5)  # the identity: lowered to no op at all
    assert qb.Plan(6, encode_gates(c.circuit_gates, 6), n_local=4).stats()["n_passes"] == 0
    with pytest.raises(F.QsvError) as e:
        qb.Plan(6, encode_gates(OracleCircuit.new(6).add_gate(G.Custom(increment, [0, 1, 2, 3, 4], "wide"), 5).circuit_gates, 6), n_local=4)
    assert e.value.code == 5
    bad = (F.QsvOp * 1)()
    bad[0].kind = F.GATE_CNOT
    bad[0].target = 1
    bad[0].n_controls = 0
    fake = qb.EncodedOps(bad, [], [])
    fake.n_ops = 1
    with pytest.raises(F.QsvError) as e:
        qb.Plan(3, fake)
    assert e.value.code == 1
    bad[0].kind = 99
    with pytest.raises(F.QsvError):
        qb.Plan(3, fake)


@pytest.mark.parametrize("n,tile_bits,low_bits", [(6, 5, 2), (9, 7, 3), (13, 12, 0), (14, 11, 3), (12, 8, 0)])
@pytest.mark.parametrize("mode", [1, 2])
def test_basis_initialisation_fused_into_the_first_pass(n, tile_bits, low_bits, mode):
    """Fused initialisation (pass_core.h PassInit): the first pass synthesises its tiles instead of reading the register
    (mode 2 also writes all-zero tiles without arithmetic).  The buffer starts as NaN to prove it is never read."""
    import ctypes as C
    from helpers import emu_lib
    lib = emu_lib()
    lib.qsv_emu_run_plan_fused_init.restype = C.c_int
    lib.qsv_emu_run_plan_fused_init.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64, C.c_uint64, C.c_uint32]
    rng = np.random.default_rng(n * 10 + mode)
    for trial in range(3):
        x = int(rng.integers(0, 1 << n))
        c = qft_circuit(OracleCircuit, G, n) if trial == 0 else random_any_gate_circuit(OracleCircuit, G, n, 40, rng)
        enc = encode_gates(c.circuit_gates, n)
        reg = np.zeros(1 << n, dtype=np.complex128)
        reg[x] = 1.0
        ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=2)
        plan = qb.Plan(n, enc, tile_bits=tile_bits, low_bits=low_bits, lib=lib)
        n_alloc = lib.qsv_emu_alloc_qubits(plan.handle)
        amps = np.full(1 << n_alloc, np.nan + 1j * np.nan, dtype=np.complex128)
        assert lib.qsv_emu_run_plan_fused_init(plan.handle, amps.ctypes.data_as(C.POINTER(C.c_double)), 0, x, mode) == 0
        assert np.max(np.abs(amps[:1 << n] - ref)) < 1e-12
        plan.close()


def test_peephole_optimiser_merges_and_cancels_single_qubit_gates():
    """H.H, X.X and Rx.Rx^-1 vanish (also across gates that commute with them); runs of rotations on one wire become one
    matrix.  Anchor: the reference walks every gate separately (src/circuit/simulation.rs:37-56)."""
    n = 6
    c = OracleCircuit.new(n)
    c.add_gate(G.H, 0).add_gate(G.X, 3).add_gate(G.CNot(1), 2).add_gate(G.H, 0).add_gate(G.X, 3)  # CNot(1->2) commutes with wires 0 and 3
    c.add_gate(G.Rx(0.4), 5).add_gate(G.Rz(0.3), 4).add_gate(G.Rx(-0.4), 5)
    enc = encode_gates(c.circuit_gates, n)
    desc = qb.Plan(n, enc, lib=__import__("helpers").emu_lib()).describe()
    assert desc["n_gates"] == 8 and desc["n_lowered_ops"] == 2  # CNot and Rz survive
    rng = np.random.default_rng(3)
    c2 = OracleCircuit.new(n)
    for _ in range(40):
        w = int(rng.integers(0, n))
        c2.add_gate([G.H, G.Rx(0.3), G.Ry(1.1), G.Rz(0.7), G.X, G.Y, G.T][int(rng.integers(0, 7))], w)
    enc2 = encode_gates(c2.circuit_gates, n)
    d2 = qb.Plan(n, enc2, lib=__import__("helpers").emu_lib()).describe()
    assert d2["n_lowered_ops"] <= 2 * n  # at most one matrix + one merged phase per wire
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc2.ops, enc2.n_ops, reg, mode="dense")
    assert np.max(np.abs(emu_simulate(n, enc2, reg) - ref)) < 1e-13


def test_controlled_gates_absorb_their_target_neighbours():
    """merge_ctrl: U2 . CNot . U1 on the target becomes one dual-matrix op (U2.X.U1 where the control holds, U2.U1
    elsewhere) - with the control among the register, thread and tile-index bits, for Toffoli, and for phase gates on the
    target; CNot.CNot vanishes.  Anchor: the reference applies every gate on its own (src/circuit/simulation.rs:37-56)."""
    emu = __import__("helpers").emu_lib()
    n = 12
    c = OracleCircuit.new(n)
    c.add_gate(G.Ry(0.3), 5).add_gate(G.CNot(2), 5).add_gate(G.Rx(1.2), 5)
    desc = qb.Plan(n, encode_gates(c.circuit_gates, n), lib=emu).describe()
    assert desc["n_lowered_ops"] == 1 and desc["n_dual_ops"] == 1
    c = OracleCircuit.new(n)
    c.add_gate(G.CNot(2), 5).add_gate(G.H, 7).add_gate(G.CNot(2), 5)  # X.X on the target: nothing left but the H
    desc = qb.Plan(n, encode_gates(c.circuit_gates, n), lib=emu).describe()
    assert desc["n_lowered_ops"] == 1 and desc["n_dual_ops"] == 0
    rng = np.random.default_rng(11)
    for trial in range(6):
        c = OracleCircuit.new(n)
        for _ in range(60):
            w = [int(x) for x in rng.permutation(n)[:3]]
            k = int(rng.integers(0, 7))
            th = float(rng.uniform(-3, 3))
            if k == 0:
                c.add_gate(G.CNot(w[0]), w[1])
            elif k == 1:
                c.add_gate(G.Toffoli(w[0], w[1]), w[2])
            elif k == 2:
                c.add_gate(G.CZ(w[0]), w[1])
            else:
                c.add_gate([G.H, G.Rx(th), G.Ry(th), G.Rz(th)][k - 3], w[int(rng.integers(0, 3))])
        enc = encode_gates(c.circuit_gates, n)
        reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        reg /= np.linalg.norm(reg)
        ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
        for tile_bits, low_bits in [(0, 0), (6, 2), (8, 3), (10, 3)]:
            out, desc = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=low_bits, describe=True)
            assert desc["n_dual_ops"] > 0
            assert np.max(np.abs(out - ref)) < 1e-12, (trial, tile_bits, low_bits)


def test_permutation_rounds_for_runs_of_x_gates():
    """A run of X / CNot / Toffoli gates on more than four targets becomes one gather through the tile (ROUND_PERM):
    targets and controls on any tile bit, controls outside the tile per tile; first, middle and last round of a pass
    (direct store).  Anchor: the reference applies every gate on its own (src/circuit/simulation.rs:37-56); Grover's
    ancilla V-chain (tests/grovers.rs:75-155 writes it with Custom multi-CNOTs) is the workload this is for."""
    n = 13
    rng = np.random.default_rng(21)
    seen_perm = 0
    for trial in range(8):
        c = OracleCircuit.new(n)
        for block in range(3):
            if block != 1 or trial % 2:
                for _ in range(4):  # something that is not a permutation in front / between / behind
                    c.add_gate([G.H, G.Ry(0.7), G.Rz(0.4), G.T][int(rng.integers(0, 4))], int(rng.integers(0, n)))
            for _ in range(int(rng.integers(6, 20))):
                w = [int(x) for x in rng.permutation(n)[:3]]
                k = int(rng.integers(0, 3))
                if k == 0:
                    c.add_gate(G.X, w[0])
                elif k == 1:
                    c.add_gate(G.CNot(w[0]), w[1])
                else:
                    c.add_gate(G.Toffoli(w[0], w[1]), w[2])
        enc = encode_gates(c.circuit_gates, n)
        reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        reg /= np.linalg.norm(reg)
        ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
        for tile_bits, low_bits in [(0, 0), (8, 3), (10, 3), (11, 3), (12, 4)]:
            out, desc = emu_simulate(n, enc, reg, tile_bits=tile_bits, low_bits=low_bits, describe=True)
            assert np.max(np.abs(out - ref)) < 1e-12, (trial, tile_bits, low_bits)
            seen_perm += sum(1 for p in desc["passes"] for r in p["rounds"] if r["type"] == 2)
    assert seen_perm > 0


@pytest.mark.parametrize("n,world,tile_bits,low_bits,mode,tma", [(16, 1, 6, 2, 2, 0), (16, 1, 8, 3, 1, 0), (17, 1, 11, 3, 2, 1), (16, 1, 11, 4, 2, 1),
                                                               (17, 2, 7, 3, 2, 0), (18, 4, 11, 3, 2, 1), (19, 8, 11, 3, 2, 1), (18, 8, 7, 2, 2, 0)])
def test_prefix_folded_over_the_top_local_qubits(n, world, tile_bits, low_bits, mode, tma, monkeypatch):
    """Leading gates on the top qubits of a basis state - the rank-id qubits and up to 14 local ones - see a product state:
    the host applies them to the 2^(g+k) non-zero amplitudes (plan.cpp prefix_amplitudes) and the first pass synthesises its
    tiles from that table (PassInit::amp_tbl) instead of reading the register; QFT-33 drops from 4 passes to 2.  Checked here
    on small registers (QSV_PREFIX_MIN_LOCAL=0) for QFT and for random circuits (support bits inside and outside the first
    pass's tile), fused initialisation in every mode and through the tensor-map walk; buffers start as NaN."""
    import ctypes as C
    from helpers import emu_lib, emu_simulate_sharded
    monkeypatch.setenv("QSV_PREFIX_MIN_LOCAL", "0")
    monkeypatch.setenv("QSV_PREFIX_KEEP_BITS", "8")
    lib = emu_lib()
    lib.qsv_emu_run_plan_fused_init.restype = C.c_int
    lib.qsv_emu_run_plan_fused_init.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64, C.c_uint64, C.c_uint32]
    lib.qsv_emu_set_tma_mode(tma)
    try:
        rng = np.random.default_rng(n * 7 + world)
        g = world.bit_length() - 1
        nl = n - g
        folded = 0
        for trial in range(4):
            x = int(rng.integers(0, 1 << n))
            if trial == 0:
                c = qft_circuit(OracleCircuit, G, n)
            else:
                c = OracleCircuit.new(n)
                for w in range(n):  # a column of one-wire gates (config 3's first layer), then anything
                    c.add_gate([G.H, G.Rx(0.3 + w), G.Ry(1.1 * w), G.Rz(0.7)][int(rng.integers(0, 4))], w)
                c2 = random_any_gate_circuit(OracleCircuit, G, n, 30, rng)
                c.circuit_gates.extend(c2.circuit_gates)
            enc = encode_gates(c.circuit_gates, n)
            reg = np.zeros(1 << n, dtype=np.complex128)
            reg[x] = 1.0
            ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=2)
            plan = qb.Plan(n, enc, n_local=nl, tile_bits=tile_bits, low_bits=low_bits, free_layout=True, lib=lib)
            desc = plan.describe()
            folded += desc["prefix_local_bits"]
            only_passes = all(k == "pass" for k, _ in plan.steps())
            if only_passes and plan.layout(False) == list(range(n)):
                shards = []
                for r in range(world):
                    amps = np.full(1 << nl, np.nan + 1j * np.nan, dtype=np.complex128)
                    assert lib.qsv_emu_run_plan_fused_init(plan.handle, amps.ctypes.data_as(C.POINTER(C.c_double)), r, x, mode) == 0
                    shards.append(amps)
                from helpers import physical_index_table
                out = np.concatenate(shards)[physical_index_table(n, plan.layout(True))]
                assert np.max(np.abs(out - ref)) < 1e-12, (trial, "fused")
            plan.close()
            # the same plan with the prefix amplitudes written to the register first (what materialize + the scatter kernel do)
            out2, _, _ = emu_simulate_sharded(n, enc, world, basis_index=x, tile_bits=tile_bits, low_bits=low_bits)
            assert np.max(np.abs(out2 - ref)) < 1e-12, (trial, "scattered")
        assert folded > 0
    finally:
        lib.qsv_emu_set_tma_mode(0)


@pytest.mark.parametrize("n,world", [(16, 1), (18, 2), (17, 4), (20, 1), (19, 8)])
def test_wide_prefix_runs_as_a_plan_on_the_support_qubits(n, world, monkeypatch):
    """Prefixes wider than the host table (more than 14 local qubits: QFT-33 folds 25 of its 33 stages) are applied on the
    device, as a plan of their own on a sub-register of the support qubits (plan.cpp build_prefix_subplan: controls on the
    constant low bits resolved against the basis state, their phases folded into constants).  That plan's final state must
    be the table the host computes (prefix_amplitudes)."""
    import ctypes as C
    from helpers import emu_lib
    monkeypatch.setenv("QSV_PREFIX_MIN_LOCAL", "0")
    monkeypatch.setenv("QSV_PREFIX_KEEP_BITS", "8")
    lib = emu_lib()
    lib.qsv_emu_prefix_subplan_table.restype = C.c_int
    lib.qsv_emu_prefix_subplan_table.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_double), C.c_size_t]
    rng = np.random.default_rng(n + world)
    nl = n - (world.bit_length() - 1)
    for trial in range(4):
        if trial == 0:
            c = qft_circuit(OracleCircuit, G, n)
        else:
            c = OracleCircuit.new(n)
            for w in range(n):
                c.add_gate([G.H, G.Rx(0.3 + w), G.Ry(1.1 * w), G.Rz(0.7)][int(rng.integers(0, 4))], w)
            c.circuit_gates.extend(random_any_gate_circuit(OracleCircuit, G, n, 30, rng).circuit_gates)
        enc = encode_gates(c.circuit_gates, n)
        x = int(rng.integers(0, 1 << n))
        plan = qb.Plan(n, enc, n_local=nl, free_layout=True, lib=lib, tile_bits=6, low_bits=2)
        assert plan.prefix_local_bits() > 0
        want = plan.initial_amplitudes(x)
        got = np.zeros(max(len(want), 16), dtype=np.complex128)
        assert lib.qsv_emu_prefix_subplan_table(plan.handle, x, got.ctypes.data_as(C.POINTER(C.c_double)), len(got)) == 0
        assert np.max(np.abs(got[:len(want)] - want)) < 1e-13
        assert abs(np.sum(np.abs(want) ** 2) - 1.0) < 1e-12
        plan.close()


@pytest.mark.parametrize("open_control", [None, 4])
def test_wide_multicnot_compact_columns(open_control):
    """multicnot::<15> (tests/grovers.rs:157-172 generalised): 15 wires go to the scheduler as compact columns (qsv.h,
    iparam = 1) and come back as one controlled op - no dense round, no 13-wire limit."""
    from workloads import wide_multicnot_circuit
    n = 15
    c, expect = wide_multicnot_circuit(OracleCircuit, G, st, n, open_control)
    enc = encode_gates(c.circuit_gates, n)
    assert enc.ops[enc.n_ops - 1].iparam == 1
    out, desc = emu_simulate(n, enc, None, describe=True)
    want = np.zeros(1 << n, dtype=np.complex128)
    for k, v in expect.items():
        want[k] = v
    assert np.max(np.abs(out - want)) < 1e-14
    assert all(r["type"] != 1 for p in desc["passes"] for r in p["rounds"])  # no dense round (qsv_types.h ROUND_DENSE = 1)
