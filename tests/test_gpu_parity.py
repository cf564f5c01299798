"""Parity tests proper: the sm_100a path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): amplitudes within 1e-12 max-abs of the oracle on identical circuits;
measure_all bin counts pass a chi-square test against the oracle's probabilities.
"""
import os

import numpy as np
import pytest

from golden import reference_vectors as rv
from helpers import (OracleCircuit, encode_gates, orc, qb, qft_circuit, qft_expected, random_any_gate_circuit,
                     random_layered_circuit, st)
from quantr_b200 import _ffi as F

pytestmark = pytest.mark.gpu
G, Circuit = qb.Gate, qb.Circuit
TOL = 1e-12


def device_run(n, enc, reg=None, **opts):
    s = qb.DeviceState(n)
    for k, v in opts.items():
        s.set_option(k, v)
    if reg is not None:
        s.upload(reg)
    stats = s.apply(enc)
    out = s.download()
    s.close()
    return out, stats


@pytest.mark.parametrize("vec", rv.VECTORS, ids=[v["name"] for v in rv.VECTORS])
def test_reference_golden_vectors_on_device(vec):
    """The reference's own tests, replayed through Circuit::simulate -> libqsv.so."""
    amps = vec["build"](Circuit, G, st).simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps - np.array(vec["expect"]))) < TOL


@pytest.mark.parametrize("seed", range(16))
def test_random_circuits_match_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 19))
    c = random_any_gate_circuit(OracleCircuit, G, n, int(rng.integers(1, 200)), rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    for opts in ({}, {"tile_bits": 8, "low_bits": 2}, {"tile_bits": 13}, {"fuse": 0, "tile_bits": 10}):
        out, _ = device_run(n, enc, reg, **opts)
        assert np.max(np.abs(out - ref)) < TOL, (n, opts)


@pytest.mark.parametrize("seed", range(6))
def test_random_x_gate_runs_and_merged_controls_match_oracle(seed):
    """Runs of X / CNot / Toffoli gates on many wires (permutation rounds, ROUND_PERM) next to rotations that merge into the
    controlled gates around them (dual-matrix ops): controls on register, thread and tile-index bits, first / middle / last
    round of a pass.  Anchor: the reference applies every gate on its own (src/circuit/simulation.rs:37-56)."""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(12, 18))
    c = OracleCircuit.new(n)
    for block in range(4):
        for _ in range(int(rng.integers(0, 8))):
            w = [int(x) for x in rng.permutation(n)[:2]]
            th = float(rng.uniform(-3, 3))
            k = int(rng.integers(0, 6))
            if k < 4:
                c.add_gate([G.H, G.Rx(th), G.Ry(th), G.Rz(th)][k], w[0])
            elif k == 4:
                c.add_gate(G.CZ(w[0]), w[1])
            else:
                c.add_gate(G.CRk(int(rng.integers(2, 6)), w[0]), w[1])
        for _ in range(int(rng.integers(5, 24))):
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
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    for opts in ({}, {"tile_bits": 8, "low_bits": 3}, {"tile_bits": 11, "low_bits": 3}, {"tile_bits": 12, "low_bits": 4}, {"tile_bits": 13}):
        out, _ = device_run(n, enc, reg, **opts)
        assert np.max(np.abs(out - ref)) < TOL, (n, opts)


@pytest.mark.parametrize("n", [20, 22, 24])
def test_layered_circuit_matches_oracle_at_scale(n):
    """BASELINE config 3's generator (H/Rx/Ry/Rz/CNot/Toffoli) at sizes the dense oracle finishes in seconds."""
    c = random_layered_circuit(OracleCircuit, G, n, 8, seed=30)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=8)
    out, stats = device_run(n, enc)
    assert np.max(np.abs(out - ref)) < TOL
    assert stats["n_passes"] < stats["n_gates"]


def test_qft16_config2_matches_oracle_and_closed_form():
    """BASELINE config 2: QFT on 16 qubits from |0xACE1>, plus the variant with a 3-wire Custom QFT."""
    n, x = 16, 0xACE1
    c = qft_circuit(Circuit, G, n, x)
    enc = encode_gates(c.circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[x] = 1
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    amps = c.simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps - ref)) < TOL
    assert np.max(np.abs(amps - qft_expected(n, x))) < TOL
    # variant: the last three wires' sub-QFT as one Gate::Custom exactly as tests/qft.rs:27-28,51-67
    c2 = Circuit.new(n)
    for pos in range(n - 3):
        c2.add_gate(G.H, pos)
        for k in range(2, n - pos + 1):
            c2.add_gate(G.CRk(k, pos + k - 1), pos)
    c2.add_gate(G.Custom(rv.make_qft_closure(Circuit, G), [13, 14], "QFT"), 15)
    c2.change_register(st.ProductState.binary_basis(x, n))
    amps2 = c2.simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps2 - ref)) < TOL


@pytest.mark.parametrize("n", [26, 28])
def test_qft_closed_form_large(n):
    """Sizes beyond the oracle's reach in test time: closed-form QFT amplitudes on random indices + norm."""
    x = 0x123456789 & ((1 << n) - 1)
    c = qft_circuit(OracleCircuit, G, n)
    enc = encode_gates(c.circuit_gates, n)
    s = qb.DeviceState(n)
    s.init_basis(x)
    stats = s.apply(enc)
    rng = np.random.default_rng(n)
    idx = rng.integers(0, 1 << n, size=4096, dtype=np.uint64)
    got = s.gather(idx)
    rev = np.zeros_like(idx)
    for b in range(n):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
    phase = np.array([(int(r) * x) % (1 << n) for r in rev], dtype=np.float64) * (2 * np.pi / (1 << n))
    expect = (np.cos(phase) + 1j * np.sin(phase)) / np.sqrt(float(1 << n))
    assert np.max(np.abs(got - expect)) < TOL
    assert abs(s.norm_sqr() - 1.0) < 1e-12
    assert stats["n_passes"] <= 4
    s.close()


def test_x3sudoko_on_device():
    """tests/grovers.rs:75-155 through the device path, incl. the reference's own assertions."""
    qb.seed(0)
    sim = rv.build_x3sudoko(Circuit, G, st).simulate()
    sim.print_warnings(True)
    amps = sim.get_state().take().get_amplitudes()
    ref = rv.build_x3sudoko(OracleCircuit, G, st).simulate().get_state().take().get_amplitudes()
    assert np.max(np.abs(amps - ref)) < TOL
    bins = sim.measure_all(5000).take()
    assert sum(bins.values()) == 5000
    for state, count in bins.items():
        if state.to_string()[:6] in rv.SUDOKU_SOLUTIONS:
            assert count > 150
        else:
            assert count < 150


@pytest.mark.parametrize("open_control", [None, 5])
def test_wide_multi_controlled_custom_gate_on_device(open_control):
    """multicnot::<17> (tests/grovers.rs:157-172 generalised): a Custom gate on 17 wires is passed as compact columns
    (qsv.h, iparam = 1) and lowered to one controlled op of the fused pass - no dense round, no 13-wire limit.  Run twice:
    the second call takes the handle's plan cache."""
    from workloads import wide_multicnot_circuit
    n = 17
    c, expect = wide_multicnot_circuit(Circuit, G, st, n, open_control)
    enc = encode_gates(c.circuit_gates, n)
    s = qb.DeviceState(n)
    try:
        for _ in range(2):
            s.init_basis(0)
            stats = s.apply(enc)
            amps = s.download()
            want = np.zeros(1 << n, dtype=np.complex128)
            for k, v in expect.items():
                want[k] = v
            assert np.max(np.abs(amps - want)) < TOL
            assert stats["n_passes"] >= 1
    finally:
        s.close()


def test_grovers_config1_measure_all():
    """BASELINE config 1 (examples/grovers.rs) + the assertions of tests/grovers.rs:62-69."""
    qb.seed(0)
    sim = rv.build_grovers_3qubit(Circuit, G, st).simulate()
    bins = sim.measure_all(500).take()
    for state, count in bins.items():
        assert state.to_string() in ("011", "111") and count > 200
    sim1 = rv.build_example_grovers(Circuit, G, st).simulate()
    p = np.abs(sim1.get_state().take().get_amplitudes()) ** 2
    assert abs(p[6] - 0.5) < TOL and abs(p[7] - 0.5) < TOL


def chi_square_p(counts, probs, shots):
    """Pearson chi-square with bins of expected count < 5 merged (SURVEY.md 8d protocol)."""
    from scipy import stats as sps
    expected = probs * shots
    order = np.argsort(expected)
    obs_m, exp_m, acc_o, acc_e = [], [], 0.0, 0.0
    for i in order:
        acc_o += counts[i]
        acc_e += expected[i]
        if acc_e >= 5:
            obs_m.append(acc_o)
            exp_m.append(acc_e)
            acc_o = acc_e = 0.0
    if acc_e > 0 and exp_m:
        obs_m[-1] += acc_o
        exp_m[-1] += acc_e
    obs_m, exp_m = np.array(obs_m), np.array(exp_m)
    chi2 = np.sum((obs_m - exp_m) ** 2 / exp_m)
    return float(sps.chi2.sf(chi2, len(obs_m) - 1))


@pytest.mark.parametrize("n", [5, 12, 14])
def test_measure_all_chi_square_and_exact_inverse_cdf(n):
    rng = np.random.default_rng(7 + n)
    c = random_any_gate_circuit(OracleCircuit, G, n, 60, rng)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    s = qb.DeviceState(n)
    s.apply(enc)
    shots = 200000
    u = rng.random(shots)
    got = s.sample(u)
    assert not np.any(got == np.uint64(F.UINT64_MAX))
    counts = np.bincount(got.astype(np.int64), minlength=1 << n).astype(np.float64)
    probs = np.abs(ref) ** 2
    assert chi_square_p(counts, probs / probs.sum(), shots) > 1e-3
    # the same uniforms through the reference's sequential rule: identical except within rounding of a boundary
    want = orc.measure_all(n, s.download(), u)
    assert np.mean(got != want) < 1e-4
    s.close()


def test_sampling_large_register_chunked_scan():
    """n = 27: 2^15 block sums, scanned by the chunked (multi-CTA) scan.  After H on every wire |amp|^2 = 2^-n exactly,
    so the sample for u is floor(u * 2^n) up to the rounding of the cumulative sums; the total must be 1."""
    n = 27
    c = OracleCircuit.new(n)
    for w in range(n):
        c.add_gate(G.H, w)
    s = qb.DeviceState(n)
    s.apply(encode_gates(c.circuit_gates, n))
    u = np.random.default_rng(27).random(20000)
    u[:4] = [0.0, 0.5, 0.999999999, 1.0 - 2.0 ** -40]
    got = s.sample(u).astype(np.float64)
    assert np.max(np.abs(got - np.floor(u * float(1 << n)))) <= 2.0
    assert abs(s.norm_sqr() - 1.0) < 1e-12
    s.close()


def test_failed_collapse_and_strictness_on_device():
    """Non-unitary register (total probability 0.25): shots with u >= total return UINT64_MAX."""
    s = qb.DeviceState(2)
    s.upload(np.array([0.5, 0, 0, 0], dtype=np.complex128))
    got = s.sample(np.array([0.1, 0.2499, 0.25, 0.9]))
    assert [int(g) for g in got] == [0, 0, F.UINT64_MAX, F.UINT64_MAX]
    assert abs(s.norm_sqr() - 0.25) < 1e-15
    s.upload(np.array([0.5, 0.5, 0.5, 0.5], dtype=np.complex128))
    got = s.sample(np.array([0.0, 0.2499999, 0.25, 0.5, 0.74, 0.75, 0.999999]))
    assert [int(g) for g in got] == [0, 0, 1, 2, 2, 3, 3]
    s.close()


def test_post_select_example_and_measure_all_without_cache():
    """examples/post_select.rs:39-49 (non-unitary Custom) and simulated_circuit.rs:81-114."""
    def post_select(prod):
        if prod.get_qubits()[0] == st.Qubit.Zero:
            return st.SuperPosition.new_with_amplitudes_unchecked([np.sqrt(2.0), 0.0])
        return st.SuperPosition.new_with_amplitudes_unchecked([0.0, 0.0])

    c = Circuit.new(2)
    c.add_gate(G.H, 0).add_gate(G.H, 1).add_gate(G.Custom(post_select, [], "P"), 1)
    sim = c.simulate()
    sim.print_warnings(True)
    assert np.allclose(sim.get_state().take().get_amplitudes(), [np.sqrt(0.5), 0, np.sqrt(0.5), 0], atol=1e-15)
    qb.seed(1)
    bins = sim.measure_all_without_cache(40).take()
    assert set(k.to_string() for k in bins) <= {"00", "10"} and sum(bins.values()) == 40
    assert sim.resimulated_shots == 0  # a deterministic closure encodes to the same ops every shot: the register is reused


def test_measure_all_without_cache_resimulates_stochastic_closures():
    """simulated_circuit.rs:81-114: 'potentially allows for mixed states to be simulated, through the implementation of
    Gate::Custom' - a closure that flips its wire with probability 1/4 per *shot* must give a 3:1 mixture of |0> and |1>, which
    only comes out if every shot whose closure differs re-runs the circuit; shots whose ops are byte-identical to the ones
    that built the register in HBM reuse it."""
    rng = np.random.default_rng(12)
    state = {"flip": False, "calls": 0}

    def noisy(prod):
        if state["calls"] % 2 == 0:  # one decision per shot (the closure is called once per basis sub-state)
            state["flip"] = rng.random() < 0.25
        state["calls"] += 1
        one = prod.get_qubits()[0] == st.Qubit.One
        out = (not one) if state["flip"] else one
        return st.SuperPosition.new_with_amplitudes_unchecked([0.0, 1.0] if out else [1.0, 0.0])

    c = Circuit.new(3)
    c.add_gate(G.Custom(noisy, [], "N"), 1)
    sim = c.simulate()
    sim.print_warnings(False)
    qb.seed(3)
    shots = 400
    bins = {k.to_string(): v for k, v in sim.measure_all_without_cache(shots).take().items()}
    assert set(bins) <= {"000", "010"} and sum(bins.values()) == shots
    assert 60 <= bins.get("010", 0) <= 140  # binomial(400, 1/4): 100 +- 4 sigma
    assert 0 < sim.resimulated_shots < shots  # only the shots whose closure changed its mind


def _none_on_zero_closure(prod):
    """None on |..0>, (|..0>+|..1>)/sqrt2 on |..1> of the LAST qubit of the sub-state: the untouched state |0> is also
    part of the image of |1>, the case where the reference's overwrite rule (simulation.rs:120-133) differs from a
    linear map."""
    q = prod.get_qubits()
    if q[-1] == st.Qubit.Zero:
        return None
    amps = np.zeros(1 << len(q), dtype=np.complex128)
    base = 0
    for b in q[:-1]:
        base = (base << 1) | (1 if b == st.Qubit.One else 0)
    amps[base << 1] = amps[(base << 1) | 1] = np.sqrt(0.5)
    return st.SuperPosition.new_with_amplitudes_unchecked(amps)


def test_none_overwrite_rule_on_device():
    """App. B.4 on the GPU: input (|0>+|1>)/sqrt2 -> [0.7071, 0.5] (the additive reading would give [1.2071, 0.5])."""
    c = Circuit.new(1)
    c.add_gate(G.H, 0).add_gate(G.Custom(_none_on_zero_closure, [], "N"), 0)
    sim = c.simulate()
    assert np.max(np.abs(sim.get_state().take().get_amplitudes() - np.array([np.sqrt(0.5), 0.5]))) < 1e-15


@pytest.mark.parametrize("tile_bits", [8, 13])
def test_none_overwrite_rule_three_wire_custom_at_n12(tile_bits):
    """The same closure as a 3-wire Custom (two controls) inside a 12-qubit random circuit, device vs the oracle's
    faithful restatement of simulation.rs:120-133, for a tile smaller than and as large as the register."""
    n = 12
    rng = np.random.default_rng(1204)
    c = random_any_gate_circuit(OracleCircuit, G, n, 30, rng)
    c.add_gate(G.Custom(_none_on_zero_closure, [3, 9], "N3"), 6)
    for w in (0, 5, 11):
        c.add_gate(G.Ry(0.3 + w), w)
    c.add_gate(G.Custom(_none_on_zero_closure, [11, 0], "N3b"), 4)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="faithful")
    assert np.max(np.abs(orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense") - ref)) < 1e-14
    out, _ = device_run(n, enc, reg, tile_bits=tile_bits)
    assert np.max(np.abs(out - ref)) < TOL


def test_config3_depth100_live_oracle():
    """BASELINE config 3's generator at its full depth (100 layers) against the live dense oracle at n = 22."""
    import os
    n = 22
    c = random_layered_circuit(OracleCircuit, G, n, 100, seed=30)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=os.cpu_count() or 1)
    out, stats = device_run(n, enc)
    assert np.max(np.abs(out - ref)) < TOL
    assert stats["n_gates"] == enc.n_ops


@pytest.mark.parametrize("n", [26, 28])
def test_config3_depth100_against_oracle_fixture(n):
    """Config 3 at depth 100 and n = 26 / 28: 4,096 amplitudes + norm against the dense oracle's result, computed once
    by tests/golden/make_config3_fixtures.py (the oracle needs ~7 / ~30 minutes at these sizes) and committed."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"config3_n{n}_d100.npz")
    if not os.path.exists(path):
        pytest.skip(f"fixture {path} not generated")
    fx = np.load(path)
    c = random_layered_circuit(OracleCircuit, G, n, 100, seed=30)
    enc = encode_gates(c.circuit_gates, n)
    assert enc.n_ops == int(fx["n_gates"])
    s = qb.DeviceState(n)
    s.apply(enc)
    got = s.gather(fx["indices"])
    assert np.max(np.abs(got - fx["amps"])) < TOL
    assert abs(s.norm_sqr() - float(fx["norm_sqr"])) < 1e-11
    s.close()


def test_config3_full_size_three_schedules_agree():
    """Config 3 at its stated size (n = 30, depth 100, 4,000 gates; 16 GiB register): no CPU result can exist, so three
    independent schedules (tile 11, tile 12, one gate per pass) must agree on 4,096 gathered amplitudes and the norm."""
    n = 30
    c = random_layered_circuit(OracleCircuit, G, n, 100, seed=30)
    enc = encode_gates(c.circuit_gates, n)
    assert enc.n_ops == 4000
    idx = np.random.default_rng(30).integers(0, 1 << n, size=4096, dtype=np.uint64)
    results = []
    for opts in ({"tile_bits": 11}, {"tile_bits": 12}, {"fuse": 0}):
        s = qb.DeviceState(n)
        for k, v in opts.items():
            s.set_option(k, v)
        s.apply(enc)
        results.append((s.gather(idx), s.norm_sqr()))
        s.close()
    for got, norm in results:
        assert abs(norm - 1.0) < 1e-11
        assert np.max(np.abs(got - results[0][0])) < TOL


def test_two_handles_from_two_threads():
    """include/qsv.h: different handles may be used from different threads.  Two threads each run their own
    circuits (different tile sizes, so different kernel instantiations and launch parameters) against the oracle."""
    import threading
    errors = []

    def work(seed, tile_bits):
        try:
            rng = np.random.default_rng(seed)
            for it in range(6):
                n = 14 + (it % 3)
                c = random_any_gate_circuit(OracleCircuit, G, n, 120, rng)
                enc = encode_gates(c.circuit_gates, n)
                ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
                out, _ = device_run(n, enc, None, tile_bits=tile_bits)
                err = float(np.max(np.abs(out - ref)))
                if err >= TOL:
                    errors.append((seed, it, err))
        except Exception as e:  # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=work, args=(s, t)) for s, t in ((1, 10), (2, 12), (3, 11), (4, 10))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("n,iterations", [(16, 3), (24, 2)])
def test_grover_from_native_gates(n, iterations):
    """The Grover workload of bench.py --workload grover (tests/workloads.py) on one GPU: the oracle at n = 16, the closed
    form on the (marked, unmarked) plane at both sizes."""
    from workloads import grover_circuit, grover_expected_amplitudes
    c, info = grover_circuit(OracleCircuit, G, n, iterations=iterations)
    enc = encode_gates(c.circuit_gates, n)
    s = qb.DeviceState(n)
    s.apply(enc)
    am, ao, probe = grover_expected_amplitudes(info, iterations)
    got = s.gather(np.array(probe["indices"], dtype=np.uint64))
    assert np.max(np.abs(got - np.array(probe["expect"]))) < TOL
    assert abs(s.norm_sqr() - 1.0) < 1e-12
    if n <= 16:
        ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense", threads=4)
        assert np.max(np.abs(s.download() - ref)) < TOL
    s.close()


def test_plan_rerun_and_range_access():
    n = 18
    c = qft_circuit(OracleCircuit, G, n)
    enc = encode_gates(c.circuit_gates, n)
    plan = qb.Plan(n, enc)
    s = qb.DeviceState(n)
    outs = []
    for x in (1, 77777):
        s.init_basis(x)
        s.run_plan(plan)
        outs.append(s.download(1000, 5000))
        assert np.max(np.abs(outs[-1] - qft_expected(n, x)[1000:6000])) < TOL
    s.close()


def test_checkpoint_round_trip(tmp_path):
    """qsv_save / qsv_load: a state carried from one circuit to the next through a file equals running them back to back."""
    n = 14
    rng = np.random.default_rng(14)
    c1 = random_any_gate_circuit(OracleCircuit, G, n, 80, rng)
    c2 = random_any_gate_circuit(OracleCircuit, G, n, 80, rng)
    e1, e2 = encode_gates(c1.circuit_gates, n), encode_gates(c2.circuit_gates, n)
    s = qb.DeviceState(n)
    s.apply(e1)
    path = str(tmp_path / "state.qsv")
    s.save(path)
    assert os.path.getsize(path) == 256 + (16 << n)
    s.apply(e2)
    want = s.download()
    s.close()
    t = qb.DeviceState(n)
    t.load(path)
    t.apply(e2)
    assert np.max(np.abs(t.download() - want)) == 0.0
    with pytest.raises(F.QsvError):
        qb.DeviceState(n + 1).load(path)
    t.close()


def test_error_paths_on_device():
    s = qb.DeviceState(3)
    with pytest.raises(F.QsvError):
        s.init_basis(8)
    with pytest.raises(F.QsvError):
        s.download(4, 8)
    with pytest.raises(F.QsvError):
        s.set_option("tile_bits", 99)
    with pytest.raises(F.QsvError):
        s.gather([9])
    s.close()
