"""Sharded registers: the state is split on the top log2(P) index bits, gates on qubits held in the
rank id trigger a global-qubit remap (EXCHANGE step).  The reference has no distributed code, so correctness is
self-consistency: the sharded plan must reproduce the single-rank oracle amplitude for amplitude.

CPU-only: the product's scheduler produces the plan; the kernel's per-thread code runs in the host emulation
(tests/emu); the exchange is done either in-process (all ranks emulated) or between two gloo ranks.
"""
import os
import sys

import numpy as np
import pytest

from helpers import (OracleCircuit, emu_lib, emu_simulate_sharded, emu_simulate_sharded_overlapped, encode_gates, orc, qb, qft_circuit, qft_expected,
                     random_any_gate_circuit, random_layered_circuit)
from quantr_b200 import _ffi as F

G = qb.Gate


@pytest.mark.parametrize("n,world", [(8, 2), (9, 4), (10, 8), (12, 4)])
def test_qft_needs_no_remap_from_a_basis_state(n, world):
    """QFT-n on P ranks from a basis state: the first log2 P stages act on the qubits in the rank id while the state is
    still a product state, so the scheduler folds them into the ranks' initial amplitudes; every later stage is local.
    Canonical layout, no EXCHANGE."""
    enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    for x in (0, 5, (1 << n) - 3):
        out, plan, n_exchanges = emu_simulate_sharded(n, enc, world, basis_index=x, tile_bits=5, low_bits=2)
        assert n_exchanges == 0
        assert np.max(np.abs(out - qft_expected(n, x))) < 1e-13
    assert plan.layout(False) == list(range(n)) and plan.layout(True) == list(range(n))
    assert plan.describe()["prefix_ops"] > 0
    amps = plan.initial_amplitudes(5)
    assert abs(np.sum(np.abs(amps) ** 2) - 1.0) < 1e-14 and np.all(np.abs(np.abs(amps) - 1 / np.sqrt(len(amps))) < 1e-14)


@pytest.mark.parametrize("n,world", [(8, 2), (9, 4), (10, 8), (12, 4)])
def test_qft_needs_exactly_one_remap_without_prefix_folding(n, world, monkeypatch):
    """The same with the folding switched off: every H target must be local once -> one EXCHANGE, the
    last-targeted qubits start in the rank id (free initial layout)."""
    monkeypatch.setenv("QSV_FOLD_PREFIX", "0")
    enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    for x in (0, 5, (1 << n) - 3):
        out, plan, n_exchanges = emu_simulate_sharded(n, enc, world, basis_index=x, tile_bits=5, low_bits=2)
        assert n_exchanges == 1
        assert np.max(np.abs(out - qft_expected(n, x))) < 1e-13
    g = world.bit_length() - 1
    lay = plan.layout(False)
    assert sorted(lay) == list(range(n))
    assert sorted(lay[b] for b in range(g)) == list(range(n - g, n))  # the last-targeted qubits start in the rank id
    st = plan.stats()
    assert st["n_exchanges"] == 1 and st["exchange_bytes"] == (16 << (n - g)) * (world - 1) // world


@pytest.mark.parametrize("seed", range(10))
def test_random_circuits_sharded_match_oracle(seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(6, 12))
    world = int(2 ** rng.integers(1, 3))
    c = random_any_gate_circuit(OracleCircuit, G, n, 60, rng)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    out, _, _ = emu_simulate_sharded(n, enc, world, tile_bits=5, low_bits=2)
    assert np.max(np.abs(out - ref)) < 1e-12
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref2 = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    out2, plan2, _ = emu_simulate_sharded(n, enc, world, register=reg, tile_bits=5, low_bits=2)
    assert plan2.layout(False) == list(range(n))  # an uploaded register keeps the canonical layout
    assert np.max(np.abs(out2 - ref2)) < 1e-12


def test_layered_circuit_sharded():
    n, world = 12, 4
    c = random_layered_circuit(OracleCircuit, G, n, 5, seed=30)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    out, plan, n_exchanges = emu_simulate_sharded(n, enc, world, tile_bits=6, low_bits=2)
    assert np.max(np.abs(out - ref)) < 1e-12
    assert n_exchanges >= 1


def test_diagonal_only_circuit_needs_no_remap():
    """Controls and diagonal gates on qubits in the rank id never move data."""
    n, world = 9, 4
    c = OracleCircuit.new(n)
    for w in range(2, n):
        c.add_gate(G.H, w)
    for w in range(n):
        c.add_gate(G.Rz(0.1 * (w + 1)), w)
    c.add_gate(G.CNot(0), 5).add_gate(G.Toffoli(0, 1), 6).add_gate(G.CRk(3, 1), 0).add_gate(G.CZ(0), 1)
    enc = encode_gates(c.circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[0b110000000] = 1
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    out, _, n_exchanges = emu_simulate_sharded(n, enc, world, register=reg, tile_bits=5, low_bits=2)
    assert n_exchanges == 0
    assert np.max(np.abs(out - ref)) < 1e-12


def _gloo_worker(rank, world, port, n, seed, result_dir):
    """One process per rank: local passes through the emulator, EXCHANGE steps as pairwise send/recv over gloo."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import emu_lib, logical_to_physical

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = emu_lib()
    g = world.bit_length() - 1
    nl = n - g
    rng = np.random.default_rng(seed)
    c = random_any_gate_circuit(OracleCircuit, G, n, 50, rng)
    for w in range(n):  # make sure every qubit is targeted at least once -> at least one remap
        c.add_gate(G.H, w)
    enc = encode_gates(c.circuit_gates, n)
    plan = qb.Plan(n, enc, n_local=nl, tile_bits=5, low_bits=2, free_layout=True, lib=lib)
    lay0 = plan.layout(False)
    shard = np.zeros(1 << nl, dtype=np.complex128)
    phys0 = logical_to_physical(3, lay0)
    shard[phys0 & ((1 << nl) - 1)] = plan.initial_amplitudes(3)[rank]
    for kind, arg in plan.steps():
        if kind == "pass":
            assert lib.qsv_emu_run_pass(plan.handle, arg, shard.ctypes.data_as(C.POINTER(C.c_double)), rank) == 0
            continue
        # rank bit j <-> local bit arg[j]: the block whose partner bits spell `peer` is swapped with the peer's block that spells `rank`
        idx = np.arange(1 << nl, dtype=np.uint64)
        spelled = np.zeros_like(idx)
        for j, p in enumerate(arg):
            spelled |= ((idx >> np.uint64(p)) & np.uint64(1)) << np.uint64(j)
        for step in range(1, world):
            peer = rank ^ step
            sel = np.nonzero(spelled == peer)[0]
            send = torch.from_numpy(np.ascontiguousarray(shard[sel]).view(np.float64))
            recv = torch.empty_like(send)
            reqs = [dist.isend(send, peer), dist.irecv(recv, peer)]
            for r in reqs:
                r.wait()
            shard[sel] = recv.numpy().view(np.complex128)
    np.save(os.path.join(result_dir, f"shard{rank}.npy"), shard)
    if rank == 0:
        np.save(os.path.join(result_dir, "layout.npy"), np.array(plan.layout(True)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 9), (4, 10)])
def test_gloo_ranks_exchange_matches_oracle(world, n, tmp_path):
    """world_size-2/4 gloo run of the N>1 host path: plan steps + pairwise amplitude exchange between processes."""
    import torch.multiprocessing as mp
    from helpers import physical_index_table
    port = 29500 + (os.getpid() % 2000) + world
    seed = 77
    mp.spawn(_gloo_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    g = world.bit_length() - 1
    full = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    layout = [int(x) for x in np.load(tmp_path / "layout.npy")]
    out = full[physical_index_table(n, layout)]
    rng = np.random.default_rng(seed)
    c = random_any_gate_circuit(OracleCircuit, G, n, 50, rng)
    for w in range(n):
        c.add_gate(G.H, w)
    enc = encode_gates(c.circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[3] = 1
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    assert np.max(np.abs(out - ref)) < 1e-12


@pytest.mark.parametrize("g,n_local,partners", [(1, 5, [4]), (1, 6, [0]), (2, 6, [1, 4]), (2, 5, [3, 4]), (3, 7, [0, 3, 6]), (3, 6, [3, 4, 5])])
def test_peer_memory_exchange_index_math(g, n_local, partners):
    """The NVLink peer-memory exchange (peer_swap.h: what peer_swap_kernel runs per pair of ranks, walked here by the
    emulation harness) against the reference semantics of an EXCHANGE step, at world sizes 2, 4 and 8."""
    import ctypes as C
    from helpers import emu_lib, exchange_bits_global
    lib = emu_lib()
    lib.qsv_emu_peer_exchange.restype = C.c_int
    lib.qsv_emu_peer_exchange.argtypes = [C.POINTER(C.POINTER(C.c_double)), C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32]
    world = 1 << g
    rng = np.random.default_rng(100 * g + n_local)
    shards = [(rng.standard_normal(1 << n_local) + 1j * rng.standard_normal(1 << n_local)).astype(np.complex128) for _ in range(world)]
    want = exchange_bits_global([s.copy() for s in shards], n_local, partners)
    ptrs = (C.POINTER(C.c_double) * world)(*[s.ctypes.data_as(C.POINTER(C.c_double)) for s in shards])
    part = (C.c_uint8 * g)(*partners)
    assert lib.qsv_emu_peer_exchange(ptrs, n_local, part, g) == 0
    for r in range(world):
        assert np.array_equal(shards[r], want[r]), f"rank {r}"


def test_shards_below_four_qubits_are_rejected():
    """Shards smaller than one register group (2^4 amplitudes) are padded with idle index bits that would collide with
    the rank bits: the scheduler refuses them instead of producing wrong amplitudes."""
    lib = emu_lib()
    for n, world in ((4, 2), (5, 4), (6, 8)):
        enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
        with pytest.raises(F.QsvError) as e:
            qb.Plan(n, enc, n_local=n - (world.bit_length() - 1), lib=lib)
        assert e.value.code == F.ERR_UNSUPPORTED


@pytest.mark.parametrize("seed", range(6))
def test_tiny_shards_find_exchange_partners(seed):
    """n_local = 4 with up to 3 rank bits and 3-wire gates: exchange partners must be found below the top-10 window."""
    rng = np.random.default_rng(900 + seed)
    g = int(rng.integers(1, 4))
    n = 4 + g
    c = random_any_gate_circuit(OracleCircuit, G, n, 40, rng)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    out, _, _ = emu_simulate_sharded(n, enc, 1 << g, tile_bits=4, low_bits=1)
    assert np.max(np.abs(out - ref)) < 1e-12


@pytest.mark.parametrize("n,world,iterations", [(10, 2, 2), (12, 4, 1), (13, 8, 1)])
def test_grover_from_native_gates_sharded(n, world, iterations):
    """BASELINE configs[4]'s Grover workload (tests/workloads.py: Toffoli V-chain + CZ, no Custom gate) on emulated ranks:
    closed-form amplitudes, the oracle, several remaps per circuit and Toffoli controls held in the rank id."""
    from workloads import grover_circuit, grover_expected_amplitudes
    c, info = grover_circuit(OracleCircuit, G, n, iterations=iterations)
    enc = encode_gates(c.circuit_gates, n)
    ref = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    _, _, probe = grover_expected_amplitudes(info, iterations, n_probe=256)
    assert np.max(np.abs(ref[np.array(probe["indices"])] - np.array(probe["expect"]))) < 1e-13
    out, plan, n_exchanges = emu_simulate_sharded(n, enc, world, tile_bits=6, low_bits=3)
    assert np.max(np.abs(out - ref)) < 1e-12
    assert n_exchanges >= 2


@pytest.mark.parametrize("seed", range(8))
def test_pipelined_exchange_slices_match_oracle(seed):
    """The pipelined remap (state_api.cu run_overlapped; north_star: global-qubit swaps overlapped with local fused
    passes): the pass before a remap, the remap and the pass after it run slice by slice along index bits none of the three
    touches.  Emulated here with the product's own slice enumeration (tma_tile.h slice_tile_id), swap index arithmetic
    (peer_swap.h) and group decision (plan.cpp plan_overlap_group), slices in a scrambled order."""
    rng = np.random.default_rng(100 + seed)
    world = [2, 4, 8][seed % 3]
    g = world.bit_length() - 1
    n = 13 + g + seed % 2
    c = random_any_gate_circuit(OracleCircuit, G, n, 90, rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    total = pipelined = 0
    for tile_bits, low_bits, k in [(6, 2, 1), (7, 3, 2), (6, 3, 3)]:
        out, n_exch, n_over = emu_simulate_sharded_overlapped(n, enc, world, register=reg, tile_bits=tile_bits, low_bits=low_bits, log2_slices=k, rng=rng)
        assert np.max(np.abs(out - ref)) < 1e-12, (world, n, tile_bits, low_bits, k)
        total += n_exch
        pipelined += n_over
    assert total > 0 and pipelined > 0
