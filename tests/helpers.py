"""Shared test plumbing: oracle- and emulator-backed variants of the host `Circuit`.

`OracleCircuit` keeps the product's builder (column layout, validation, encoding are host
logic under test) but runs the gate list through the CPU oracle instead of the device, so the
reference's tests can be replayed on a CPU-only box and used as the checker on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import quantr_b200 as qb  # noqa: E402
from quantr_b200 import _ffi as F  # noqa: E402
from quantr_b200 import states as st  # noqa: E402
from quantr_b200.circuit import Measurement, encode_gates  # noqa: E402
from oracle import oracle as orc  # noqa: E402


class _HostSimulated:
    """SimulatedCircuit look-alike over a host amplitude vector (oracle / emulator results)."""

    def __init__(self, gates, n, amps):
        self.circuit_gates, self.num_qubits, self.amps = gates, n, amps

    def get_state(self):
        return Measurement.NonObservable(st.SuperPosition._raw(self.amps, self.num_qubits))

    take_state = get_state

    def measure_all(self, shots, rng=None):
        rng = rng or np.random.default_rng(0)
        idx = orc.measure_all(self.num_qubits, self.amps, rng.random(shots))
        bins = {}
        for i in idx:
            if int(i) == F.UINT64_MAX:
                continue
            key = st.ProductState.binary_basis(int(i), self.num_qubits)
            bins[key] = bins.get(key, 0) + 1
        return Measurement.Observable(bins)


class OracleCircuit(qb.Circuit):
    mode = "dense"

    @staticmethod
    def new(n):
        return OracleCircuit(n)

    def _simulate(self, gates, register):
        enc = encode_gates(gates, self.num_qubits)
        reg = None if register is None else register.get_amplitudes()
        amps = orc.simulate(self.num_qubits, enc.ops, enc.n_ops, reg, mode=self.mode)
        return _HostSimulated(gates, self.num_qubits, amps)


class FaithfulOracleCircuit(OracleCircuit):
    mode = "faithful"

    @staticmethod
    def new(n):
        return FaithfulOracleCircuit(n)


_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libqsv_emu.so"))
        lib.qsv_emu_run_plan.restype = C.c_int
        lib.qsv_emu_run_plan.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64]
        lib.qsv_emu_alloc_qubits.restype = C.c_uint32
        lib.qsv_emu_alloc_qubits.argtypes = [C.c_void_p]
        lib.qsv_plan_create.restype = C.c_int
        lib.qsv_plan_create.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.POINTER(F.QsvOp), C.c_size_t,
                                        C.c_uint32, C.c_uint32, C.c_int]
        lib.qsv_plan_destroy.argtypes = [C.c_void_p]
        lib.qsv_emu_run_pass.restype = C.c_int
        lib.qsv_emu_run_pass.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.c_uint64]
        vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
        lib.qsv_plan_create_ex.restype = C.c_int
        lib.qsv_plan_create_ex.argtypes = [C.POINTER(vp), u32, u32, C.POINTER(F.QsvOp), sz, u32, u32, i32, C.POINTER(C.c_uint8), i32]
        lib.qsv_plan_num_steps.argtypes = [vp, C.POINTER(sz)]
        lib.qsv_plan_get_step.argtypes = [vp, sz, C.POINTER(i32), C.POINTER(u32), C.POINTER(C.c_uint8), sz]
        lib.qsv_plan_get_layout.argtypes = [vp, i32, C.POINTER(C.c_uint8), sz]
        lib.qsv_plan_stats.argtypes = [vp, C.POINTER(F.QsvStats)]
        lib.qsv_plan_last_error.restype = C.c_char_p
        lib.qsv_plan_serialize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.qsv_emu_set_tma_mode.argtypes = [C.c_int]
        lib.qsv_emu_tma_passes.restype = C.c_uint32
        lib.qsv_emu_tma_max_boxes.restype = C.c_uint32
        _emu = lib
    return _emu


def emu_simulate(n, enc, register=None, *, tile_bits=0, low_bits=0, fuse=True, n_local=None, rank=0, describe=False):
    """Schedules `enc` with the product's scheduler and executes the pass blobs with the host
    emulation of the kernel's per-thread code (tests/emu/qsv_emu.cpp)."""
    import json
    lib = emu_lib()
    plan = C.c_void_p()
    nl = n if n_local is None else n_local
    rc = lib.qsv_plan_create(C.byref(plan), n, nl, enc.ops, enc.n_ops, tile_bits, low_bits, 1 if fuse else 0)
    if rc != 0:
        raise F.QsvError(rc, lib.qsv_plan_last_error().decode())
    try:
        n_alloc = lib.qsv_emu_alloc_qubits(plan)
        amps = np.zeros(1 << n_alloc, dtype=np.complex128)
        if register is None:
            if rank == 0:
                amps[0] = 1.0
        else:
            reg = np.asarray(register, dtype=np.complex128)
            amps[: reg.shape[0]] = reg
        assert lib.qsv_emu_run_plan(plan, amps.ctypes.data_as(C.POINTER(C.c_double)), rank) == 0
        out = amps[: 1 << nl].copy()
        if describe:
            size = C.c_size_t()
            lib.qsv_plan_serialize(plan, None, 0, C.byref(size))
            buf = C.create_string_buffer(size.value)
            lib.qsv_plan_serialize(plan, buf, size.value, C.byref(size))
            return out, json.loads(buf.raw[: size.value].decode())
        return out
    finally:
        lib.qsv_plan_destroy(plan)


class EmuCircuit(qb.Circuit):
    tile_bits = 0
    low_bits = 0

    @staticmethod
    def new(n):
        return EmuCircuit(n)

    def _simulate(self, gates, register):
        enc = encode_gates(gates, self.num_qubits)
        reg = None if register is None else register.get_amplitudes()
        amps = emu_simulate(self.num_qubits, enc, reg, tile_bits=self.tile_bits, low_bits=self.low_bits)
        return _HostSimulated(gates, self.num_qubits, amps)


# ---- workload generators ---------------------------------------------------

def qft_circuit(C, G, n, x=None):
    """QFT as the reference writes it (tests/qft.rs:55-62): H(pos) then CRk(k, pos+k-1), no final swaps."""
    c = C.new(n)
    for pos in range(n):
        c.add_gate(G.H, pos)
        for k in range(2, n - pos + 1):
            c.add_gate(G.CRk(k, pos + k - 1), pos)
    if x is not None:
        c.change_register(st.ProductState.binary_basis(x, n))
    return c


def qft_expected(n, x):
    """Closed form: amp[y] = 2^{-n/2} exp(2 pi i x bitrev_n(y) / 2^n)."""
    y = np.arange(1 << n, dtype=np.uint64)
    rev = np.zeros_like(y)
    for b in range(n):
        rev |= ((y >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
    phase = (rev.astype(object) * x) % (1 << n)
    ang = np.array([float(p) for p in phase]) * (2.0 * np.pi / (1 << n))
    return (np.cos(ang) + 1j * np.sin(ang)) / np.sqrt(float(1 << n))


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)


def random_layered_circuit(C, G, n, depth, seed=None):
    """BASELINE config 3 generator: per layer one column of random H/Rx/Ry/Rz, then
    n//3 CNot/Toffoli on a shuffled wire order."""
    rng = SplitMix64(n if seed is None else seed)
    c = C.new(n)
    for _ in range(depth):
        col = []
        for _q in range(n):
            kind = rng.next() % 4
            theta = 2.0 * np.pi * (rng.next() >> 11) / float(1 << 53)
            col.append([G.H, G.Rx(theta), G.Ry(theta), G.Rz(theta)][kind])
        c.add_gates(col)
        perm = list(range(n))
        for i in range(n - 1, 0, -1):
            j = rng.next() % (i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        for j in range(n // 3):
            a, b, t = perm[3 * j], perm[3 * j + 1], perm[3 * j + 2]
            if rng.next() % 4 == 0:
                c.add_gate(G.Toffoli(a, b), t)
            else:
                c.add_gate(G.CNot(a), b)
    return c


ALL_SINGLE = ["H", "X", "Y", "Z", "S", "Sdag", "T", "Tdag", "X90", "Y90", "MX90", "MY90"]


def random_any_gate_circuit(C, G, n, n_gates, rng, custom=None):
    """Random circuit over every standard gate kind (for differential tests)."""
    c = C.new(n)
    for _ in range(n_gates):
        r = rng.integers(0, 24 if n >= 3 else (22 if n >= 2 else 16))
        wires = rng.permutation(n)
        t = int(wires[0])
        theta = float(rng.uniform(-2 * np.pi, 2 * np.pi))
        if r < 12:
            g = getattr(G, ALL_SINGLE[r])
        elif r == 12:
            g = G.Rx(theta)
        elif r == 13:
            g = G.Ry(theta)
        elif r == 14:
            g = G.Rz(theta)
        elif r == 15:
            g = G.Phase(theta)
        elif r == 16:
            g = G.CR(theta, int(wires[1]))
        elif r == 17:
            g = G.CRk(int(rng.integers(1, 8)), int(wires[1]))
        elif r == 18:
            g = G.CZ(int(wires[1]))
        elif r == 19:
            g = G.CY(int(wires[1]))
        elif r == 20:
            g = G.CNot(int(wires[1]))
        elif r == 21:
            g = G.Swap(int(wires[1]))
        else:
            g = G.Toffoli(int(wires[1]), int(wires[2]))
        c.add_gate(g, t)
    return c


# ---- sharded registers on the CPU: the product's plan (layouts, passes, EXCHANGE steps) driven by the emulator --------

def logical_to_physical(i, layout):
    p = 0
    for b, pos in enumerate(layout):
        p |= ((i >> b) & 1) << pos
    return p


def physical_index_table(n, layout):
    """phys[i] = physical index of canonical index i under `layout` (vectorised)."""
    idx = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(idx)
    for b, pos in enumerate(layout):
        phys |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(pos)
    return phys


def exchange_bits_global(shards, n_local, partners):
    """Reference semantics of an EXCHANGE step on the list of per-rank shards: rank bit j <-> local bit partners[j]."""
    g = len(partners)
    full = np.concatenate(shards)  # physical order: rank bits on top
    n = n_local + g
    idx = np.arange(1 << n, dtype=np.uint64)
    src = idx.copy()
    for j, p in enumerate(partners):
        a, b = np.uint64(n_local + j), np.uint64(p)
        ba, bb = (src >> a) & np.uint64(1), (src >> b) & np.uint64(1)
        diff = ba ^ bb
        src ^= (diff << a) | (diff << b)
    out = full[src]
    return [out[r << n_local:(r + 1) << n_local].copy() for r in range(1 << g)]


def emu_simulate_sharded(n, enc, world, *, basis_index=0, register=None, tile_bits=0, low_bits=0):
    """Runs a sharded plan on `world` emulated ranks in this process.  Returns the canonical-order state vector."""
    lib = emu_lib()
    g = world.bit_length() - 1
    nl = n - g
    plan = qb.Plan(n, enc, n_local=nl, tile_bits=tile_bits, low_bits=low_bits, free_layout=register is None, lib=lib)
    lay0, lay1 = plan.layout(False), plan.layout(True)
    assert lib.qsv_emu_alloc_qubits(plan.handle) == nl  # shards are never padded (sharded registers need >= 4 local qubits)
    full = np.zeros(1 << n, dtype=np.complex128)
    if register is None:
        # every rank starts from its own amplitude at the basis state's local index (one rank holds 1 unless the plan
        # folded the circuit's leading gates on the rank-id qubits into the initial state)
        k = plan.prefix_local_bits()  # the folded prefix may also cover the top k local bits
        low = logical_to_physical(basis_index, lay0) & ((1 << (nl - k)) - 1)
        for j, a in enumerate(plan.initial_amplitudes(basis_index)):
            full[(j << (nl - k)) | low] = a
    else:
        assert lay0 == list(range(n))
        full[:] = register
    shards = [full[r << nl:(r + 1) << nl].copy() for r in range(world)]
    n_exchanges = 0
    for kind, arg in plan.steps():
        if kind == "pass":
            for r in range(world):
                assert lib.qsv_emu_run_pass(plan.handle, arg, shards[r].ctypes.data_as(C.POINTER(C.c_double)), r) == 0
        else:
            shards = exchange_bits_global(shards, nl, arg)
            n_exchanges += 1
    phys = physical_index_table(n, lay1)
    out = np.concatenate(shards)[phys]
    return out, plan, n_exchanges


def emu_simulate_sharded_overlapped(n, enc, world, *, register, tile_bits=0, low_bits=0, log2_slices=2, rng=None):
    """emu_simulate_sharded with every EXCHANGE step run the way state_api.cu run_overlapped pipelines it: the shard is cut
    into slices along bits that neither the remap nor its neighbouring passes touch (plan_overlap_group); the pass
    before the remap, the remap and the pass after it each run slice by slice - here in a scrambled order that only
    respects 'a slice is exchanged after every rank has passed it and before any rank passes it again'.
    Returns (canonical state vector, number of remaps, number of remaps that were pipelined)."""
    lib = emu_lib()
    rng = rng or np.random.default_rng(0)
    g = world.bit_length() - 1
    nl = n - g
    plan = qb.Plan(n, enc, n_local=nl, tile_bits=tile_bits, low_bits=low_bits, lib=lib)
    assert plan.layout(False) == list(range(n))
    shards = [np.array(register[r << nl:(r + 1) << nl], dtype=np.complex128) for r in range(world)]
    steps = plan.steps()
    sliceable = (C.c_uint8 * len(steps))(*[1 if k == "pass" else 0 for k, _ in steps])
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    done = [False] * len(steps)
    n_exchanges = n_overlapped = 0
    for i, (kind, arg) in enumerate(steps):
        if kind == "pass":
            continue
        n_exchanges += 1
        out = (C.c_uint32 * 6)()
        if lib.qsv_emu_overlap_group(plan.handle, i, sliceable, log2_slices, out) != 1:
            continue
        n_overlapped += 1
        steps[i] = ("group", (arg, bool(out[0]), bool(out[1]), int(out[2]), [int(out[3 + k]) for k in range(out[2])]))
        if out[0]:
            done[i - 1], sliceable[i - 1] = True, 0
        if out[1]:
            done[i + 1], sliceable[i + 1] = True, 0
    for i, (kind, arg) in enumerate(steps):
        if kind == "pass":
            if not done[i]:
                for r in range(world):
                    assert lib.qsv_emu_run_pass(plan.handle, arg, ptr(shards[r]), r) == 0
        elif kind == "exchange":
            shards = exchange_bits_global(shards, nl, arg)
        else:
            partners, slice_prev, slice_next, nb, bits = arg
            cbits = (C.c_uint8 * 3)(*(bits + [0] * (3 - nb)))
            cpart = (C.c_uint8 * len(partners))(*partners)
            order = [int(v) for v in rng.permutation(1 << nb)]
            for v in order:  # a fast slice may be exchanged and even passed again before a slow one has been touched
                if slice_prev:
                    for r in range(world):
                        assert lib.qsv_emu_run_pass_slice(plan.handle, steps[i - 1][1], ptr(shards[r]), r, nb, cbits, v) == 0
                arr = (C.POINTER(C.c_double) * world)(*[ptr(s) for s in shards])
                assert lib.qsv_emu_peer_exchange_slice(arr, nl, cpart, g, nb, cbits, v) == 0
                if slice_next:
                    for r in range(world):
                        assert lib.qsv_emu_run_pass_slice(plan.handle, steps[i + 1][1], ptr(shards[r]), r, nb, cbits, v) == 0
    phys = physical_index_table(n, plan.layout(True))
    return np.concatenate(shards)[phys], n_exchanges, n_overlapped
