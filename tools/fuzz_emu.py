"""CPU fuzzer: random circuits (every standard gate kind + Custom closures of several shapes) scheduled by the product's
plan.cpp and executed by the host emulation of the kernel's per-thread code (tests/emu), on 1..8 emulated ranks, from
basis states and uploaded registers, against the CPU oracle.  Test infrastructure only.

    python tools/fuzz_emu.py [first_seed] [n_seeds]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import (OracleCircuit, emu_simulate, emu_simulate_sharded, emu_simulate_sharded_overlapped, encode_gates, orc, qb,  # noqa: E402
                     random_any_gate_circuit, st)

G = qb.Gate
refused = []
NMIN, NMAX = int(os.environ.get("FUZZ_NMIN", 1)), int(os.environ.get("FUZZ_NMAX", 13))


def random_custom(rng, n):
    """One Custom gate of a random shape: dense unitary, permutation with phases, multi-controlled pair gate, partial
    (None on some inputs), non-unitary."""
    shape = int(rng.integers(0, 6))
    k = int(rng.integers(1, min(n, 10 if shape in (2, 5) else 5) + 1))  # structured closures may be wide
    wires = [int(w) for w in rng.permutation(n)[:k]]
    target, controls = wires[-1], wires[:-1]
    dim = 1 << k
    if shape == 0:
        m = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        q, _ = np.linalg.qr(m)
        cols = {i: q[:, i].copy() for i in range(dim)}
    elif shape == 1:
        perm = rng.permutation(dim)
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi, size=dim))
        cols = {}
        for i in range(dim):
            v = np.zeros(dim, dtype=np.complex128)
            v[perm[i]] = ph[i]
            cols[i] = v
    elif shape == 2:  # acts only when every control is 1: 2x2 unitary on the target
        m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
        q, _ = np.linalg.qr(m)
        cols = {}
        base = dim - 2
        for b in range(2):
            v = np.zeros(dim, dtype=np.complex128)
            v[base] = q[0, b]
            v[base + 1] = q[1, b]
            cols[base + b] = v
    elif shape == 5:  # multi-controlled gate that fires on a random control pattern: a phased flip of the target
        pat = int(rng.integers(0, dim >> 1)) << 1
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi, size=2))
        cols = {}
        for b in range(2):
            v = np.zeros(dim, dtype=np.complex128)
            v[pat | (b ^ 1)] = ph[b]
            cols[pat | b] = v
    elif shape == 3:  # some inputs untouched (None), the others a dense image
        cols = {}
        for i in range(dim):
            if rng.random() < 0.5:
                v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
                cols[i] = v / np.linalg.norm(v)
    else:  # non-unitary: sparse random images
        cols = {}
        for i in range(dim):
            if rng.random() < 0.7:
                v = np.zeros(dim, dtype=np.complex128)
                for _ in range(int(rng.integers(1, 3))):
                    v[int(rng.integers(0, dim))] = rng.normal() + 1j * rng.normal()
                cols[i] = v

    def closure(prod, cols=cols, k=k):
        idx = 0
        for q in prod.get_qubits():
            idx = (idx << 1) | (1 if q == st.Qubit.One else 0)
        v = cols.get(idx)
        if v is None:
            return None
        return st.SuperPosition.new_with_amplitudes_unchecked(list(v))

    return G.Custom(closure, controls, "F"), target


def build(rng, n, n_gates, p_custom):
    c = random_any_gate_circuit(OracleCircuit, G, n, n_gates, rng)
    gates = list(c.circuit_gates)
    c2 = OracleCircuit.new(n)
    pos = 0
    # replay column by column, inserting Custom gates between columns
    cols = [gates[i:i + n] for i in range(0, len(gates), n)]
    for col in cols:
        c2.add_gates(col)
        if rng.random() < p_custom:
            g, t = random_custom(rng, n)
            c2.add_gate(g, t)
        pos += 1
    return c2


def one(seed):
    try:
        _one(seed)
    finally:
        for key in ("QSV_PREFIX_MIN_LOCAL", "QSV_PREFIX_KEEP_BITS"):
            os.environ.pop(key, None)


def _one(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(NMIN, NMAX + 1))
    n_gates = int(rng.integers(1, 100))
    c = build(rng, n, n_gates, p_custom=float(rng.choice([0.0, 0.1, 0.4])))
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    basis = int(rng.integers(0, 1 << n))
    breg = np.zeros(1 << n, dtype=np.complex128)
    breg[basis] = 1
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense")
    ref0 = orc.simulate(n, enc.ops, enc.n_ops, None, mode="dense")
    refb = orc.simulate(n, enc.ops, enc.n_ops, breg, mode="dense")
    scale = max(1.0, float(np.max(np.abs(ref))), float(np.max(np.abs(ref0))), float(np.max(np.abs(refb))))
    tol = 1e-11 * scale
    if seed % 2:  # fold leading gates on the top local qubits of a basis state on these small registers too (plan.cpp build_plan)
        os.environ["QSV_PREFIX_MIN_LOCAL"] = "4"
        os.environ["QSV_PREFIX_KEEP_BITS"] = str(int(rng.integers(0, 8)))
    cfgs = [(0, 0, True)] + [(int(rng.integers(4, 14)), int(rng.integers(1, 4)), bool(rng.integers(0, 2))) for _ in range(2)]
    for tb, lb, fuse in cfgs:
        try:
            out = emu_simulate(n, enc, reg, tile_bits=tb, low_bits=lb, fuse=fuse)
        except qb._ffi.QsvError as e:  # a forced tile smaller than a dense Custom gate: documented refusal
            assert tb and e.code == qb._ffi.ERR_UNSUPPORTED and "wider than the tile" in str(e), (seed, str(e))
            continue
        assert np.max(np.abs(out - ref)) < tol, ("single/register", seed, n, tb, lb, fuse)
        out = emu_simulate(n, enc, None, tile_bits=tb, low_bits=lb, fuse=fuse)
        assert np.max(np.abs(out - ref0)) < tol, ("single/zero", seed, n, tb, lb, fuse)
        if n < 4:  # the harness cannot address padded registers
            continue
        out, _, _ = emu_simulate_sharded(n, enc, 1, basis_index=basis, tile_bits=tb, low_bits=lb)  # folded prefix, one device
        assert np.max(np.abs(out - refb)) < tol, ("single/basis", seed, n, tb, lb, basis)
    for world in (2, 4, 8):
        g = world.bit_length() - 1
        if n - g < 4:
            continue
        tb, lb = int(rng.integers(4, 9)), int(rng.integers(1, 4))
        try:
            out, _, _ = emu_simulate_sharded(n, enc, world, basis_index=basis, tile_bits=tb, low_bits=lb)
            assert np.max(np.abs(out - refb)) < tol, ("sharded/basis", seed, n, world, tb, lb, basis)
            out, _, _ = emu_simulate_sharded(n, enc, world, register=reg, tile_bits=tb, low_bits=lb)
            assert np.max(np.abs(out - ref)) < tol, ("sharded/register", seed, n, world, tb, lb)
            if n - g >= 8:
                out, _, _ = emu_simulate_sharded_overlapped(n, enc, world, register=reg, tile_bits=min(tb, 7), low_bits=lb,
                                                            log2_slices=int(rng.integers(1, 4)), rng=rng)
                assert np.max(np.abs(out - ref)) < tol, ("sharded/overlapped", seed, n, world, tb, lb)
        except qb._ffi.QsvError as e:
            if e.code != qb._ffi.ERR_UNSUPPORTED:
                raise
            # a dense Custom gate wider than the shard cannot be made local: a documented refusal, not a wrong answer
            assert "rank" in str(e) or "local" in str(e) or "Custom" in str(e) or "shard" in str(e), (seed, str(e))
            refused.append(str(e))


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    bad = 0
    if os.environ.get("FUZZ_TMA"):
        from helpers import emu_lib
        emu_lib().qsv_emu_set_tma_mode(1)
    for s in range(first, first + count):
        try:
            one(s)
        except AssertionError as e:
            bad += 1
            print("FAIL", e.args, flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("ERROR", s, type(e).__name__, e, flush=True)
    print(f"seeds {first}..{first + count - 1}: {bad} failures, {len(refused)} sharded refusals: {sorted(set(refused))}")
    sys.exit(1 if bad else 0)
