"""Workload generators beyond QFT (BASELINE configs[4]: "QFT + Grover 36 qubits sharded across 8xB200").

Grover search from the reference's native gates.  The reference writes Grover with Gate::Custom multi-controlled NOTs
(/root/reference/tests/grovers.rs:75-155, multicnot at :157-172; examples/grovers.rs:20-69 for three qubits with
CZ/Toffoli); a Custom gate is a dense 2^k x 2^k closure table, which cannot exist for 19+ wires (SURVEY.md 7.2 hard
part 7).  Here the multi-controlled Z over the s search wires is a Toffoli V-chain: the AND of the first s-1 search wires
is accumulated into s-2 clean ancilla wires, a native CZ between the last ancilla and the last search wire applies the
sign, and the chain is uncomputed.  n wires = s search + (s-2) ancilla (+ 1 idle wire when n is odd):
n = 36 -> 19 search + 17 ancilla; n = 34 -> 18 + 16; n = 33 -> 17 + 15 + 1 idle.

One Grover iteration = oracle (X on the search wires where the marked item has a 0, MCZ, X again) + diffusion
(H, X, MCZ, X, H on the search wires): 4(s-2) Toffoli + 2 CZ + at most 6s single-wire gates.
"""
from __future__ import annotations

import numpy as np


def grover_layout(n):
    s = (n + 2) // 2
    a = s - 2
    return s, a, n - s - a


def _mcz(c, G, search, ancilla):
    """Z on |1..1> of the search wires through the ancilla V-chain (ancillas start and end in |0>)."""
    s = len(search)
    c.add_gate(G.Toffoli(search[0], search[1]), ancilla[0])
    for i in range(2, s - 1):
        c.add_gate(G.Toffoli(search[i], ancilla[i - 2]), ancilla[i - 1])
    c.add_gate(G.CZ(ancilla[s - 3]), search[s - 1])
    for i in range(s - 2, 1, -1):
        c.add_gate(G.Toffoli(search[i], ancilla[i - 2]), ancilla[i - 1])
    c.add_gate(G.Toffoli(search[0], search[1]), ancilla[0])


def grover_circuit(C, G, n, iterations=1, marked=None):
    """Returns (circuit, info).  Search wires are 0..s-1 (the top index bits), ancillas follow."""
    s, a, idle = grover_layout(n)
    if s < 3:
        raise ValueError("Grover workload needs at least 4 qubits")
    if marked is None:
        marked = 0x5A5A5A5A5A & ((1 << s) - 1)
    search = list(range(s))
    ancilla = list(range(s, s + a))
    zeros = [w for w in search if not (marked >> (s - 1 - w)) & 1]
    c = C.new(n)
    for w in search:
        c.add_gate(G.H, w)
    for _ in range(iterations):
        for w in zeros:
            c.add_gate(G.X, w)
        _mcz(c, G, search, ancilla)
        for w in zeros:
            c.add_gate(G.X, w)
        for w in search:
            c.add_gate(G.H, w)
        for w in search:
            c.add_gate(G.X, w)
        _mcz(c, G, search, ancilla)
        for w in search:
            c.add_gate(G.X, w)
        for w in search:
            c.add_gate(G.H, w)
    return c, {"n": n, "search": s, "ancilla": a, "idle": idle, "marked": marked}


def grover_expected_amplitudes(info, iterations, n_probe=4096, seed=99):
    """Closed form on the (marked, unmarked) plane: oracle a_m -> -a_m; diffusion = I - 2|psi><psi| (the circuit's
    H X MCZ X H, sign included): a_i -> a_i - 2*mean.  Returns (marked amplitude, unmarked amplitude,
    {"indices": canonical indices to probe, "expect": their amplitudes}); amplitudes with a non-zero ancilla are 0."""
    n, s, marked = info["n"], info["search"], info["marked"]
    big = float(1 << s)
    am = ao = 1.0 / np.sqrt(big)
    for _ in range(iterations):
        am = -am
        mean = (am + (big - 1.0) * ao) / big
        am, ao = am - 2.0 * mean, ao - 2.0 * mean
    rng = np.random.default_rng(seed)
    low = n - s
    items = rng.integers(0, 1 << s, size=n_probe // 2, dtype=np.uint64)
    items[0] = marked
    idx = [int(i) << low for i in items]
    exp = [am if int(i) == marked else ao for i in items]
    junk = rng.integers(1, 1 << low, size=n_probe - len(idx), dtype=np.uint64) if low > 0 else []
    for j, i in zip(junk, rng.integers(0, 1 << s, size=len(junk), dtype=np.uint64)):
        idx.append((int(i) << low) | int(j))
        exp.append(0.0)
    return am, ao, {"indices": idx, "expect": [complex(e) for e in exp]}


def wide_multicnot_circuit(C, G, st, n, open_control=None):
    """The reference's multicnot::<N> closure (tests/grovers.rs:157-172) on all n wires: n - 1 controls, flips the last wire
    when every control is |1>, None otherwise.  X on the controls first (all but `open_control`), H + S on the target so
    the flip is visible in both amplitudes.  Returns the circuit and its expected non-zero amplitudes {index: value}."""
    def multicnot(prod):
        q = prod.get_qubits()
        if all(x == st.Qubit.One for x in q[:-1]):
            return prod.clone().invert_digit(len(q) - 1).into_super_position()
        return None

    c = C.new(n)
    for w in range(n - 1):
        if w != open_control:
            c.add_gate(G.X, w)
    c.add_gate(G.H, n - 1).add_gate(G.S, n - 1)
    c.add_gate(G.Custom(multicnot, list(range(n - 1)), "X"), n - 1)
    base = 0
    for w in range(n - 1):
        if w != open_control:
            base |= 1 << (n - 1 - w)
    r = 0.5 ** 0.5
    a0, a1 = r, 1j * r  # H then S on |0>
    if open_control is None:
        a0, a1 = a1, a0
    return c, {base: a0, base | 1: a1}
